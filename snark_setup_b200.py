"""Import shim: the package directory is `snark-setup_b200/` (the name the build contract fixes),
which is not a valid Python identifier; `import snark_setup_b200` loads it under this name."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "snark-setup_b200")
_spec = importlib.util.spec_from_file_location("snark_setup_b200", os.path.join(_d, "__init__.py"),
                                               submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["snark_setup_b200"] = _mod
_spec.loader.exec_module(_mod)
