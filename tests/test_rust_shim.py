"""The Rust shim crate (rust/setup-utils-cuda) cannot be compiled in this image (no rustc), so it is kept honest
mechanically: every `extern "C"` prototype in src/ffi.rs is parsed and compared with the C prototype of
include/snark_setup_b200.h (name, argument count, integer widths, pointer constness, struct layouts), ffi.rs must be
what tools/gen_rust_ffi.py generates, every `ffi::ss_*` call in src/lib.rs must name a bound function with the right
number of arguments, and the error map must cover every ss_status the header defines."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "snark_setup_b200.h")
CRATE = os.path.join(ROOT, "rust", "setup-utils-cuda")

# independent of tools/gen_rust_ffi.py on purpose: (width class, pointer?, const?)
C_CLASS = {"int": "i32", "uint32_t": "u32", "uint64_t": "u64", "size_t": "usize", "uint8_t": "u8", "void": "void", "char": "char",
           "ss_phase1_params": "SsPhase1Params", "ss_phase1_sizes": "SsPhase1Sizes", "ss_error_info": "SsErrorInfo",
           "ss_profile_entry": "SsProfileEntry", "double": "f64"}
R_CLASS = {"c_int": "i32", "u32": "u32", "u64": "u64", "usize": "usize", "u8": "u8", "c_void": "void", "c_char": "char",
           "SsPhase1Params": "SsPhase1Params", "SsPhase1Sizes": "SsPhase1Sizes", "SsErrorInfo": "SsErrorInfo",
           "SsProfileEntry": "SsProfileEntry", "f64": "f64"}


def c_type(t):
    t = t.strip()
    ptr = t.endswith("*")
    t = t.rstrip("*").strip()
    const = t.startswith("const ")
    t = t[6:].strip() if const else t
    return (C_CLASS[t], ptr, const and ptr)


def r_type(t):
    t = t.strip()
    if t.startswith("*const "):
        return (R_CLASS[t[7:].strip()], True, True)
    if t.startswith("*mut "):
        return (R_CLASS[t[5:].strip()], True, False)
    return (R_CLASS[t], False, False)


def header_protos():
    text = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"(?m)^\s*((?:const\s+)?\w+\s*\*?)\s*(ss_\w+)\s*\(([^;{]*?)\)\s*;", text):
        ret, name, args = m.groups()
        args = " ".join(args.split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                mm = re.match(r"^(.*?)(\w+)$", a.strip())
                params.append(c_type(mm.group(1)))
        out[name] = (c_type(ret), params)
    return out


def rust_protos():
    text = open(os.path.join(CRATE, "src", "ffi.rs")).read()
    block = text[text.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (ss_\w+)\((.*?)\)(?:\s*->\s*([^;]+))?;", block, flags=re.S):
        name, args, ret = m.groups()
        params = [r_type(a.split(":", 1)[1]) for a in args.split(",") if a.strip()]
        out[name] = (r_type(ret) if ret else ("void", False, False), params)
    return out


def test_every_prototype_matches_the_header():
    h, r = header_protos(), rust_protos()
    assert len(h) >= 45
    assert sorted(h) == sorted(r), (sorted(set(h) - set(r)), sorted(set(r) - set(h)))
    for name in h:
        assert h[name][0] == r[name][0], (name, "return", h[name][0], r[name][0])
        assert len(h[name][1]) == len(r[name][1]), (name, "argument count")
        for i, (a, b) in enumerate(zip(h[name][1], r[name][1])):
            assert a == b, (name, i, a, b)


def test_struct_layouts_match_the_header():
    hdr = open(HDR).read()
    ffi = open(os.path.join(CRATE, "src", "ffi.rs")).read()

    def c_fields(name):
        body = re.search(r"typedef struct \{([^}]*)\}\s*" + name + r"\s*;", hdr).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            ty, rest = decl.split(" ", 1)
            for nm in rest.split(","):
                nm = nm.strip()
                arr = re.match(r"(\w+)\[(\d+)\]", nm)
                out.append((arr.group(1), C_CLASS[ty], int(arr.group(2))) if arr else (nm, C_CLASS[ty], 0))
        return out

    def r_fields(name):
        body = re.search(r"pub struct " + name + r" \{(.*?)\n\}", ffi, flags=re.S).group(1)
        out = []
        for m in re.finditer(r"pub (\w+): ([^,\n]+),", body):
            nm, ty = m.groups()
            arr = re.match(r"\[(\w+); (\d+)\]", ty)
            out.append((nm, R_CLASS[arr.group(1)], int(arr.group(2))) if arr else (nm, R_CLASS[ty.strip()], 0))
        return out

    assert c_fields("ss_error_info") == r_fields("SsErrorInfo")
    assert c_fields("ss_phase1_params") == r_fields("SsPhase1Params")
    assert c_fields("ss_phase1_sizes") == r_fields("SsPhase1Sizes")
    assert c_fields("ss_profile_entry") == r_fields("SsProfileEntry")


def test_ffi_rs_is_the_generated_file():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"])
    assert p.returncode == 0, "rust/setup-utils-cuda/src/ffi.rs is stale: run tools/gen_rust_ffi.py"


def _call_args(text, start):
    """number of top-level arguments of the call whose '(' is at text[start]"""
    depth, n, i, seen = 0, 0, start, False
    while True:
        c = text[i]
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
            if depth == 0:
                return n + (1 if seen else 0)
        elif c == "," and depth == 1:
            n += 1
            seen = False
        elif not c.isspace() and depth >= 1 and not (depth == 1 and c == "("):
            seen = True
        i += 1


def test_lib_rs_calls_bound_functions_with_the_right_arity():
    r = rust_protos()
    lib = open(os.path.join(CRATE, "src", "lib.rs")).read()
    lib = re.sub(r"//.*", "", lib)
    calls = list(re.finditer(r"ffi::(ss_\w+)\s*\(", lib))
    assert len(calls) >= 18
    for m in calls:
        name = m.group(1)
        assert name in r, f"lib.rs calls {name}, which ffi.rs does not bind"
        assert _call_args(lib, m.end() - 1) == len(r[name][1]), (name, _call_args(lib, m.end() - 1), len(r[name][1]))
    # the reference-facing wrappers north_star / SURVEY §8b name
    for fn in ("generate_powers_of_tau", "batch_exp", "batch_mul", "merge_pairs", "power_pairs", "check_subgroup", "same_ratio",
               "check_same_ratio", "apply_powers", "phase1_computation", "phase1_verification_vectors", "phase1_initialization",
               "groth16_params_new", "phase1_computation_shard", "phase1_verification_vectors_shard", "reduce_partial_pairs"):
        assert re.search(r"pub fn " + fn + r"\b", lib), fn


def test_error_map_covers_every_status():
    hdr = open(HDR).read()
    codes = {int(v) for v in re.findall(r"SS_ERR_\w+\s*=\s*(\d+)", hdr)}
    lib = open(os.path.join(CRATE, "src", "lib.rs")).read()
    body = lib[lib.index("fn check(rc: c_int)"):]
    body = body[:body.index("\n}\n")]
    mapped = {int(v) for v in re.findall(r"^\s*(\d+)\s*=>", body, flags=re.M)}
    # 8 (invalid argument) and 9 (device) have no setup_utils::Error variant: they take the panicking arm
    assert codes - mapped == {8, 9} and 0 in mapped


def test_crate_files_exist():
    for f in ("Cargo.toml", "build.rs", "src/ffi.rs", "src/lib.rs", "tests/parity.rs"):
        assert os.path.exists(os.path.join(CRATE, f)), f
    build = open(os.path.join(CRATE, "build.rs")).read()
    assert "rustc-link-lib=dylib=snarksetup_b200" in build and "make" in build
