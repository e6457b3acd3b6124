"""One power_pairs pass over a 2^18-element BLS12-377 G1 vector (for ncu on the MSM kernels)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref as R
import snark_setup_b200 as S
from snark_setup_b200 import ffi
g = R.BLS12_377.g1
n = (1 << int(sys.argv[1] if len(sys.argv) > 1 else 18)) + 1
gen = g.encode(g.gen, False) * n
pts = S.apply_powers(S.BLS12_377, S.G1, gen, False, S.CHECK_NO, True, n, tau=0x1234567890abcdef1234567890abcdef, first_power=1)
for it in range(2):
    ffi.profile_reset(); ffi.profile_enable(True)
    t = time.perf_counter()
    out, s, sx = S.check_and_ratio(S.BLS12_377, S.G1, pts, True, seed=bytes(32))
    dt = time.perf_counter() - t
    print(round(dt * 1e3, 1), "ms", {k: round(v["ms"], 2) for k, v in ffi.profile_read().items()})
