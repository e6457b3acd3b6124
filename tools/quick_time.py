"""Quick wall-clock timing of ss_apply_powers per group (host buffers), for early sizing."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref as R
import snark_setup_b200 as S

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for cid, cv in ((S.BLS12_377, R.BLS12_377), (S.BW6_761, R.BW6_761)):
    for gid, g in ((S.G1, cv.g1), (S.G2, cv.g2)):
        n = 1 << (logn if cid == S.BLS12_377 else logn - 2)
        gen = g.encode(g.gen, False) * n
        tau = 0x1234567890abcdef1234567890abcdef % cv.r
        t = time.perf_counter()
        pts = S.apply_powers(cid, gid, gen, False, S.CHECK_NO, False, n, tau=tau, first_power=1)
        t_first = time.perf_counter() - t
        best = 1e9
        for _ in range(2):
            t = time.perf_counter()
            out = S.apply_powers(cid, gid, pts, False, S.CHECK_NO, True, n, tau=tau, first_power=12345, coeff=tau + 1)
            best = min(best, time.perf_counter() - t)
        print(json.dumps({"group": g.name, "n": n, "first_s": round(t_first, 4), "best_s": round(best, 4),
                          "elems_per_s": round(n / best)}), flush=True)
