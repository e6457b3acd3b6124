"""Secondary measurements (not the driver's bench line): BW6-761 contribute/verify and the phase-2
delta^-1 batch_mul + H/L ratio MSM, through the host-buffer C ABI (copies included).  One JSON line each.

  python tools/extra_bench.py [bw6_power=16] [phase2_log2=20]
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import coracle as O  # noqa: E402  (checker only)
import pyref as R  # noqa: E402
import snark_setup_b200 as S  # noqa: E402


def scalar(label, r):
    return int.from_bytes(hashlib.blake2b(label, digest_size=64).digest(), "little") % (r - 2) + 2


def best(f, reps=3):
    f()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        t.append(time.perf_counter() - t0)
    return min(t)


def bw6(power):
    cv, cid = R.BW6_761, S.BW6_761
    rp = R.Phase1Parameters(cv, power, 256)
    sp = S.Phase1Parameters(cid, power, 256)
    N = 1 << power
    k0 = [scalar(b"bw6-0-%d" % i, cv.r) for i in range(3)]
    k1 = [scalar(b"bw6-1-%d" % i, cv.r) for i in range(3)]
    blank = bytes(R.phase1_initialization(rp, False))
    chal = bytearray(sp.get_length(False))
    S.phase1_computation(sp, blank, chal, False, False, S.CHECK_NO, *k0)
    chal = bytes(chal)
    resp = bytearray(sp.get_length(True))
    t_c = best(lambda: S.phase1_computation(sp, chal, resp, False, True, S.CHECK_NO, *k1))
    # spot parity vs the oracle: 4 tau_g1 elements
    o, _, sz = rp.split_offsets(False)[0]
    oo, _, szo = rp.split_offsets(True)[0]
    i0 = 1234 % (2 * N - 5)
    want = O.apply_powers(1, 0, chal[o + i0 * sz:o + (i0 + 4) * sz], False, 3, True, 4, tau=k1[0], first_power=i0)
    ok = bytes(resp[oo + i0 * szo:oo + (i0 + 4) * szo]) == want
    newc = bytearray(sp.get_length(False))
    seed = bytes(range(32))
    pairs = []
    t_v = best(lambda: pairs.append(S.phase1_verification_vectors(sp, bytes(resp), True, newc, False, seed=seed)))
    tau = k0[0] * k1[0] % cv.r
    vok = all(O.apply_powers(1, g, s, False, 3, False, 1, powers=[tau]) == sx for (s, sx), g in zip(pairs[-1], (0, 1, 0, 0)))
    print(json.dumps({"bench": "bw6_761 phase1", "power": power, "contribute_powers_per_s": N / t_c, "contribute_ms": t_c * 1e3,
                      "verify_powers_per_s": N / t_v, "verify_ms": t_v * 1e3, "parity_spot_check": ok, "ratio_check": vok,
                      "path": "host buffers through ss_phase1_computation / ss_phase1_verification_vectors"}), flush=True)


def phase2(log2n):
    cv, cid, g = R.BLS12_377, S.BLS12_377, R.BLS12_377.g1
    n = 1 << log2n
    gen = g.encode(g.gen, False) * n
    h_before = S.apply_powers(cid, S.G1, gen, False, S.CHECK_NO, False, n, tau=scalar(b"p2-tau", cv.r), first_power=1)
    dinv = scalar(b"p2-delta-inv", cv.r)
    buf = bytearray(h_before)
    t_mul = best(lambda: (buf.__setitem__(slice(None), h_before), S.batch_mul(cid, S.G1, buf, dinv)))
    t_copy = best(lambda: buf.__setitem__(slice(None), h_before))
    S.batch_mul(cid, S.G1, buf, dinv)
    want = O.apply_powers(0, 0, h_before[:96 * 8], False, 3, False, 8, powers=[dinv] * 8)
    ok = bytes(buf[:96 * 8]) == want
    seed = bytes(range(32))
    res = []
    t_mp = best(lambda: res.append(S.merge_pairs(cid, S.G1, h_before, bytes(buf), False, seed=seed)))
    s, sx = res[-1]
    rok = O.apply_powers(0, 0, s, False, 3, False, 1, powers=[dinv]) == sx
    print(json.dumps({"bench": "phase2 delta^-1 batch_mul + H ratio MSM (BLS12-377 G1)", "n": n,
                      "batch_mul_elems_per_s": n / (t_mul - t_copy), "batch_mul_ms": (t_mul - t_copy) * 1e3,
                      "merge_pairs_pairs_per_s": n / t_mp, "merge_pairs_ms": t_mp * 1e3, "parity_spot_check": ok,
                      "ratio_check": rok, "path": "host buffers through ss_batch_mul / ss_merge_pairs"}), flush=True)


if __name__ == "__main__":
    bw6(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
    phase2(int(sys.argv[2]) if len(sys.argv) > 2 else 20)
