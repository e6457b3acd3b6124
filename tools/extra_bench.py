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
    from snark_setup_b200 import ffi as F
    F.set_concurrent_vectors(False)  # serialised, so the event brackets see one kernel at a time
    F.profile_enable(True)
    F.profile_reset()
    S.phase1_computation(sp, chal, resp, False, True, S.CHECK_NO, *k1)
    prof_c = {k: round(v["ms"], 2) for k, v in sorted(F.profile_read().items())}
    F.profile_reset()
    S.phase1_verification_vectors(sp, bytes(resp), True, newc, False, seed=seed)
    prof_v = {k: round(v["ms"], 2) for k, v in sorted(F.profile_read().items())}
    F.profile_enable(False)
    F.set_concurrent_vectors(True)
    print(json.dumps({"bench": "bw6_761 phase1", "power": power, "contribute_powers_per_s": N / t_c, "contribute_ms": t_c * 1e3,
                      "contribute_kernels_ms_serialised": prof_c, "verify_kernels_ms_serialised": prof_v,
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


def prepare_phase2(curve, power):
    """Groth16Params::new + ::write (prepare_phase2) of a 2^power accumulator with phase2_size = 2^power."""
    from snark_setup_b200 import ffi as F
    cv = R.CURVES[curve]
    cid = S.BLS12_377 if curve == "bls12_377" else S.BW6_761
    rp = R.Phase1Parameters(cv, power, 256)
    sp = S.Phase1Parameters(cid, power, 256)
    m = 1 << power
    keys = [scalar(b"pp2-%d" % i, cv.r) for i in range(3)]
    acc = bytearray(sp.get_length(False))
    S.phase1_computation(sp, bytes(R.phase1_initialization(rp, False)), acc, False, False, S.CHECK_NO, *keys)
    acc = bytes(acc)
    out = []
    S.groth16_params_new(sp, acc, False, m, False)  # warm-up (slab allocation)
    t0 = time.perf_counter()
    out.append(S.groth16_params_new(sp, acc, False, m, False))
    t = time.perf_counter() - t0
    F.profile_enable(True)
    F.profile_reset()
    S.groth16_params_new(sp, acc, False, m, False)
    prof = {k: round(v["ms"], 3) for k, v in sorted(F.profile_read().items())}
    F.profile_enable(False)
    # forward-evaluation check of the tau_g1 coefficients at w^1: sum_j w^j coeffs_j = tau*G
    s1, s2 = cv.g1.size(False), cv.g2.size(False)
    w = R.get_root_of_unity(cv.r, m)
    rho, x = [], 1
    for _ in range(m):
        rho.append(x)
        x = x * w % cv.r
    coeffs = out[0][2 * s1 + s2:2 * s1 + s2 + m * s1]
    s_, _ = S.merge_pairs(cid, S.G1, coeffs, coeffs, False, rho=rho)
    o = rp.split_offsets(False)[0][0]
    ok = s_ == acc[o + s1:o + 2 * s1]
    smuls = 4 * ((power - 1) * m // 2 + 2)
    # CPU port: oracle.cpp::group_ifft (reference algorithm: n/2*log n + n double-and-add scalar multiplications per
    # vector, OpenMP over all host cores) timed on a 2^ks-element sample of the same vectors, scaled by the exact
    # operation count ratio (the work per scalar multiplication does not depend on n)
    ks = min(power, 10)
    o1, o2 = rp.split_offsets(False)[0][0], rp.split_offsets(False)[1][0]
    t1 = time.perf_counter()
    O.group_ifft(cid, 0, acc[o1:o1 + (s1 << ks)], False, False)
    t_g1 = time.perf_counter() - t1
    t1 = time.perf_counter()
    O.group_ifft(cid, 1, acc[o2:o2 + (s2 << ks)], False, False)
    t_g2 = time.perf_counter() - t1
    scale = (m * power / 2 + m) / ((1 << ks) * ks / 2 + (1 << ks))
    cpu_s = scale * (3 * t_g1 + t_g2)
    print(json.dumps({"bench": "prepare_phase2 (Groth16Params::new + write)", "curve": curve, "power": power,
                      "phase2_size": m, "seconds": round(t, 4), "coefficients_per_s": round(4 * m / t),
                      "scalar_muls": smuls, "output_bytes": len(out[0]), "forward_evaluation_check": ok,
                      "kernels_ms": prof,
                      "cpu_port_estimate_s": round(cpu_s, 1), "cpu_cores": O.threads(),
                      "cpu_note": "oracle.cpp::group_ifft (reference algorithm) on a 2^%d-element sample of each vector on all "
                                  "host cores, scaled by the n/2*log n + n operation count (3 G1 + 1 G2 transforms)" % ks,
                      "path": "host buffers through ss_groth16_params_new"}), flush=True)


def pairing_latency(curve):
    """Latency of the device check_same_ratio (one warp per check): 1 check and the 4 checks of a response."""
    from snark_setup_b200 import ffi as F
    cv = R.CURVES[curve]
    cid = S.BLS12_377 if curve == "bls12_377" else S.BW6_761
    g1, g2 = cv.g1, cv.g2
    x = scalar(b"pair-x", cv.r)
    p1 = g1.write_batch([g1.gen, g1.mul(g1.gen, x)], False)
    p2 = g2.write_batch([g2.gen, g2.mul(g2.gen, x)], False)
    S.check_same_ratio(cid, p1, p2)
    t1 = best(lambda: S.check_same_ratio(cid, p1, p2))
    t4 = best(lambda: S.check_same_ratio_batch(cid, p1 * 4, p2 * 4))
    t32 = best(lambda: S.check_same_ratio_batch(cid, p1 * 32, p2 * 32))
    F.profile_enable(True)
    F.profile_reset()
    S.check_same_ratio_batch(cid, p1 * 4, p2 * 4)
    prof = {k: round(v["ms"], 3) for k, v in F.profile_read().items()}
    F.profile_enable(False)
    print(json.dumps({"bench": "check_same_ratio on device (reduced pairing product, one warp per check)",
                      "curve": curve, "one_check_ms": round(t1 * 1e3, 2), "four_checks_ms": round(t4 * 1e3, 2),
                      "thirty_two_checks_ms": round(t32 * 1e3, 2), "kernels_ms": prof}), flush=True)


def qap(log2m):
    """dot_product_vec over a synthetic R1CS-shaped matrix: 2^log2m constraints and variables, 1-4 entries per
    variable (90 % +-1, 10 % general coefficients) plus the constant-one variable that touches every constraint."""
    import random
    from snark_setup_b200 import ffi as F
    cv, cid, g = R.BLS12_377, S.BLS12_377, R.BLS12_377.g1
    m = 1 << log2m
    rng = random.Random(11)
    gen = g.encode(g.gen, False) * m
    bases = S.apply_powers(cid, S.G1, gen, False, S.CHECK_NO, False, m, tau=scalar(b"qap-tau", cv.r), first_power=0)
    rows = [[(1, i) for i in range(m)]]
    for _ in range(m - 1):
        row = []
        for _ in range(rng.randrange(1, 5)):
            u = rng.random()
            c = 1 if u < 0.6 else cv.r - 1 if u < 0.9 else rng.randrange(cv.r)
            row.append((c, rng.randrange(m)))
        rows.append(row)
    nnz = sum(len(r) for r in rows)
    S.qap_dot_product(cid, S.G1, bases, False, rows[:1024], True)  # warm-up
    t0 = time.perf_counter()
    out = S.qap_dot_product(cid, S.G1, bases, False, rows, True)
    t = time.perf_counter() - t0
    F.profile_enable(True)
    F.profile_reset()
    t1 = time.perf_counter()
    S.qap_dot_product(cid, S.G1, bases, False, rows, True)
    t_prof = time.perf_counter() - t1
    prof = {k: round(v["ms"], 3) for k, v in sorted(F.profile_read().items())}
    F.profile_enable(False)
    ok = True
    for v in (1, m // 2, m - 1):
        idx = [i for _, i in rows[v]]
        pts = b"".join(bases[i * 96:(i + 1) * 96] for i in idx)
        want = O.msm(0, 0, pts, False, len(idx), [c for c, _ in rows[v]])
        ok = ok and S.transcode(cid, S.G1, want, False, S.CHECK_NO, True) == out[v * 48:(v + 1) * 48]
    print(json.dumps({"bench": "qap dot_product_vec (BLS12-377 G1)", "constraints": m, "variables": len(rows), "nnz": nnz,
                      "seconds_incl_python_marshalling": round(t, 3), "kernels_ms": prof, "device_ms": round(sum(prof.values()), 1),
                      "spot_check_vs_oracle_msm": ok, "path": "host buffers through ss_qap_dot_product"}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "qap":
        qap(int(sys.argv[2]) if len(sys.argv) > 2 else 20)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pairing":
        pairing_latency("bls12_377")
        pairing_latency("bw6_761")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "prepare_phase2":
        prepare_phase2(sys.argv[2] if len(sys.argv) > 2 else "bls12_377", int(sys.argv[3]) if len(sys.argv) > 3 else 18)
        sys.exit(0)
    bw6(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
    phase2(int(sys.argv[2]) if len(sys.argv) > 2 else 20)
