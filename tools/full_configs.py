"""BASELINE.json configs at their FULL sizes through the host-buffer C ABI, timed, with size-independent checks.

  python tools/full_configs.py c4 [power=22] [chunks=8]   BLS12-377: contribute, verify, inverse round trip (full mode),
                                                          then the same ceremony as `chunks` chunked contributions
                                                          aggregated and compared byte for byte with the full-mode response
  python tools/full_configs.py c3 [power=21]              BW6-761: contribute + verify_transform + inverse round trip

One JSON line per measurement.  The oracle is used as the checker for a few spot elements only.
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


def blank(cv, rp):
    out = bytearray(rp.get_length(False))
    for vec, (o, c, s) in enumerate(rp.split_offsets(False)):
        g = cv.g2 if vec in (1, 4) else cv.g1
        out[o:o + c * s] = g.encode(g.gen, False) * c
    return bytes(out)


def timed(f):
    t0 = time.perf_counter()
    r = f()
    return r, time.perf_counter() - t0


def emit(**kw):
    print(json.dumps(kw), flush=True)


def spot_parity(cid, rp, chal, resp, keys):
    for vec, ((o, c, s), (oo, _, so)) in enumerate(zip(rp.split_offsets(False), rp.split_offsets(True))):
        grp = 1 if vec in (1, 4) else 0
        coeff = [None, None, keys[1], keys[2], keys[2]][vec]
        for i0 in sorted({0, c // 3, max(0, c - 2)}):
            n = min(2, c - i0)
            tau = 1 if vec == 4 else keys[0]
            want = O.apply_powers(cid, grp, chal[o + i0 * s:o + (i0 + n) * s], False, 3, True, n, tau=tau, first_power=i0,
                                  coeff=coeff)
            if bytes(resp[oo + i0 * so:oo + (i0 + n) * so]) != want:
                return False
    return True


def full_mode(name, curve, power):
    cv = R.CURVES[curve]
    cid = S.BLS12_377 if curve == "bls12_377" else S.BW6_761
    rp = R.Phase1Parameters(cv, power, 256)
    sp = S.Phase1Parameters(cid, power, 256)
    N = 1 << power
    k0 = [scalar(b"full-0-%d" % i, cv.r) for i in range(3)]
    k1 = [scalar(b"full-1-%d" % i, cv.r) for i in range(3)]
    chal = bytearray(sp.get_length(False))
    _, t0 = timed(lambda: S.phase1_computation(sp, blank(cv, rp), chal, False, False, S.CHECK_NO, *k0))
    chal = bytes(chal)
    resp = bytearray(sp.get_length(True))
    _, t_c = timed(lambda: S.phase1_computation(sp, chal, resp, False, True, S.CHECK_NO, *k1))
    ok = spot_parity(cid, rp, chal, resp, k1)
    emit(config=name, step="contribute", curve=curve, power=power, challenge_bytes=len(chal), response_bytes=len(resp),
         seconds=round(t_c, 3), powers_per_s=round(N / t_c), parity_spot_check=ok, first_call_seconds=round(t0, 3))
    newc = bytearray(sp.get_length(False))
    pairs, t_v = timed(lambda: S.phase1_verification_vectors(sp, bytes(resp), True, newc, False,
                                                             seed=hashlib.blake2b(b"rho", digest_size=32).digest()))
    tau_acc = k0[0] * k1[0] % cv.r
    vok = all(O.apply_powers(cid, g, s, False, 3, False, 1, powers=[tau_acc]) == sx for (s, sx), g in zip(pairs, (0, 1, 0, 0)))
    emit(config=name, step="verify_transform", curve=curve, power=power, seconds=round(t_v, 3), powers_per_s=round(N / t_v),
         ratio_check=vok, what="OnlyNonZero decode + subgroup check + power_pairs (s,sx) per vector + uncompressed new challenge")
    back = bytearray(sp.get_length(False))
    k1inv = [pow(x, -1, cv.r) for x in k1]
    _, t_b = timed(lambda: S.phase1_computation(sp, bytes(newc), back, False, False, S.CHECK_NO, *k1inv))
    emit(config=name, step="inverse_contribution_roundtrip", seconds=round(t_b, 3), byte_exact=bytes(back[64:]) == chal[64:])
    return cv, cid, k1, chal, bytes(resp)


def chunked(name, cv, cid, power, nchunks, keys, chal, full_resp):
    """The same contribution done as `nchunks` chunked contributions (phase1/src/objects/parameters.rs:248-294),
    aggregated (aggregation.rs:11-180): must equal the full-mode response byte for byte."""
    N = 1 << power
    chunk_size = (2 * N - 1 + nchunks - 1) // nchunks
    fullp = S.Phase1Parameters(cid, power, 256)
    agg = bytearray(fullp.get_length(True))
    t_total = 0.0
    for ci in range(nchunks):
        cp = S.Phase1Parameters(cid, power, 256, mode=S.ffi.MODE_CHUNKED, chunk_index=ci, chunk_size=chunk_size)
        cchal = S.phase1_split_chunk(cp, chal, False, False)
        cresp = bytearray(cp.get_length(True))
        _, t = timed(lambda: S.phase1_computation(cp, cchal, cresp, False, True, S.CHECK_NO, *keys))
        t_total += t
        S.phase1_aggregate_chunk(cp, bytes(cresp), True, agg, True)
    emit(config=name, step="chunked_contribute", chunks=nchunks, chunk_size=chunk_size, seconds=round(t_total, 3),
         powers_per_s=round(N / t_total), equals_full_mode=bytes(agg[64:]) == full_resp[64:])


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c4"
    if which == "c4":
        power = int(sys.argv[2]) if len(sys.argv) > 2 else 22
        nchunks = int(sys.argv[3]) if len(sys.argv) > 3 else 8
        cv, cid, k1, chal, resp = full_mode("C4", "bls12_377", power)
        chunked("C4", cv, cid, power, nchunks, k1, chal, resp)
    else:
        power = int(sys.argv[2]) if len(sys.argv) > 2 else 21
        full_mode("C3", "bw6_761", power)
