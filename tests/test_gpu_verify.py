"""GPU parity for the verification side: merge_pairs / power_pairs (bucket MSM), the fused
check-subgroup + ratio + re-emit pass, and the per-vector loop of Phase1::verification.

The reference's own tests for this path (setup-utils/src/helpers.rs:334-368: power_pairs accepts a
geometric sequence and rejects a tampered one; phase1/src/verification.rs:783-921: chained
contributions verify) need a pairing for the final verdict.  The pairing stays on the host in the
product (check_same_ratio); here the verdict e(s, tau*G2) == e(sx, G2) is checked in its equivalent
discrete-log form  sx == tau * s  with the oracle's scalar multiplication, which is what the pairing
equation states for points of the prime-order subgroup."""
import os
import random

import pytest

import coracle as O
import pyref as R
import snark_setup_b200 as S

pytestmark = pytest.mark.gpu

CURVES = [(S.BLS12_377, R.BLS12_377), (S.BW6_761, R.BW6_761)]
GROUPS = [(cid, cv, gid, g) for cid, cv in CURVES for gid, g in ((S.G1, cv.g1), (S.G2, cv.g2))]
IDS = [g.name for _, _, _, g in GROUPS]


def geometric(g, n, tau, rng):
    P = g.mul(g.gen, rng.randrange(1, g.r))
    out = []
    for _ in range(n):
        out.append(P)
        P = g.mul(P, tau)
    return out


@pytest.mark.parametrize("cid,cv,gid,g", GROUPS, ids=IDS)
def test_merge_pairs_explicit_rho_matches_oracle_msm(cid, cv, gid, g):
    rng = random.Random(31 + cid * 2 + gid)
    n = 70
    v1 = [g.mul(g.gen, rng.randrange(1, g.r)) for _ in range(n)]
    v2 = [g.mul(g.gen, rng.randrange(1, g.r)) for _ in range(n)]
    v1[11] = None
    rho = [0, 1, cv.r - 1] + [rng.randrange(cv.r) for _ in range(n - 3)]
    for comp in (False, True):
        b1, b2 = g.write_batch(v1, comp), g.write_batch(v2, comp)
        s, sx = S.merge_pairs(cid, gid, b1, b2, comp, rho=rho)
        assert s == O.msm(cid, gid, b1, comp, n, rho)
        assert sx == O.msm(cid, gid, b2, comp, n, rho)
    # power_pairs = merge_pairs(v[..n-1], v[1..])
    b = g.write_batch(v1, False)
    s, sx = S.power_pairs(cid, gid, b, False, rho=rho[:-1])
    assert s == O.msm(cid, gid, b[:(n - 1) * g.usize], False, n - 1, rho[:-1])
    assert sx == O.msm(cid, gid, b[g.usize:], False, n - 1, rho[:-1])


@pytest.mark.parametrize("cid,cv,gid,g", GROUPS, ids=IDS)
def test_power_pairs_verdict(cid, cv, gid, g):
    """helpers.rs:334-368: a true geometric sequence passes the ratio check, a tampered one fails."""
    rng = random.Random(7 + gid)
    tau = rng.randrange(2, cv.r)
    n = 45
    v = geometric(g, n, tau, rng)
    seed = bytes(rng.randrange(256) for _ in range(32))
    buf = g.write_batch(v, True)
    out, s, sx = S.check_and_ratio(cid, gid, buf, True, seed=seed, out_compressed=False)
    assert out == g.write_batch(v, False)
    S_pt, SX_pt = g.decode(s, False), g.decode(sx, False)
    assert S_pt is not None and g.mul(S_pt, tau) == SX_pt
    # same seed -> same combination (device ChaCha20 is deterministic)
    _, s2, sx2 = S.check_and_ratio(cid, gid, buf, True, seed=seed)
    assert (s2, sx2) == (s, sx)
    # tamper v[1] (helpers.rs:363-367)
    v2 = list(v)
    v2[1] = g.mul(v2[1], rng.randrange(2, cv.r))
    _, s, sx = S.check_and_ratio(cid, gid, g.write_batch(v2, True), True, seed=seed)
    assert g.mul(g.decode(s, False), tau) != g.decode(sx, False)
    # swap two adjacent elements
    v3 = list(v)
    v3[20], v3[21] = v3[21], v3[20]
    _, s, sx = S.check_and_ratio(cid, gid, g.write_batch(v3, True), True, seed=seed)
    assert g.mul(g.decode(s, False), tau) != g.decode(sx, False)


@pytest.mark.parametrize("cid,cv,gid,g", GROUPS, ids=IDS)
def test_check_and_ratio_errors(cid, cv, gid, g):
    rng = random.Random(3)
    v = geometric(g, 20, 12345, rng)
    seed = bytes(32)
    bad = list(v)
    bad[9] = None
    with pytest.raises(S.PointAtInfinity) as ei:
        S.check_and_ratio(cid, gid, g.write_batch(bad, True), True, seed=seed)
    assert ei.value.index == 9
    x = 5
    while True:
        xx = x if g.F.degree == 1 else (x, 3)
        try:
            P = g.decode(g.F.to_bytes(xx, 0, 2), True, R.NO)
            if not g.in_subgroup(P):
                break
        except R.InvalidData:
            pass
        x += 1
    bad = list(v)
    bad[14] = P
    with pytest.raises(S.IncorrectSubgroup) as ei:
        S.check_and_ratio(cid, gid, g.write_batch(bad, True), True, seed=seed)
    assert ei.value.index == 14
    # SubgroupCheckMode::No skips the r-multiplication
    S.check_and_ratio(cid, gid, g.write_batch(bad, True), True, subgroup_mode=3, seed=seed)
    with pytest.raises(S.SetupError):
        S.check_and_ratio(cid, gid, g.write_batch(v[:1], True), True, seed=seed)  # BatchTooSmall


def test_multi_tile_paths(monkeypatch):
    """Force tiny device tiles so vectors span many tiles (overlap element, persistent buckets)."""
    # SS_TILE_LOG2 is read once per process: run in a subprocess
    import subprocess
    import sys
    code = r'''
import os, sys, random
sys.path.insert(0, %r); sys.path.insert(0, %r)
import pyref as R, coracle as O, snark_setup_b200 as S
rng = random.Random(5)
cv, g = R.BLS12_377, R.BLS12_377.g1
tau = rng.randrange(2, cv.r)
n = 1000
P = g.mul(g.gen, 99); v = []
for _ in range(n):
    v.append(P); P = g.mul(P, tau)
buf = g.write_batch(v, True)
rho = [rng.randrange(cv.r) for _ in range(n - 1)]
out, s, sx = S.check_and_ratio(S.BLS12_377, S.G1, buf, True, rho=rho, out_compressed=False)
assert out == g.write_batch(v, False)
ub = g.write_batch(v, False)
assert s == O.msm(0, 0, ub[:(n - 1) * 96], False, n - 1, rho)
assert sx == O.msm(0, 0, ub[96:], False, n - 1, rho)
got = S.apply_powers(S.BLS12_377, S.G1, ub, False, S.CHECK_NO, True, n, tau=tau, first_power=7)
assert got == O.apply_powers(0, 0, ub, False, 3, True, n, tau=tau, first_power=7)
print("ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
       os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    env = dict(os.environ, SS_TILE_LOG2="8")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("cid,cv", CURVES, ids=["bls12_377", "bw6_761"])
def test_phase1_contribute_then_verify(cid, cv):
    """phase1/src/verification.rs:783-921 shape: initialise -> contribute -> contribute -> verify the last
    response (compressed) into a new uncompressed challenge."""
    power, batch = 4, 8
    rng = random.Random(cid + 100)
    rp = R.Phase1Parameters(cv, power, batch)
    sp = S.Phase1Parameters(cid, power, batch)
    N = 1 << power
    k0 = [rng.randrange(2, cv.r) for _ in range(3)]
    k1 = [rng.randrange(2, cv.r) for _ in range(3)]
    acc0 = bytes(R.phase1_initialization(rp, False))
    acc1 = bytearray(sp.get_length(False))
    S.phase1_computation(sp, acc0, acc1, False, False, S.CHECK_NO, *k0)
    resp = bytearray(sp.get_length(True))
    S.phase1_computation(sp, bytes(acc1), resp, False, True, S.CHECK_NO, *k1)
    assert bytes(resp) == O.phase1_computation(cid, bytes(acc1), sp.get_length(True), False, True, 3, rp.g1_chunk_size,
                                               rp.other_chunk_size, 0, *k1)
    newc = bytearray(sp.get_length(False))
    seed = bytes(range(32))
    pairs = S.phase1_verification_vectors(sp, bytes(resp), True, newc, False, seed=seed)
    # new challenge == decompressed response (verification.rs:271-274 + beta_g2 :199-201)
    want = O.phase1_computation(cid, bytes(acc1), sp.get_length(False), False, False, 3, rp.g1_chunk_size,
                                rp.other_chunk_size, 0, *k1)
    assert bytes(newc[64:]) == want[64:]
    # accumulated tau after two contributions
    tau = k0[0] * k1[0] % cv.r
    for (s, sx), g in zip(pairs, (cv.g1, cv.g2, cv.g1, cv.g1)):
        Sp, SXp = g.decode(s, False), g.decode(sx, False)
        assert Sp is not None and g.mul(Sp, tau) == SXp
    # tamper one alpha_g1 element of the response: that vector's ratio must fail, others still pass
    offs = rp.split_offsets(True)
    o, c, sz = offs[2]
    bad = bytearray(resp)
    other = cv.g1.encode(cv.g1.mul(cv.g1.gen, 4242), True)
    bad[o + 5 * sz:o + 6 * sz] = other
    pairs = S.phase1_verification_vectors(sp, bytes(bad), True, None, False, seed=seed)
    verdicts = [g.mul(g.decode(s, False), tau) == g.decode(sx, False) for (s, sx), g in zip(pairs, (cv.g1, cv.g2, cv.g1, cv.g1))]
    assert verdicts == [True, True, False, True]
    # an element outside the subgroup in tau_g2 -> IncorrectSubgroup, infinity -> PointAtInfinity
    o, c, sz = offs[1]
    bad = bytearray(resp)
    bad[o + 3 * sz:o + 4 * sz] = cv.g2.encode(None, True)
    with pytest.raises(S.PointAtInfinity):
        S.phase1_verification_vectors(sp, bytes(bad), True, None, False, seed=seed)


def test_aggregation_split_decompress():
    """phase1/src/aggregation.rs:355-846 shape: full accumulator -> split into chunks -> aggregate back, in
    every compression combination, plus decompress (helpers/accumulator.rs:352-388 round trip)."""
    cv, cid = R.BLS12_377, S.BLS12_377
    power, batch, csz = 4, 8, 8
    rng = random.Random(321)
    rp = R.Phase1Parameters(cv, power, batch)
    sp = S.Phase1Parameters(cid, power, batch)
    keys = [rng.randrange(2, cv.r) for _ in range(3)]
    full_u = O.phase1_computation(0, bytes(R.phase1_initialization(rp, False)), rp.get_length(False), False, False, 3,
                                  rp.g1_chunk_size, rp.other_chunk_size, 0, *keys)
    full_c = bytearray(rp.get_length(True))  # the same accumulator, compressed (oracle transcode per vector)
    for vec, ((o, c, sz), (co, cc, csz_)) in enumerate(zip(rp.split_offsets(False), rp.split_offsets(True))):
        full_c[co:co + cc * csz_] = O.transcode(0, 1 if vec in (1, 4) else 0, full_u[o:o + c * sz], False, 3, True, c)
    full_c = bytes(full_c)
    assert S.phase1_decompress(sp, full_c)[64:] == full_u[64:]
    nchunks = (rp.powers_g1_length + csz - 1) // csz
    for comp_full, full in ((False, full_u), (True, full_c)):
        offs_full = rp.split_offsets(comp_full)
        for comp_chunk in (False, True):
            rebuilt = bytearray(len(full))
            for ci in range(nchunks):
                cp_r = R.Phase1Parameters(cv, power, batch, R.CHUNKED_MODE, ci, csz)
                cp_s = S.Phase1Parameters(cid, power, batch, 1, ci, csz)
                chunk = S.phase1_split_chunk(cp_s, full, comp_full, comp_chunk)
                # expected chunk: slices of the full vectors, transcoded by the oracle
                for vec, ((o, c, sz), (fo, fc, fsz)) in enumerate(zip(cp_r.split_offsets(comp_chunk), offs_full)):
                    grp = 1 if vec in (1, 4) else 0
                    first = 0 if vec == 4 else ci * csz
                    src = full[fo + first * fsz:fo + (first + c) * fsz]
                    want = O.transcode(0, grp, src, comp_full, 3, comp_chunk, c) if c else b""
                    assert chunk[o:o + c * sz] == want, (comp_full, comp_chunk, ci, vec)
                S.phase1_aggregate_chunk(cp_s, chunk, comp_chunk, rebuilt, comp_full)
            assert bytes(rebuilt[64:]) == full[64:], (comp_full, comp_chunk)
