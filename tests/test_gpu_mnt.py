"""MNT4-753 / MNT6-753 through the C ABI (the reference exposes both curves: setup-utils/src/converters.rs:18-45,
phase1-cli/src/bin/phase1.rs:146-151): a != 0 curves, G2 over Fq2 / Fq3, 95-byte field elements and scalars.  Functional
parity with the big-integer oracle for every hot-path entry point: read_batch / write_batch, batch_exp / batch_mul /
apply_powers, check_subgroup, merge_pairs / power_pairs, Phase1::computation and the verification vectors."""
import random

import pytest

import pyref as R
import snark_setup_b200 as S

pytestmark = pytest.mark.gpu

CURVES = [("mnt4_753", S.MNT4_753), ("mnt6_753", S.MNT6_753)]


def _pts(g, cv, rng, n):
    base = g.mul(g.gen, rng.randrange(1, cv.r))
    out = [base]
    for _ in range(n - 1):
        out.append(g.add(out[-1], base))  # cheap distinct subgroup points
    return out


@pytest.mark.parametrize("cname,cid", CURVES)
def test_sizes(cname, cid):
    cv = R.curve_by_name(cname)
    assert [S.element_size(cid, g, c) for g in (0, 1) for c in (0, 1)] == [cv.g1.size(False), cv.g1.size(True),
                                                                           cv.g2.size(False), cv.g2.size(True)]
    assert S.scalar_size(cid) == 95
    for k, bs, mode, ci, cs in ((3, 4, 0, 0, 0), (4, 8, 1, 1, 8), (4, 8, 1, 3, 8)):
        a, b = S.Phase1Parameters(cid, k, bs, mode, ci, cs), R.Phase1Parameters(cv, k, bs, mode, ci, cs)
        assert (a.accumulator_size, a.contribution_size, a.public_key_size) == (b.accumulator_size, b.contribution_size, b.public_key_size)


@pytest.mark.parametrize("cname,cid", CURVES)
@pytest.mark.parametrize("gname,gid", [("g1", 0), ("g2", 1)])
def test_codec_and_batch_exp(cname, cid, gname, gid):
    cv = R.curve_by_name(cname)
    g = getattr(cv, gname)
    rng = random.Random(cid * 2 + gid)
    n = 9
    pts = _pts(g, cv, rng, n)
    pts[4] = None  # the point at infinity travels through CheckForCorrectness::No
    for cin in (False, True):
        for cout in (False, True):
            assert S.transcode(cid, gid, g.write_batch(pts, cin), cin, S.CHECK_NO, cout) == g.write_batch(pts, cout)
    with pytest.raises(S.PointAtInfinity) as ei:
        S.transcode(cid, gid, g.write_batch(pts, True), True, S.CHECK_FULL, False)
    assert ei.value.index == 4
    pts[4] = g.mul(g.gen, 5)
    # apply_powers: tau^(first + i) (* coeff), compressed out — and explicit 95-byte scalars (batch_exp)
    tau, coeff = rng.randrange(2, cv.r), rng.randrange(2, cv.r)
    got = S.apply_powers(cid, gid, g.write_batch(pts, False), False, S.CHECK_NO, True, n, tau=tau, first_power=3, coeff=coeff)
    want = g.write_batch(R.batch_exp(g, pts, R.generate_powers_of_tau(cv, tau, 3, 3 + n), coeff), True)
    assert got == want
    exps = [0, 1, cv.r - 1] + [rng.randrange(cv.r) for _ in range(n - 3)]
    buf = bytearray(g.write_batch(pts, False))
    S.batch_exp(cid, gid, buf, exps, coeff=None)
    assert bytes(buf) == g.write_batch(R.batch_exp(g, pts, exps), False)
    buf = bytearray(g.write_batch(pts[:4], False))
    S.batch_mul(cid, gid, buf, coeff)
    assert bytes(buf) == g.write_batch(R.batch_mul(g, pts[:4], coeff), False)
    assert S.generate_powers_of_tau(cid, tau, 5, 12) == R.generate_powers_of_tau(cv, tau, 5, 12)


@pytest.mark.parametrize("cname,cid", CURVES)
def test_subgroup_and_ratio(cname, cid):
    cv = R.curve_by_name(cname)
    rng = random.Random(17 + cid)
    tau = rng.randrange(2, cv.r)
    for g, gid in ((cv.g1, 0), (cv.g2, 1)):
        n = 12
        base = g.mul(g.gen, rng.randrange(1, cv.r))
        v = [g.mul(base, pow(tau, i, cv.r)) for i in range(n)]
        buf = g.write_batch(v, True)
        S.check_subgroup(cid, gid, buf, True)
        # generated rho: the pair satisfies the ratio
        s, sx = S.power_pairs(cid, gid, buf, True, seed=bytes(range(32)))
        assert g.mul(g.decode(s, False), tau) == g.decode(sx, False)
        # explicit full-width rho: the exact sums of the oracle
        rho = [rng.randrange(cv.r) for _ in range(n - 1)]
        s, sx = S.power_pairs(cid, gid, buf, True, rho=rho)
        ws, wsx = R.power_pairs(g, v, rho)
        assert (s, sx) == (g.encode(ws, False), g.encode(wsx, False))
        bad = list(v)
        bad[5] = g.mul(base, 424242)
        s, sx = S.power_pairs(cid, gid, g.write_batch(bad, True), True, seed=bytes(range(32)))
        assert g.mul(g.decode(s, False), tau) != g.decode(sx, False)
    # G2: a twist point outside the order-r subgroup
    g = cv.g2
    c = 1
    while True:
        x = (c, 1) if g.F.degree == 2 else (c, 1, 0)
        y = g.F.sqrt(g.rhs(x))
        if y is not None and g.mul((x, y), cv.r) is not None:
            break
        c += 1
    pts = [g.mul(g.gen, 3), (x, y), g.mul(g.gen, 4)]
    with pytest.raises(S.IncorrectSubgroup) as ei:
        S.check_subgroup(cid, 1, g.write_batch(pts, False), False)
    assert ei.value.index == 1
    with pytest.raises(S.InvalidData):
        S.transcode(cid, 1, g.write_batch(pts, True), True, S.CHECK_FULL, False)


@pytest.mark.parametrize("cname,cid", CURVES)
def test_phase1_round(cname, cid):
    """Phase1::computation + the verification vectors on a 2^2-power accumulator, full and as two shards."""
    cv = R.curve_by_name(cname)
    rng = random.Random(99 + cid)
    power, batch = 2, 4
    rp, sp = R.Phase1Parameters(cv, power, batch), S.Phase1Parameters(cid, power, batch)
    k0 = [rng.randrange(2, cv.r) for _ in range(3)]
    k1 = [rng.randrange(2, cv.r) for _ in range(3)]
    blank = bytes(R.phase1_initialization(rp, False))  # the G2 slots hold pyref's derived generator (test input only)
    chal = bytes(R.phase1_computation(rp, blank, False, False, R.NO, *k0))
    want = bytes(R.phase1_computation(rp, chal, False, True, R.NO, *k1))
    resp = bytearray(sp.get_length(True))
    S.phase1_computation(sp, chal, resp, False, True, S.CHECK_NO, *k1)
    assert bytes(resp[64:]) == want[64:]
    sharded = bytearray(sp.get_length(True))
    for r in range(2):
        S.phase1_computation(sp, chal, sharded, False, True, S.CHECK_NO, *k1, shard=(r, 2))
    assert sharded == resp
    newc = bytearray(sp.get_length(False))
    pairs = S.phase1_verification_vectors(sp, bytes(resp), True, newc, False, seed=bytes(32))
    assert bytes(newc[64:]) == bytes(R.phase1_computation(rp, chal, False, False, R.NO, *k1))[64:]
    tau = k0[0] * k1[0] % cv.r
    for (s, sx), g in zip(pairs, (cv.g1, cv.g2, cv.g1, cv.g1)):
        assert g.mul(g.decode(s, False), tau) == g.decode(sx, False)


@pytest.mark.parametrize("cname,cid", CURVES)
def test_what_is_not_available_fails_loudly(cname, cid):
    sp = S.Phase1Parameters(cid, 2, 4)
    with pytest.raises(S.InvalidArgument):
        S.phase1_initialization(sp, False)  # arkworks' G2 generator constant is not known to this build
    cv = R.curve_by_name(cname)
    g1p = cv.g1.write_batch([cv.g1.gen, cv.g1.gen], False)
    g2p = cv.g2.write_batch([cv.g2.gen, cv.g2.gen], False)
    with pytest.raises(S.SetupError):
        S.same_ratio(cid, g1p, g2p)  # no device pairing for the MNT curves: the (s, sx) pairs go to the host's pairing
