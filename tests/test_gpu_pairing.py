"""GPU tests of same_ratio / check_same_ratio (setup-utils/src/helpers.rs:406-424) on the device against the
oracle's textbook Tate pairing (oracle/pyref.py) and against the structural cases of the reference's own tests
(helpers.rs:334-368: test_same_ratio / test_power_pairs — accept a consistent ratio, reject a perturbed one)."""
import hashlib
import random

import pytest

import pyref as R
import snark_setup_b200 as S

pytestmark = pytest.mark.gpu
CID = {"bls12_377": S.BLS12_377, "bw6_761": S.BW6_761}


def _pairs(cv, a0, a1, b0, b1):
    return cv.g1.write_batch([a0, a1], False), cv.g2.write_batch([b0, b1], False)


@pytest.mark.parametrize("curve", ["bls12_377", "bw6_761"])
def test_same_ratio_matches_oracle(curve):
    cv, cid = R.CURVES[curve], CID[curve]
    g1, g2 = cv.g1, cv.g2
    rng = random.Random(41)
    s, t, x = (rng.randrange(1, cv.r) for _ in range(3))
    a0 = g1.mul(g1.gen, s)
    b0 = g2.mul(g2.gen, t)
    good = (a0, g1.mul(a0, x), b0, g2.mul(b0, x))
    bad = (a0, g1.mul(a0, x), b0, g2.mul(b0, (x + 1) % cv.r))
    swapped = (g1.mul(a0, x), a0, b0, g2.mul(b0, x))
    for case in (good, bad, swapped):
        want = R.same_ratio(cv, case[:2], case[2:])
        got = S.same_ratio(cid, *_pairs(cv, *case))
        assert got == want
    assert S.same_ratio(cid, *_pairs(cv, *good)) is True
    assert S.same_ratio(cid, *_pairs(cv, *bad)) is False
    S.check_same_ratio(cid, *_pairs(cv, *good))
    with pytest.raises(S.InvalidRatio):
        S.check_same_ratio(cid, *_pairs(cv, *bad))


@pytest.mark.parametrize("curve", ["bls12_377", "bw6_761"])
def test_same_ratio_many_random_ratios(curve):
    """Self-consistency over many scalars (no oracle needed: the ratio is known by construction)."""
    cv, cid = R.CURVES[curve], CID[curve]
    g1, g2 = cv.g1, cv.g2
    rng = random.Random(43)
    n = 6
    p1, p2, expect = b"", b"", []
    for i in range(n):
        s, t, x = (rng.randrange(1, cv.r) for _ in range(3))
        y = x if i % 2 == 0 else rng.randrange(1, cv.r)
        a0, b0 = g1.mul(g1.gen, s), g2.mul(g2.gen, t)
        q1, q2 = _pairs(cv, a0, g1.mul(a0, x), b0, g2.mul(b0, y))
        p1, p2 = p1 + q1, p2 + q2
        expect.append(i % 2 == 0)
    us1, us2 = 2 * g1.size(False), 2 * g2.size(False)
    for i in range(n):
        assert S.same_ratio(cid, p1[i * us1:(i + 1) * us1], p2[i * us2:(i + 1) * us2]) == expect[i]
    with pytest.raises(S.InvalidRatio) as e:
        S.check_same_ratio_batch(cid, p1, p2)
    assert e.value.index == 1
    S.check_same_ratio_batch(cid, p1[:us1] + p1[2 * us1:3 * us1], p2[:us2] + p2[2 * us2:3 * us2])


def test_check_same_ratio_rejects_zero():
    """helpers.rs:415-418: any zero element => InvalidRatio, even when both pairings would be 1."""
    cv, cid = R.BLS12_377, S.BLS12_377
    g1, g2 = cv.g1, cv.g2
    for case in ((None, None, g2.gen, g2.gen), (g1.gen, g1.gen, None, g2.gen), (g1.gen, None, g2.gen, g2.gen)):
        with pytest.raises(S.InvalidRatio):
            S.check_same_ratio(cid, *_pairs(cv, *case))
    # same_ratio itself follows e(O, .) = 1 (helpers.rs:406-408 has no zero check)
    assert S.same_ratio(cid, *_pairs(cv, None, None, g2.gen, g2.gen)) is True
    assert S.same_ratio(cid, *_pairs(cv, g1.gen, None, g2.gen, g2.gen)) is False


@pytest.mark.parametrize("curve,power", [("bls12_377", 10), ("bw6_761", 6)])
def test_power_pairs_feed_check_same_ratio(curve, power):
    """The verification flow end to end on the device (accumulator.rs:56-91): power_pairs of every vector of a
    contributed accumulator against (g2, tau*g2) resp. (g1, tau*g1) — and a tampered vector is rejected."""
    cv, cid = R.CURVES[curve], CID[curve]
    rp = R.Phase1Parameters(cv, power, 256)
    sp = S.Phase1Parameters(cid, power, 256)
    keys = [int.from_bytes(hashlib.blake2b(b"pair" + bytes([i]), digest_size=64).digest(), "little") % (cv.r - 2) + 2
            for i in range(3)]
    resp = bytearray(sp.get_length(True))
    S.phase1_computation(sp, bytes(R.phase1_initialization(rp, False)), resp, False, True, S.CHECK_NO, *keys)
    newc = bytearray(sp.get_length(False))
    seed = hashlib.blake2b(b"rho", digest_size=32).digest()
    pairs = S.phase1_verification_vectors(sp, bytes(resp), True, newc, False, seed=seed)
    offs = rp.split_offsets(False)
    s1, s2 = cv.g1.size(False), cv.g2.size(False)
    tau_g1 = bytes(newc[offs[0][0]:offs[0][0] + 2 * s1])   # (g1, tau*g1)
    tau_g2 = bytes(newc[offs[1][0]:offs[1][0] + 2 * s2])   # (g2, tau*g2)
    (t1, t2, al, be) = pairs
    g1_pairs = t1[0] + t1[1] + tau_g1 + al[0] + al[1] + be[0] + be[1]
    g2_pairs = tau_g2 + t2[0] + t2[1] + tau_g2 + tau_g2
    S.check_same_ratio_batch(cid, g1_pairs, g2_pairs)
    # tamper: swap two adjacent alpha_g1 elements of the response -> that vector's ratio check must fail
    bad = bytearray(resp)
    oc = rp.split_offsets(True)[2][0]
    c1 = cv.g1.size(True)
    bad[oc + 3 * c1:oc + 4 * c1], bad[oc + 4 * c1:oc + 5 * c1] = resp[oc + 4 * c1:oc + 5 * c1], resp[oc + 3 * c1:oc + 4 * c1]
    pairs = S.phase1_verification_vectors(sp, bytes(bad), True, newc, False, seed=seed)
    (t1, t2, al, be) = pairs
    g1_pairs = t1[0] + t1[1] + tau_g1 + al[0] + al[1] + be[0] + be[1]
    with pytest.raises(S.InvalidRatio) as e:
        S.check_same_ratio_batch(cid, g1_pairs, g2_pairs)
    assert e.value.index == 2


@pytest.mark.parametrize("curve,power", [("bls12_377", 12), ("bw6_761", 7)])
def test_phase1_verification_ratios_verdicts(curve, power):
    """verification.rs tests (:783-1106) accept a correct contribution and reject tampered ones; the verdict and
    the failing vector must match what the reference reports (PointAtInfinity / IncorrectSubgroup / InvalidRatio)."""
    cv, cid = R.CURVES[curve], CID[curve]
    rp = R.Phase1Parameters(cv, power, 256)
    sp = S.Phase1Parameters(cid, power, 256)
    keys = [int.from_bytes(hashlib.blake2b(b"verd" + bytes([i]), digest_size=64).digest(), "little") % (cv.r - 2) + 2
            for i in range(3)]
    chal = bytearray(sp.get_length(False))
    S.phase1_computation(sp, bytes(R.phase1_initialization(rp, False)), chal, False, False, S.CHECK_NO, *keys)
    resp = bytearray(sp.get_length(True))
    S.phase1_computation(sp, bytes(chal), resp, False, True, S.CHECK_NO, *[k + 1 for k in keys])
    resp = bytes(resp)
    newc = bytearray(sp.get_length(False))
    seed = bytes(range(32))
    S.phase1_verification_ratios(sp, resp, True, newc, False, seed=seed)
    # the new challenge is the decompressed response
    want = bytearray(sp.get_length(False))
    S.phase1_computation(sp, bytes(chal), want, False, False, S.CHECK_NO, *[k + 1 for k in keys])
    assert bytes(newc[64:]) == bytes(want[64:])
    offs = rp.split_offsets(True)
    c1, c2 = cv.g1.size(True), cv.g2.size(True)
    # swap two adjacent elements in each vector in turn -> InvalidRatio on that vector
    for vec, csz in ((0, c1), (1, c2), (2, c1), (3, c1)):
        o = offs[vec][0]
        bad = bytearray(resp)
        bad[o + 5 * csz:o + 6 * csz], bad[o + 6 * csz:o + 7 * csz] = resp[o + 6 * csz:o + 7 * csz], resp[o + 5 * csz:o + 6 * csz]
        with pytest.raises(S.InvalidRatio) as e:
            S.phase1_verification_ratios(sp, bytes(bad), True, None, False, seed=seed)
        assert e.value.index == vec
    # one element replaced by the wrong power (tau_g1[9] <- tau_g1[8]) -> InvalidRatio on tau_g1
    o = offs[0][0]
    bad = bytearray(resp)
    bad[o + 9 * c1:o + 10 * c1] = resp[o + 8 * c1:o + 9 * c1]
    with pytest.raises(S.InvalidRatio) as e:
        S.phase1_verification_ratios(sp, bytes(bad), True, None, False, seed=seed)
    assert e.value.index == 0
    # infinity in beta_g1 -> PointAtInfinity before any pairing
    o = offs[3][0]
    bad = bytearray(resp)
    bad[o + 4 * c1:o + 5 * c1] = cv.g1.encode(None, True)
    with pytest.raises(S.PointAtInfinity):
        S.phase1_verification_ratios(sp, bytes(bad), True, None, False, seed=seed)


def test_phase2_contribute_then_verify_with_verdict():
    """BASELINE config C5 in miniature (phase2/src/parameters.rs:286-307,393-407): contribute = delta^-1 batch_mul of
    the H and L queries, verify = merge_pairs(before, after) fed to check_same_ratio against
    (delta_after * G2, delta_before * G2) — all on the device; a tampered query element flips the verdict."""
    cv, cid = R.BLS12_377, S.BLS12_377
    g1, g2 = cv.g1, cv.g2
    n = 1 << 12
    rng = random.Random(2024)
    tau, delta = rng.randrange(2, cv.r), rng.randrange(2, cv.r)
    gens = g1.encode(g1.gen, False) * n
    seed = hashlib.blake2b(b"phase2", digest_size=32).digest()
    g2_pair = g2.write_batch([g2.mul(g2.gen, delta), g2.gen], False)   # (after.delta_g2, before.delta_g2)
    for label, first in ((b"h", 1), (b"l", 7)):
        before = S.apply_powers(cid, S.G1, gens, False, S.CHECK_NO, False, n, tau=tau, first_power=first)
        after = bytearray(before)
        S.batch_mul(cid, S.G1, after, pow(delta, -1, cv.r))
        s, sx = S.merge_pairs(cid, S.G1, before, bytes(after), False, seed=seed)
        S.check_same_ratio(cid, s + sx, g2_pair)
        bad = bytearray(after)
        bad[96 * 100:96 * 101] = after[96 * 101:96 * 102]
        s, sx = S.merge_pairs(cid, S.G1, before, bytes(bad), False, seed=seed)
        with pytest.raises(S.InvalidRatio):
            S.check_same_ratio(cid, s + sx, g2_pair)
