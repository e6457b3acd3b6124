"""GPU parity tests of prepare_phase2: the group IFFT `to_coeffs`, the H query and
Groth16Params::new + ::write (setup-utils/src/groth16_utils.rs:44-168), through the C ABI, against
oracle/pyref.py (definition of the inverse DFT and an independent recursive algorithm), against the C++
oracle through points with known discrete logs, and — at sizes no oracle reaches — through the
forward-evaluation identity sum_j w^(j*k) * coeffs_j = P_k checked with the bucket MSM.  The structure follows
the reference's own tests (groth16_utils.rs:253-365: first_half_powers, phase2_equal_to_powers,
large_phase2_fails)."""
import hashlib
import os
import random

import pytest

import coracle as O
import pyref as R
import snark_setup_b200 as S

pytestmark = pytest.mark.gpu
FULL = os.environ.get("SS_TEST_FULL") == "1"
CID = {"bls12_377": S.BLS12_377, "bw6_761": S.BW6_761}


def _group(cv, grp):
    return cv.g2 if grp else cv.g1


def _random_points(g, n, rng):
    return [g.mul(g.gen, rng.randrange(1, g.r)) for _ in range(n)]


@pytest.mark.parametrize("curve,grp,n", [("bls12_377", 0, 8), ("bls12_377", 1, 4), ("bw6_761", 0, 4), ("bw6_761", 1, 4)])
def test_group_ifft_matches_definition(curve, grp, n):
    cv = R.CURVES[curve]
    g = _group(cv, grp)
    pts = _random_points(g, n, random.Random(11 + n))
    want = R.group_ifft(g, pts)
    for cin, cout in ((False, True), (True, False)):
        got = S.group_ifft(CID[curve], grp, g.write_batch(pts, cin), cin, cout)
        assert got == g.write_batch(want, cout)


@pytest.mark.parametrize("curve,grp", [("bls12_377", 0), ("bls12_377", 1), ("bw6_761", 0)])
@pytest.mark.parametrize("n", [1, 2, 16])
def test_group_ifft_sizes(curve, grp, n):
    if curve == "bw6_761" and n > 2:
        n = 8
    cv = R.CURVES[curve]
    g = _group(cv, grp)
    pts = _random_points(g, n, random.Random(100 + n))
    got = S.group_ifft(CID[curve], grp, g.write_batch(pts, False), False, False)
    assert got == g.write_batch(R.group_ifft_fast(g, pts), False)


def test_group_ifft_exceptional_inputs():
    """Identity elements and equal / opposite points drive every butterfly into the doubling and
    cancellation branches of the mixed addition."""
    cv = R.BLS12_377
    g = cv.g1
    P = g.mul(g.gen, 12345)
    cases = [
        [P] * 8,                                            # lo + t doubles, lo - t cancels
        [P, g.neg(P)] * 4,
        [None] * 8,                                         # all identity
        [P, None, None, g.neg(P), None, P, P, None],
        [None, g.gen, None, g.gen, None, g.gen, None, g.gen],
    ]
    for pts in cases:
        got = S.group_ifft(S.BLS12_377, S.G1, g.write_batch(pts, False), False, True)
        assert got == g.write_batch(R.group_ifft_fast(g, pts), True), pts


def test_group_ifft_rejects_bad_input():
    g = R.BLS12_377.g1
    pts = _random_points(g, 3, random.Random(1))
    with pytest.raises(S.SetupError):  # 3 is not a radix-2 domain size
        S.group_ifft(S.BLS12_377, S.G1, g.write_batch(pts, False), False, False)
    buf = bytearray(g.write_batch(pts + pts[:1], True))
    buf[48 * 2:48 * 3] = g.encode(None, True)
    with pytest.raises(S.PointAtInfinity) as e:  # read with CheckForCorrectness::Full, like prepare_phase2 can
        S.group_ifft(S.BLS12_377, S.G1, bytes(buf), True, False, check=S.CHECK_FULL)
    assert e.value.index == 2


@pytest.mark.parametrize("curve,grp,log_n", [("bls12_377", 0, 12), ("bls12_377", 1, 10), ("bw6_761", 0, 9), ("bw6_761", 1, 9)])
def test_group_ifft_known_discrete_logs(curve, grp, log_n):
    """P_i = s_i * G  =>  coeffs = scalar_ifft(s) * G; both sides' scalar multiplications by the C++ oracle."""
    cv = R.CURVES[curve]
    g = _group(cv, grp)
    cid = CID[curve]
    n = 1 << log_n
    rng = random.Random(log_n)
    s = [rng.randrange(cv.r) for _ in range(n)]
    s[3] = 0  # an identity element among the inputs
    gens = g.encode(g.gen, False) * n
    pts = O.apply_powers(cid, grp, gens, False, 3, False, n, powers=s)
    want = O.apply_powers(cid, grp, gens, False, 3, True, n, powers=R.scalar_ifft(cv.r, s))
    assert S.group_ifft(cid, grp, pts, False, True) == want


@pytest.mark.parametrize("curve,grp,log_n", [("bls12_377", 0, 11), ("bls12_377", 1, 9), ("bw6_761", 0, 8), ("bw6_761", 1, 8)])
def test_group_ifft_matches_cpp_oracle(curve, grp, log_n):
    """Random subgroup points (no structure), device vs oracle.cpp::group_ifft (iterative decimation in frequency)."""
    cv = R.CURVES[curve]
    g = _group(cv, grp)
    cid = CID[curve]
    n = 1 << log_n
    gens = g.encode(g.gen, False) * n
    pts = O.apply_powers(cid, grp, gens, False, 3, False, n, tau=0x1234567 + log_n, first_power=3, coeff=987654321)
    for cout in (True, False):
        assert S.group_ifft(cid, grp, pts, False, cout) == O.group_ifft(cid, grp, pts, False, cout)


@pytest.mark.parametrize("curve", ["bls12_377", "bw6_761"])
def test_h_query(curve):
    cv = R.CURVES[curve]
    g = cv.g1
    pts = _random_points(g, 15, random.Random(3))
    pts[9] = pts[1]  # h_1 = identity
    for degree in (1, 2, 8):
        got = S.h_query_groth16(CID[curve], g.write_batch(pts, False), False, degree, True)
        assert got == g.write_batch(R.h_query_groth16(g, pts, degree), True)
    with pytest.raises(S.InvalidLength):
        S.h_query_groth16(CID[curve], g.write_batch(pts[:14], False), False, 8, True)


def _accumulator(cv, cid, power, seed):
    rp = R.Phase1Parameters(cv, power, 256)
    sp = S.Phase1Parameters(cid, power, 256)
    keys = [int.from_bytes(hashlib.blake2b(seed + bytes([i]), digest_size=64).digest(), "little") % (cv.r - 2) + 2
            for i in range(3)]
    acc = bytearray(sp.get_length(False))
    S.phase1_computation(sp, bytes(R.phase1_initialization(rp, False)), acc, False, False, S.CHECK_NO, *keys)
    return rp, sp, bytes(acc)


@pytest.mark.parametrize("curve,power", [("bls12_377", 3), ("bls12_377", 4), ("bw6_761", 2)])
def test_groth16_params_new_matches_oracle(curve, power):
    """groth16_utils.rs tests first_half_powers / phase2_equal_to_powers: byte parity of the written parameters."""
    cv = R.CURVES[curve]
    rp, sp, acc = _accumulator(cv, CID[curve], power, b"fft-acc")
    for phase2_size in sorted({1 << power, (1 << power) // 2, max(1, (1 << power) - 1), 1}):
        for cout in (False, True):
            got = S.groth16_params_new(sp, acc, False, phase2_size, cout)
            assert got == R.groth16_params_new(rp, acc, False, phase2_size, cout), (phase2_size, cout)
    # compressed accumulator in, read with full validation
    acc_c = bytearray(sp.get_length(True))
    for (o, c, s), (oc, _, sc), grp in zip(rp.split_offsets(False), rp.split_offsets(True), (0, 1, 0, 0, 1)):
        acc_c[oc:oc + c * sc] = S.transcode(CID[curve], grp, acc[o:o + c * s], False, S.CHECK_NO, True)
    got = S.groth16_params_new(sp, bytes(acc_c), True, 1 << power, False, check=S.CHECK_FULL)
    assert got == R.groth16_params_new(rp, acc, False, 1 << power, False)


def test_groth16_params_new_large_phase2_fails():
    """groth16_utils.rs:354-364 (large_phase2_fails: #[should_panic])."""
    cv = R.BLS12_377
    rp, sp, acc = _accumulator(cv, S.BLS12_377, 3, b"fft-acc")
    for cout in (True, False):
        with pytest.raises(S.InvalidLength):
            S.groth16_params_new(sp, acc, False, 9, cout)
    assert S.groth16_params_size(S.BLS12_377, 9, False) == (16, 2 * 96 + 192 + 3 * 16 * 96 + 16 * 192 + 15 * 96)


@pytest.mark.parametrize("curve,power", [("bls12_377", 18 if FULL else 15), ("bw6_761", 14 if FULL else 11)])
def test_groth16_params_forward_evaluation(curve, power):
    """At config sizes: evaluating the Lagrange coefficients back at w^k must return the k-th power of tau,
    sum_j w^(j*k) coeffs_j = P_k, for every vector (one MSM per probe), and sum_j coeffs_j = P_0."""
    cv = R.CURVES[curve]
    cid = CID[curve]
    rp, sp, acc = _accumulator(cv, cid, power, b"fft-eval")
    m = 1 << power
    out = S.groth16_params_new(sp, acc, False, m, False)
    s1, s2 = cv.g1.size(False), cv.g2.size(False)
    offs = rp.split_offsets(False)
    w = R.get_root_of_unity(cv.r, m)
    pos = 2 * s1 + s2
    for vec, grp in ((0, 0), (1, 1), (2, 0), (3, 0)):
        sz = s2 if grp else s1
        coeffs = out[pos:pos + m * sz]
        pos += m * sz
        o = offs[vec][0]
        for k in (0, 1, m // 2 + 5):
            wk = pow(w, k, cv.r)
            rho, t = [], 1
            for _ in range(m):
                rho.append(t)
                t = t * wk % cv.r
            s_, _ = S.merge_pairs(cid, grp, coeffs, coeffs, False, rho=rho)
            assert s_ == acc[o + k * sz:o + (k + 1) * sz], (vec, k)
    # H query against a few directly computed differences
    g = cv.g1
    tau_g1 = acc[offs[0][0]:offs[0][0] + (2 * m - 1) * s1]
    h = out[pos:pos + (m - 1) * s1]
    assert len(out) == pos + (m - 1) * s1
    for i in (0, 1, m // 3, m - 2):
        hi = g.decode(tau_g1[(i + m) * s1:(i + m + 1) * s1], False)
        lo = g.decode(tau_g1[i * s1:(i + 1) * s1], False)
        assert h[i * s1:(i + 1) * s1] == g.encode(g.add(hi, g.neg(lo)), False)


@pytest.mark.parametrize("curve", ["bls12_377", "bw6_761"])
def test_groth16_params_golden(curve):
    """Committed vectors (tests/golden/prepare_phase2_vectors.json, made from the definition of the inverse DFT)."""
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "prepare_phase2_vectors.json")))["curves"][curve]
    hot = json.load(open(os.path.join(here, "golden", "hotpath_vectors.json")))["phase1"][curve]
    sp = S.Phase1Parameters(CID[curve], hot["power"], hot["batch_size"])
    acc = bytes.fromhex(hot["challenge1"])
    for size, ent in gold["params"].items():
        assert S.groth16_params_new(sp, acc, False, int(size), True).hex() == ent["compressed"]
        unc = S.groth16_params_new(sp, acc, False, int(size), False)
        assert hashlib.blake2b(unc).hexdigest() == ent["uncompressed_blake2b"]
