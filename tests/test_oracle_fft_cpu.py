"""CPU suite for prepare_phase2: pins the oracle's three formulations of `to_coeffs`
(definition of the inverse DFT, recursive decimation in frequency, scalar-field IFFT of known discrete logs
pushed through the C++ oracle's scalar multiplication) to one another and to tests/golden/
prepare_phase2_vectors.json, and checks the evaluation-domain constants the device library was generated with."""
import hashlib
import json
import os
import random
import re

import pytest

import coracle as O
import pyref as R

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "prepare_phase2_vectors.json")))
HOT = json.load(open(os.path.join(HERE, "golden", "hotpath_vectors.json")))
CURVES = [(0, R.BLS12_377), (1, R.BW6_761)]


def _small_factors(n, limit=3000):
    f, p = [], 2
    while p < limit:
        if n % p == 0:
            f.append(p)
            while n % p == 0:
                n //= p
        p += 1
    return f


@pytest.mark.parametrize("cid,cv", CURVES, ids=["bls12_377", "bw6_761"])
def test_scalar_field_generator_and_roots(cid, cv):
    r = cv.r
    g = R.FR_GENERATOR[r]
    fac = _small_factors(r - 1)
    assert 2 in fac
    # arkworks picks the smallest multiplicative generator; no smaller candidate survives the small factors
    def survives(x):
        return all(pow(x, (r - 1) // l, r) != 1 for l in fac)
    assert survives(g) and not any(survives(x) for x in range(2, g))
    s = R.two_adicity(r)
    assert s == (47 if cid == 0 else 46)
    w = R.get_root_of_unity(r, 1 << s)
    assert w == pow(g, (r - 1) >> s, r) and pow(w, 1 << (s - 1), r) == r - 1
    for n in (1, 2, 8, 1 << 20):
        wn = R.get_root_of_unity(r, n)
        assert pow(wn, n, r) == 1 and (n == 1 or pow(wn, n // 2, r) == r - 1)
    with pytest.raises(ValueError):
        R.get_root_of_unity(r, 3)
    assert [R.domain_size(x) for x in (0, 1, 2, 3, 4, 5, 1000)] == [1, 1, 2, 4, 4, 8, 1024]
    assert GOLD["curves"][cv.name]["two_adic_root_of_unity"] == hex(w)


def test_generator_matches_recalled_montgomery_limbs():
    """ark-bls12-377 Fr: GENERATOR = 22, whose Montgomery form (R = 2^256) was recorded in the zexe sources as
    these four 64-bit limbs — an independent memory of the same constant."""
    r = R.BLS12_377_R
    limbs = [2984901390528151251, 10561528701063790279, 5476750214495080041, 898978044469942640]
    assert (22 << 256) % r == sum(l << (64 * i) for i, l in enumerate(limbs))


def test_device_constants_carry_the_same_root():
    src = open(os.path.join(ROOT, "snark-setup_b200", "csrc", "constants_gen.cuh")).read()
    for name, r, nw in (("Bls377Fr", R.BLS12_377_R, 8), ("Bls377Fq", R.BLS12_377_Q, 12)):
        body = src[src.index("struct %s {" % name):]
        m = re.search(r"fft_root\(int i\) \{ const uint32_t v\[%d\] = \{([^}]*)\}" % nw, body)
        limbs = [int(x.strip().rstrip("u"), 16) for x in m.group(1).split(",")]
        mont = sum(l << (32 * i) for i, l in enumerate(limbs))
        w = R.get_root_of_unity(r, 1 << R.two_adicity(r))
        assert mont == (w << (32 * nw)) % r


@pytest.mark.parametrize("cid,cv", CURVES, ids=["bls12_377", "bw6_761"])
def test_golden_groth16_params(cid, cv):
    v = HOT["phase1"][cv.name]
    p = R.Phase1Parameters(cv, v["power"], v["batch_size"])
    acc = bytes.fromhex(v["challenge1"])
    for size, ent in GOLD["curves"][cv.name]["params"].items():
        got = R.groth16_params_new(p, acc, False, int(size), True)  # fast algorithm vs the definition
        assert got.hex() == ent["compressed"]
        unc = R.groth16_params_new(p, acc, False, int(size), False)
        assert hashlib.blake2b(unc).hexdigest() == ent["uncompressed_blake2b"]
        m = R.domain_size(int(size))
        s1c, s2c = cv.g1.size(True), cv.g2.size(True)
        assert len(got) == 2 * s1c + s2c + 3 * m * s1c + m * s2c + (m - 1) * s1c
    with pytest.raises(R.InvalidLength):
        R.groth16_params_new(p, acc, False, 5, True)


@pytest.mark.parametrize("cid,cv", CURVES, ids=["bls12_377", "bw6_761"])
def test_known_discrete_logs_route(cid, cv):
    """P_i = s_i G => to_coeffs(P) = scalar_ifft(s) G, with the C++ oracle doing the scalar multiplications."""
    n = 16 if cid == 0 else 8
    rng = random.Random(7 + cid)
    s = [rng.randrange(cv.r) for _ in range(n)]
    s[2] = 0
    for gid, g in ((0, cv.g1), (1, cv.g2)):
        gens = g.encode(g.gen, False) * n
        pts = g.read_batch(O.apply_powers(cid, gid, gens, False, R.NO, False, n, powers=s), False)
        want = O.apply_powers(cid, gid, gens, False, R.NO, True, n, powers=R.scalar_ifft(cv.r, s))
        assert g.write_batch(R.group_ifft_fast(g, pts), True) == want
    # forward evaluation identity used by the GPU suite at config sizes
    c = R.scalar_ifft(cv.r, s)
    w = R.get_root_of_unity(cv.r, n)
    for k in (0, 1, n - 1):
        assert sum(c[j] * pow(w, j * k, cv.r) for j in range(n)) % cv.r == s[k]


@pytest.mark.parametrize("cid,cv", CURVES, ids=["bls12_377", "bw6_761"])
def test_cpp_oracle_group_ifft(cid, cv):
    """oracle.cpp::group_ifft (iterative decimation in frequency, double-and-add twiddles — the reference algorithm's
    cost profile) against the pyref algorithms and the golden Groth16Params coefficients."""
    rng = random.Random(21 + cid)
    for gid, g in ((0, cv.g1), (1, cv.g2)):
        for n in (1, 2, 16 if cid == 0 else 4):
            pts = [g.mul(g.gen, rng.randrange(1, cv.r)) for _ in range(n)]
            if n > 2:
                pts[1] = None
            want = g.write_batch(R.group_ifft_fast(g, pts), True)
            assert O.group_ifft(cid, gid, g.write_batch(pts, False), False, True) == want
            assert O.group_ifft(cid, gid, g.write_batch(pts, True), True, True) == want
    v = HOT["phase1"][cv.name]
    p = R.Phase1Parameters(cv, v["power"], v["batch_size"])
    acc = bytes.fromhex(v["challenge1"])
    gold = bytes.fromhex(GOLD["curves"][cv.name]["params"]["4"]["compressed"])
    s1c, s2c = cv.g1.size(True), cv.g2.size(True)
    (o1, _, z1), (o2, _, z2) = p.split_offsets(False)[:2]
    assert O.group_ifft(cid, 0, acc[o1:o1 + 4 * z1], False, True) == gold[2 * s1c + s2c:2 * s1c + s2c + 4 * s1c]
    assert O.group_ifft(cid, 1, acc[o2:o2 + 4 * z2], False, True) == gold[2 * s1c + s2c + 4 * s1c:2 * s1c + s2c + 4 * s1c + 4 * s2c]
    with pytest.raises(O.OracleError):
        O.group_ifft(cid, 0, cv.g1.encode(cv.g1.gen, False) * 3, False, True)
