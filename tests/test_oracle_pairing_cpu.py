"""CPU suite for the §8(f) rank-2 and rank-4 oracles: the textbook Tate pairing of oracle/pyref.py must be bilinear
and non-degenerate on G1 x G2 (which is all the verdict of same_ratio depends on — setup-utils/src/helpers.rs:334-368
tests the same two properties through arkworks), and the QAP dot products must satisfy their defining identity."""
import random

import pytest

import pyref as R

CURVES = [R.BLS12_377, R.BW6_761]


@pytest.mark.parametrize("cv", CURVES, ids=["bls12_377", "bw6_761"])
def test_untwist_lands_on_the_curve(cv):
    K, k, tw = R.pairing_tower(cv)
    F = K.F
    x, y = R.untwist(cv, cv.g2.mul(cv.g2.gen, 12345))
    b = [F.zero] * 6
    b[0] = (cv.g1.b, 0) if F.degree == 2 else cv.g1.b
    assert K.mul(y, y) == [F.add(u, v) for u, v in zip(K.mul(K.mul(x, x), x), b)]
    assert (k, tw) == ((12, "D") if cv.name == "bls12_377" else (6, "M"))
    q = cv.g1.F.p
    assert (q ** k - 1) % cv.r == 0 and all((q ** j - 1) % cv.r for j in range(1, k))  # embedding degree


def test_pairing_bilinear_nondegenerate_bls12_377():
    cv = R.BLS12_377
    K, _, _ = R.pairing_tower(cv)
    e = R.pairing(cv, cv.g1.gen, cv.g2.gen)
    assert e != K.one() and K.pow(e, cv.r) == K.one()
    assert R.pairing(cv, cv.g1.mul(cv.g1.gen, 3), cv.g2.mul(cv.g2.gen, 5)) == K.pow(e, 15)


@pytest.mark.parametrize("cv", CURVES, ids=["bls12_377", "bw6_761"])
def test_same_ratio_accepts_and_rejects(cv):
    """helpers.rs:334-349 test_same_ratio: (g1, s*g1) vs (g2, s*g2) passes, a different scalar fails."""
    rng = random.Random(2)
    s = rng.randrange(1, cv.r)
    g1p = (cv.g1.gen, cv.g1.mul(cv.g1.gen, s))
    assert R.same_ratio(cv, g1p, (cv.g2.gen, cv.g2.mul(cv.g2.gen, s)))
    assert not R.same_ratio(cv, g1p, (cv.g2.gen, cv.g2.mul(cv.g2.gen, s + 1)))
    with pytest.raises(R.InvalidRatio):
        R.check_same_ratio(cv, (None, None), (cv.g2.gen, cv.g2.gen))


def test_power_pairs_ratio_bls12_377():
    """helpers.rs:351-368 test_power_pairs: power_pairs of a powers vector has ratio tau; a corrupted vector has not."""
    cv = R.BLS12_377
    rng = random.Random(3)
    tau = rng.randrange(1, cv.r)
    v = [cv.g1.mul(cv.g1.gen, pow(tau, i, cv.r)) for i in range(6)]
    rho = [rng.randrange(cv.r) for _ in range(5)]
    g2p = (cv.g2.gen, cv.g2.mul(cv.g2.gen, tau))
    assert R.same_ratio(cv, R.power_pairs(cv.g1, v, rho), g2p)
    v[3] = cv.g1.mul(cv.g1.gen, 7)
    assert not R.same_ratio(cv, R.power_pairs(cv.g1, v, rho), g2p)


def test_qap_dot_products():
    cv, g = R.BLS12_377, R.BLS12_377.g1
    rng = random.Random(4)
    s = [rng.randrange(cv.r) for _ in range(5)]
    bases = [g.mul(g.gen, x) for x in s]
    xt = [[(rng.randrange(cv.r), rng.randrange(4)) for _ in range(2)] for _ in range(5)]  # 5 constraints, 4 variables
    rows = R.process_matrix(xt, 4)
    assert sum(len(r) for r in rows) == 10
    got = R.dot_product_vec(g, rows, bases)
    for row, P in zip(rows, got):
        assert P == g.mul(g.gen, sum(c * s[i] for c, i in row) % cv.r)
