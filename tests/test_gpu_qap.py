"""GPU parity tests of the phase-2 QAP evaluation primitives (phase2/src/polynomial.rs:11-94)
through the C ABI against oracle/pyref.py, in the shape of the reference's own test (polynomial.rs:96-179: a small
QAP evaluated against known coefficients) plus the edge cases of the CSR boundary."""
import random

import pytest

import coracle as O
import pyref as R
import snark_setup_b200 as S

pytestmark = pytest.mark.gpu
CID = {"bls12_377": S.BLS12_377, "bw6_761": S.BW6_761}


def _rows(rng, r, num_vars, n_bases, density=3):
    rows = []
    for v in range(num_vars):
        k = rng.randrange(0, density + 1)
        row = []
        for _ in range(k):
            c = rng.choice([1, r - 1, 1, r - 1, 2, rng.randrange(r), 0])
            row.append((c, rng.randrange(n_bases)))
        rows.append(row)
    return rows


@pytest.mark.parametrize("curve,grp", [("bls12_377", 0), ("bls12_377", 1), ("bw6_761", 0), ("bw6_761", 1)])
def test_dot_product_vec_matches_oracle(curve, grp):
    cv = R.CURVES[curve]
    g = cv.g2 if grp else cv.g1
    rng = random.Random(70 + grp)
    n = 12
    bases = [g.mul(g.gen, rng.randrange(1, g.r)) for _ in range(n)]
    bases[5] = None
    rows = _rows(rng, cv.r, 20, n)
    rows[3] = []                                        # unconstrained variable -> infinity
    rows[4] = [(1, 2), (cv.r - 1, 2)]                   # cancels
    rows[6] = [(1, 7), (1, 7)]                          # doubling inside the sum
    rows[7] = [(5, 1), (cv.r - 5, 1)]                   # general entries cancelling
    want = g.write_batch(R.dot_product_vec(g, rows, bases), True)
    for cin in (False, True):
        got = S.qap_dot_product(CID[curve], grp, g.write_batch(bases, cin), cin, rows, True)
        assert got == want


def test_dot_product_long_row_and_blocks():
    """The constant-one variable touches every constraint: one row much longer than a segment, next to many short
    rows; expected value through known discrete logs (bases = s_i G) and the C++ oracle."""
    cv, cid, g = R.BLS12_377, S.BLS12_377, R.BLS12_377.g1
    rng = random.Random(9)
    n = 1 << 12
    s = [rng.randrange(cv.r) for _ in range(n)]
    gens = g.encode(g.gen, False) * n
    bases = O.apply_powers(cid, 0, gens, False, 3, False, n, powers=s)
    rows = [[(rng.choice([1, cv.r - 1, rng.randrange(cv.r)]), i) for i in range(n)]]       # 4096 entries
    rows += [[(1, rng.randrange(n)), (rng.randrange(cv.r), rng.randrange(n))] for _ in range(3000)]
    rows += [[(rng.randrange(cv.r), i) for i in range(0, n, 3)]]
    expect = [sum(c * s[i] for c, i in row) % cv.r for row in rows]
    want = O.apply_powers(cid, 0, g.encode(g.gen, False) * len(rows), False, 3, True, len(rows), powers=expect)
    assert S.qap_dot_product(cid, 0, bases, False, rows, True) == want


def test_qap_eval_small_circuit():
    """polynomial.rs:11-47 on a toy QAP: a/b/c matrices over 8 constraints and 6 variables (2 public)."""
    cv, cid = R.BLS12_377, S.BLS12_377
    g1, g2 = cv.g1, cv.g2
    rng = random.Random(5)
    m, num_vars, num_inputs = 8, 6, 2
    co1 = [g1.mul(g1.gen, rng.randrange(1, cv.r)) for _ in range(m)]
    co2 = [g2.mul(g2.gen, rng.randrange(1, cv.r)) for _ in range(m)]
    al = [g1.mul(g1.gen, rng.randrange(1, cv.r)) for _ in range(m)]
    be = [g1.mul(g1.gen, rng.randrange(1, cv.r)) for _ in range(m)]

    def matrix():
        return [[(rng.choice([1, cv.r - 1, rng.randrange(cv.r)]), rng.randrange(num_vars)) for _ in range(rng.randrange(1, 4))]
                for _ in range(m)]
    at, bt, ct = (R.process_matrix(matrix(), num_vars) for _ in range(3))
    a_g1, b_g1, b_g2, gamma_abc, l = R.qap_eval(cv, co1, co2, al, be, at, bt, ct, num_inputs)
    b1 = g1.write_batch(co1, False)
    assert S.qap_dot_product(cid, 0, b1, False, at, False) == g1.write_batch(a_g1, False)
    assert S.qap_dot_product(cid, 0, b1, False, bt, False) == g1.write_batch(b_g1, False)
    assert S.qap_dot_product(cid, 1, g2.write_batch(co2, False), False, bt, False) == g2.write_batch(b_g2, False)
    # dot_product_ext: one call over [beta_coeffs | alpha_coeffs | coeffs_g1]
    ext_rows = [[(c, i) for c, i in a] + [(c, i + m) for c, i in b] + [(c, i + 2 * m) for c, i in cc]
                for a, b, cc in zip(at, bt, ct)]
    ext = S.qap_dot_product(cid, 0, g1.write_batch(be + al + co1, False), False, ext_rows, False)
    assert ext == g1.write_batch(gamma_abc + l, False)


def test_dot_product_rejects_bad_matrix():
    cv, cid, g = R.BLS12_377, S.BLS12_377, R.BLS12_377.g1
    bases = g.write_batch([g.gen] * 4, False)
    with pytest.raises(S.InvalidLength):       # coeffs[ind] out of bounds panics in the reference
        S.qap_dot_product(cid, 0, bases, False, [[(1, 4)]], False)
    with pytest.raises(S.InvalidData):         # non-canonical scalar
        S.qap_dot_product(cid, 0, bases, False, [[(cv.r, 0)]], False)
    assert S.qap_dot_product(cid, 0, bases, False, [[], []], True) == g.encode(None, True) * 2
