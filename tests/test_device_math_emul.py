"""Host-emulated build of the DEVICE math headers (ptx.cuh emulates the PTX carry flag) against
pyref: checks the exact limb algorithms the kernels run (Montgomery even/odd multiplier, Fp2, Jacobian
formulas incl. exceptional cases, codec, sqrt) on a box without a GPU."""
import ctypes
import os
import random
import subprocess

import pytest

import pyref as R

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def L():
    src = os.path.join(HERE, "emul", "emul.cpp")
    so = os.path.join(HERE, "emul", "libemul.so")
    hdrs = os.path.join(HERE, "..", "snark-setup_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(hdrs, f)) for f in os.listdir(hdrs) if f.endswith(".cuh"))
    if not os.path.exists(so) or os.path.getmtime(so) < max(newest, os.path.getmtime(src)):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    return ctypes.CDLL(so)


def tb(v, n):
    return v.to_bytes(n, "little")


FIELDS = {0: (R.BLS12_377_Q, 48), 1: (R.BLS12_377_R, 32), 2: (R.BW6_761_Q, 96)}


@pytest.mark.parametrize("fid", [0, 1, 2])
def test_fp_ops(L, fid):
    p, nb = FIELDS[fid]
    F = R.Fp(p)
    rng = random.Random(fid)
    for it in range(40):
        a, b = rng.randrange(p), rng.randrange(p)
        if it == 0:
            a = 0
        if it == 1:
            a = b = p - 1
        if it == 2:
            a, b = 1, p - 1
        exp = {0: a * b % p, 1: (a + b) % p, 2: (a - b) % p, 3: (-a) % p, 4: pow(a, p - 2, p), 6: a * a % p}
        for op, e in exp.items():
            out = ctypes.create_string_buffer(nb)
            assert L.emul_field_op(fid, op, tb(a, nb), tb(b, nb), out) == 0
            assert out.raw == tb(e, nb), (fid, op, it)
        out = ctypes.create_string_buffer(nb)
        rc = L.emul_field_op(fid, 5, tb(a, nb), tb(b, nb), out)
        if F.sqrt(a) is None:
            assert rc == 1
        else:
            g = int.from_bytes(out.raw, "little")
            assert rc == 0 and g * g % p == a


def test_fp2_ops(L):
    p, nb = R.BLS12_377_Q, 48
    f2 = R.BLS12_377.g2.F
    rng = random.Random(5)
    for it in range(30):
        a = (rng.randrange(p), rng.randrange(p))
        b = (rng.randrange(p), rng.randrange(p))
        if it == 0:
            a = (rng.randrange(p), 0)
        if it == 1:
            a = (0, rng.randrange(p))
        if it == 2:
            a = (5, 0)
        if it == 3:
            a = (p - 5, 0)
        ab, bb = tb(a[0], nb) + tb(a[1], nb), tb(b[0], nb) + tb(b[1], nb)
        exp = {0: f2.mul(a, b), 1: f2.add(a, b), 2: f2.sub(a, b), 3: f2.neg(a), 4: f2.inv(a), 6: f2.sqr(a)}
        for op, e in exp.items():
            out = ctypes.create_string_buffer(2 * nb)
            assert L.emul_field_op(3, op, ab, bb, out) == 0
            assert out.raw == tb(e[0], nb) + tb(e[1], nb), (op, it)
        out = ctypes.create_string_buffer(2 * nb)
        rc = L.emul_field_op(3, 5, ab, bb, out)
        s = f2.sqrt(a)
        if s is None:
            assert rc == 1
        else:
            g = (int.from_bytes(out.raw[:nb], "little"), int.from_bytes(out.raw[nb:], "little"))
            assert rc == 0 and f2.sqr(g) == a


GROUPS = [R.BLS12_377.g1, R.BLS12_377.g2, R.BW6_761.g1, R.BW6_761.g2]


@pytest.mark.parametrize("gid", [0, 1, 2, 3])
def test_points(L, gid):
    g = GROUPS[gid]
    rng = random.Random(gid + 10)
    nb = (g.r.bit_length() + 7) // 8
    for it in range(4):
        P = g.mul(g.gen, rng.randrange(1, g.r))
        k = [1, 0, g.r - 1, rng.randrange(g.r)][it]
        Q = g.mul(P, k)
        for ci in (0, 1):
            for co in (0, 1):
                out = ctypes.create_string_buffer(g.size(co))
                rc = L.emul_point_mul(gid, g.encode(P, ci), ci, R.FULL if it == 3 else R.NO, tb(k, nb), g.r.bit_length(), out, co)
                assert rc == 0 and out.raw == g.encode(Q, co), (gid, it, ci, co)
        P2 = g.mul(g.gen, rng.randrange(1, g.r))
        for (A, B) in ((P, P2), (P, P), (P, g.neg(P)), (None, P), (P, None), (None, None)):
            for which in (0, 1):
                out = ctypes.create_string_buffer(g.usize)
                assert L.emul_point_add(gid, g.encode(A, 0), g.encode(B, 0), which, out) == 0
                assert out.raw == g.encode(g.add(A, B), 0)
    out = ctypes.create_string_buffer(g.usize)
    inf = g.encode(None, 1)
    assert L.emul_point_mul(gid, inf, 1, R.ONLY_NON_ZERO, tb(5, nb), 8, out, 0) == 3
    assert L.emul_point_mul(gid, inf, 1, R.NO, tb(5, nb), 8, out, 0) == 0 and out.raw == g.encode(None, 0)
    bad = bytearray(g.encode(g.gen, 1))
    bad[-1] |= 0xC0
    assert L.emul_point_mul(gid, bytes(bad), 1, R.NO, tb(5, nb), 8, out, 0) == 2
    bad = bytearray(b"\xff" * g.csize)
    bad[-1] = 0x3F
    assert L.emul_point_mul(gid, bytes(bad), 1, R.NO, tb(5, nb), 8, out, 0) == 1


def test_fp_mul2_sum_of_two_products(L):
    """fp_mul2 = (a b + c d) / R with one interleaved reduction (fp.cuh), incl. operands at their documented bounds:
    a, c up to 2p and the scanned operand d up to 5p."""
    p, nb = FIELDS[0]
    rng = random.Random(77)
    for it in range(60):
        a, b, c, d = (rng.randrange(p) for _ in range(4))
        if it == 0:
            a = b = c = d = p - 1
        if it == 1:
            a, b, c, d = 0, 0, p - 1, p - 1
        if it == 2:
            a, b, c, d = p - 1, 1, 1, p - 1
        for a_plus_p in (0, 1):
            for d5 in (0, 1):
                out = ctypes.create_string_buffer(nb)
                assert L.emul_fp_mul2(tb(a, nb), tb(b, nb), tb(c, nb), tb(d, nb), a_plus_p, d5, out) == 0
                want = (a * b + c * (5 * d if d5 else d)) % p
                assert out.raw == tb(want, nb), (it, a_plus_p, d5)


def test_lane_split_fq2_product_and_square(L):
    """fp2l.cuh: the per-lane halves of the Fq2 product / square (even lane a0 b0 - 5 a1 b1, odd lane a0 b1 + a1 b0;
    complex squaring split as (a0 + a1)(a0 - 5 a1) | a0 a1) against the big-integer Fq2."""
    p, nb = FIELDS[0]
    F2 = R.BLS12_377.g2.F
    rng = random.Random(78)
    for it in range(60):
        a = (rng.randrange(p), rng.randrange(p))
        b = (rng.randrange(p), rng.randrange(p))
        if it == 0:
            a = b = (p - 1, p - 1)
        if it == 1:
            a, b = (0, p - 1), (p - 1, 0)
        if it == 2:
            a, b = (0, 0), (5, 7)
        ab = tb(a[0], nb) + tb(a[1], nb)
        bb = tb(b[0], nb) + tb(b[1], nb)
        out = ctypes.create_string_buffer(2 * nb)
        assert L.emul_fp2l_op(0, ab, bb, out) == 0
        m = F2.mul(a, b)
        assert out.raw == tb(m[0], nb) + tb(m[1], nb), it
        assert L.emul_fp2l_op(1, ab, bb, out) == 0
        q = F2.sqr(a)
        assert out.raw == tb(q[0], nb) + tb(q[1], nb), it
        # the out-of-line per-thread units (row-interleaved base products) give the same values
        assert L.emul_fp2_unit(0, ab, bb, out) == 0
        assert out.raw == tb(m[0], nb) + tb(m[1], nb), it
        assert L.emul_fp2_unit(1, ab, bb, out) == 0
        assert out.raw == tb(q[0], nb) + tb(q[1], nb), it
