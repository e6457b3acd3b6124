"""MNT4-753 / MNT6-753 (the reference exposes both: setup-utils/src/converters.rs:18-45, phase1-cli/src/bin/phase1.rs:146-151).
The curve constants were recalled, so they are VERIFIED here rather than trusted; then the device headers (a != 0
doubling, Fq2 with non-residue 13, Fq3 with non-residue 11, Tonelli-Shanks roots, the byte-granular 95-byte codec) are
checked against the big-integer oracle through the host emulation build."""
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


def _is_prime(n, rounds=24):
    if n < 2:
        return False
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    rng = random.Random(n & 0xffff)
    for _ in range(rounds):
        a = rng.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def test_recalled_constants_are_consistent():
    q, r = R.MNT753_Q, R.MNT753_R
    assert q.bit_length() == r.bit_length() == 753 and _is_prime(q) and _is_prime(r)
    m4, m6 = R.curve_by_name("mnt4_753"), R.curve_by_name("mnt6_753")
    rng = random.Random(4)
    # the curves form a cycle of prime order: #E4(Fq) = r, #E6(Fr) = q — r * P = O for random points of either curve
    # (a prime-order group inside the Hasse interval: the order IS that prime, the G1 cofactor is 1)
    for cv in (m4, m6):
        g = cv.g1
        assert g.on_curve(g.gen) and g.mul(g.gen, cv.r) is None
        for _ in range(3):
            while True:
                x = rng.randrange(g.F.p)
                y = g.F.sqrt(g.rhs(x))
                if y is not None:
                    break
            assert g.mul((x, y), cv.r) is None
        lo, hi = g.F.p + 1 - 2 * int(g.F.p ** 0.5) - 2, g.F.p + 1 + 2 * int(g.F.p ** 0.5) + 2
        assert lo < cv.r < hi
    # extension-field non-residues and the twists
    assert pow(13, (q - 1) // 2, q) == q - 1                 # 13 is a quadratic non-residue mod q  (MNT4 Fq2)
    assert (r - 1) % 3 == 0 and pow(11, (r - 1) // 3, r) != 1  # 11 is a cubic non-residue mod r      (MNT6 Fq3)
    for cv in (m4, m6):
        g2 = cv.g2
        assert g2.order_full % cv.r == 0 and g2.on_curve(g2.gen) and g2.mul(g2.gen, cv.r) is None
        # a random twist point has order dividing the twist's group order, and cofactor clearing lands in G2
        P = g2.mul(g2.gen, 7)
        assert g2.mul(P, cv.r) is None
    assert m4.g2.a == (26, 0) and m4.g2.b == (0, 13 * R.MNT4_B % q)
    assert m6.g2.a == (0, 0, 11) and m6.g2.b == (11 * R.MNT6_B % r, 0, 0)
    assert [m4.g1.size(c) for c in (True, False)] == [95, 190] and [m4.g2.size(c) for c in (True, False)] == [190, 380]
    assert [m6.g1.size(c) for c in (True, False)] == [95, 190] and [m6.g2.size(c) for c in (True, False)] == [285, 570]
    assert m4.fr_size == m6.fr_size == 95


GROUPS = [("mnt4_753", "g1", 0), ("mnt4_753", "g2", 1), ("mnt6_753", "g1", 2), ("mnt6_753", "g2", 3)]


@pytest.mark.parametrize("cname,gname,gid", GROUPS)
def test_device_headers_match_the_oracle(L, cname, gname, gid):
    cv = R.curve_by_name(cname)
    g = getattr(cv, gname)
    rng = random.Random(gid)
    us, cs = g.size(False), g.size(True)
    pts = [g.mul(g.gen, rng.randrange(1, cv.r)) for _ in range(3)]
    out_u, out_c = ctypes.create_string_buffer(us), ctypes.create_string_buffer(cs)
    for P in pts:
        # codec: compressed -> uncompressed needs the square root (Tonelli-Shanks in Fq / Fq2 / Fq3) and the sign rule
        assert L.emul_mnt_op(gid, 0, g.encode(P, True), 1, R.FULL, None, out_u, 0) == 0
        assert out_u.raw == g.encode(P, False)
        assert L.emul_mnt_op(gid, 0, g.encode(P, False), 0, R.NO, None, out_c, 1) == 0
        assert out_c.raw == g.encode(P, True)
    # infinity, both flag layouts
    assert L.emul_mnt_op(gid, 0, g.encode(None, True), 1, R.NO, None, out_u, 0) == 0 and out_u.raw == g.encode(None, False)
    assert L.emul_mnt_op(gid, 0, g.encode(None, True), 1, R.FULL, None, out_u, 0) == -3  # PointAtInfinity
    # non-canonical coordinate (>= p) and both flags set
    bad = bytearray(g.encode(pts[0], True))
    bad[-1] |= 0xC0
    assert L.emul_mnt_op(gid, 0, bytes(bad), 1, R.NO, None, out_u, 0) == -2
    big = bytearray(g.encode(pts[0], False))
    big[:95] = (g.F.p if g.F.degree == 1 else g.F.p).to_bytes(95, "little")
    assert L.emul_mnt_op(gid, 0, bytes(big), 0, R.NO, None, out_u, 0) == -1
    # scalar multiplication (double-and-add over the a != 0 formulas) incl. k = 0, 1, r - 1
    for k in (0, 1, cv.r - 1, rng.randrange(cv.r)):
        assert L.emul_mnt_op(gid, 1, g.encode(pts[0], False), 0, R.NO, k.to_bytes(95, "little"), out_u, 0) == 0
        assert out_u.raw == g.encode(g.mul(pts[0], k), False), k
    # additions incl. the exceptional cases P + P, P + (-P), O + Q, P + O
    cases = [(pts[0], pts[1]), (pts[0], pts[0]), (pts[0], g.neg(pts[0])), (None, pts[1]), (pts[0], None)]
    for P, Q in cases:
        assert L.emul_mnt_op(gid, 2, g.encode(P, False), 0, R.NO, g.encode(Q, False), out_u, 0) == 0
        assert out_u.raw == g.encode(g.add(P, Q), False)
    # subgroup test by r-multiplication: G2 twist points outside the order-r subgroup are rejected
    assert L.emul_mnt_op(gid, 3, g.encode(pts[2], False), 0, R.NO, None, out_u, 0) == 1
    if gname == "g2":
        c = 1
        while True:
            x = (c, 1) if g.F.degree == 2 else (c, 1, 0)
            y = g.F.sqrt(g.rhs(x))
            if y is not None and g.mul((x, y), cv.r) is not None:
                break
            c += 1
        assert L.emul_mnt_op(gid, 3, g.encode((x, y), False), 0, R.NO, None, out_u, 0) == 0
        assert L.emul_mnt_op(gid, 0, g.encode((x, y), False), 0, R.FULL, None, out_u, 0) == -1  # Validate::Yes fails
