"""Pure-Python big-int restatement of the snark-setup batch-exponentiation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (snark-setup_b200/) may import this file;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.

This is the *second*, independent restatement (the first is oracle/oracle.cpp, Montgomery / 64-bit
limbs).  It uses nothing but Python integers and textbook affine formulas so that a disagreement
with oracle.cpp or with the CUDA path points at an arithmetic bug rather than a shared mistake.

PARITY UNPINNED: the reference (nimiq/snark-setup) holds no golden vectors for this path and its
arithmetic lives in un-vendored arkworks 0.4 (ark-ff/ark-ec 0.4.2 @ paberr/algebra pb/0.4 1ab82cb7,
ark-serialize 0.4.2, ark-bls12-377 / ark-bw6-761 0.4.0; /root/reference/Cargo.lock:80-82,103-105,
151-153,187-189,459-461).  What is restated here is the published arkworks behaviour at the
reference's call sites:

  * generate_powers_of_tau      setup-utils/src/helpers.rs:32-37
  * batch_exp / batch_mul       setup-utils/src/helpers.rs:56-59,75-140
  * merge_pairs / power_pairs   setup-utils/src/helpers.rs:371-390
  * read_batch / read_element   setup-utils/src/io/read.rs:57-73,110-135
  * write_batch / write_element setup-utils/src/io/write.rs:30-67
  * check_subgroup              setup-utils/src/elements.rs:123-150
  * apply_powers                phase1/src/helpers/buffers.rs:77-97
  * iter_chunk                  phase1/src/helpers/buffers.rs:22-73
  * Phase1Parameters sizes      phase1/src/objects/parameters.rs:115-294
  * Phase1::computation         phase1/src/computation.rs:16-193 (Groth16 branch)
  * to_coeffs / h_query / Groth16Params::new + ::write   setup-utils/src/groth16_utils.rs:44-168
  * same_ratio / check_same_ratio                         setup-utils/src/helpers.rs:406-424 (verdict only: Tate pairing)
  * dot_product(_vec) / eval / process_matrix             phase2/src/polynomial.rs:11-94, parameters.rs:96-105
"""
from __future__ import annotations

# ----------------------------------------------------------------------------------------------
# error codes (mirror setup_utils::Error, setup-utils/src/errors.rs:11-38)
# ----------------------------------------------------------------------------------------------
class SetupError(Exception):
    pass


class InvalidData(SetupError):          # ZexeSerializationError(InvalidData)
    pass


class UnexpectedFlags(SetupError):      # ZexeSerializationError(UnexpectedFlags)
    pass


class PointAtInfinity(SetupError):      # Error::PointAtInfinity
    pass


class IncorrectSubgroup(SetupError):    # Error::IncorrectSubgroup
    pass


class InvalidLength(SetupError):        # Error::InvalidLength
    pass


# CheckForCorrectness (setup-utils/src/elements.rs:18-23)
FULL, ONLY_NON_ZERO, ONLY_IN_GROUP, NO = 0, 1, 2, 3


# ----------------------------------------------------------------------------------------------
# fields
# ----------------------------------------------------------------------------------------------
class Fp:
    """Prime field; elements are Python ints in [0, p)."""

    def __init__(self, p: int):
        self.p = p
        self.bits = p.bit_length()
        self.degree = 1
        # two-adicity data for Tonelli-Shanks
        s, t = 0, p - 1
        while t % 2 == 0:
            s += 1
            t //= 2
        self.two_adicity, self.t = s, t
        z = 2
        while pow(z, (p - 1) // 2, p) != p - 1:
            z += 1
        self.qnr = z

    zero = 0
    one = 1

    def add(self, a, b): return (a + b) % self.p
    def sub(self, a, b): return (a - b) % self.p
    def neg(self, a): return (-a) % self.p
    def mul(self, a, b): return (a * b) % self.p
    def sqr(self, a): return (a * a) % self.p
    def inv(self, a): return pow(a, self.p - 2, self.p)
    def is_zero(self, a): return a == 0
    def from_int(self, v): return v % self.p

    def sqrt(self, a):
        """Any square root or None."""
        p = self.p
        if a == 0:
            return 0
        if pow(a, (p - 1) // 2, p) != 1:
            return None
        if p % 4 == 3:
            return pow(a, (p + 1) // 4, p)
        # Tonelli-Shanks
        s, t = self.two_adicity, self.t
        z = pow(self.qnr, t, p)
        w = pow(a, (t - 1) // 2, p)
        x = a * w % p
        b = x * w % p
        v = s
        while b != 1:
            k, b2k = 0, b
            while b2k != 1:
                b2k = b2k * b2k % p
                k += 1
            wj = z
            for _ in range(v - k - 1):
                wj = wj * wj % p
            z = wj * wj % p
            b = b * z % p
            x = x * wj % p
            v = k
        return x

    def gt(self, a, b):
        """a > b as canonical integers."""
        return a > b

    # canonical little-endian bytes, ceil((bits + flag_bits)/8) long (ark-ff Fp serialize_with_flags)
    def size(self, flag_bits=0):
        return (self.bits + flag_bits + 7) // 8

    def to_bytes(self, a, flags=0, flag_bits=0):
        b = bytearray(a.to_bytes(self.size(flag_bits), "little"))
        b[-1] |= flags
        return bytes(b)

    def from_bytes(self, b, flag_bits=0):
        """Returns (element, flags_byte). Raises InvalidData for non-canonical values."""
        b = bytearray(b)
        assert len(b) == self.size(flag_bits)
        flags = 0
        if flag_bits:
            mask = (0xFF << (8 - flag_bits)) & 0xFF
            flags = b[-1] & mask
            b[-1] &= ~mask & 0xFF
            if flag_bits == 2 and flags == 0xC0:
                raise UnexpectedFlags()
        v = int.from_bytes(b, "little")
        if v >= self.p:
            raise InvalidData("field element >= modulus")
        return v, flags


class Fp2:
    """Fp[u]/(u^2 - nr); elements are (c0, c1)."""

    def __init__(self, base: Fp, nr: int):
        self.b = base
        self.p = base.p
        self.nr = nr % base.p
        self.degree = 2

    zero = (0, 0)
    one = (1, 0)

    def add(self, a, b): return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)
    def sub(self, a, b): return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)
    def neg(self, a): return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] + self.nr * a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def sqr(self, a): return self.mul(a, a)

    def inv(self, a):
        p = self.p
        n = (a[0] * a[0] - self.nr * a[1] * a[1]) % p
        ni = pow(n, p - 2, p)
        return (a[0] * ni % p, (-a[1]) * ni % p)

    def is_zero(self, a): return a[0] == 0 and a[1] == 0
    def from_int(self, v): return (v % self.p, 0)

    def sqrt(self, a):
        p, F = self.p, self.b
        if a[1] == 0:
            r = F.sqrt(a[0])
            if r is not None:
                return (r, 0)
            # a0 is a non-residue in Fp: sqrt is purely imaginary: (c1 u)^2 = c1^2 nr = a0
            r = F.sqrt(a[0] * pow(self.nr, p - 2, p) % p)
            return None if r is None else (0, r)
        norm = (a[0] * a[0] - self.nr * a[1] * a[1]) % p
        alpha = F.sqrt(norm)
        if alpha is None:
            return None
        half = pow(2, p - 2, p)
        delta = (a[0] + alpha) * half % p
        c0 = F.sqrt(delta)
        if c0 is None:
            delta = (a[0] - alpha) * half % p
            c0 = F.sqrt(delta)
            if c0 is None:
                return None
        c1 = a[1] * pow(2 * c0, p - 2, p) % p
        r = (c0, c1)
        assert self.sqr(r) == (a[0] % p, a[1] % p)
        return r

    def gt(self, a, b):
        """Lexicographic, c1 first then c0 (ark-ff QuadExtField Ord)."""
        if a[1] != b[1]:
            return a[1] > b[1]
        return a[0] > b[0]

    def size(self, flag_bits=0):
        return self.b.size(0) + self.b.size(flag_bits)

    def to_bytes(self, a, flags=0, flag_bits=0):
        return self.b.to_bytes(a[0]) + self.b.to_bytes(a[1], flags, flag_bits)

    def from_bytes(self, b, flag_bits=0):
        n0 = self.b.size(0)
        c0, _ = self.b.from_bytes(b[:n0], 0)
        c1, flags = self.b.from_bytes(b[n0:], flag_bits)
        return (c0, c1), flags


class Fp3:
    """Fp[u]/(u^3 - nr); elements are (c0, c1, c2)  (ark-ff CubicExtField, used by MNT6-753 G2)."""

    degree = 3

    def __init__(self, base: Fp, nr: int):
        self.b = base
        self.p = base.p
        self.nr = nr % base.p
        self.zero = (0, 0, 0)
        self.one = (1, 0, 0)
        self.bits = base.bits
        # Tonelli-Shanks data for the multiplicative group of order p^3 - 1 = 2^s * t
        n = self.p ** 3 - 1
        self.s = 0
        while n % 2 == 0:
            n //= 2
            self.s += 1
        self.t = n
        self._z = None

    def add(self, a, b): return tuple((x + y) % self.p for x, y in zip(a, b))
    def sub(self, a, b): return tuple((x - y) % self.p for x, y in zip(a, b))
    def neg(self, a): return tuple((-x) % self.p for x in a)

    def mul(self, a, b):
        p, nr = self.p, self.nr
        a0, a1, a2 = a
        b0, b1, b2 = b
        return ((a0 * b0 + nr * (a1 * b2 + a2 * b1)) % p,
                (a0 * b1 + a1 * b0 + nr * a2 * b2) % p,
                (a0 * b2 + a1 * b1 + a2 * b0) % p)

    def sqr(self, a): return self.mul(a, a)

    def pow(self, a, e):
        r = self.one
        for bit in bin(e)[2:]:
            r = self.mul(r, r)
            if bit == "1":
                r = self.mul(r, a)
        return r

    def inv(self, a):
        # a^-1 = a^(p^3 - 2); slow but this is the oracle
        return self.pow(a, self.p ** 3 - 2)

    def is_zero(self, a): return a == (0, 0, 0)
    def from_int(self, v): return (v % self.p, 0, 0)

    def sqrt(self, a):
        """Any square root or None (Tonelli-Shanks in Fp3*)."""
        if a == self.zero:
            return a
        if self.pow(a, (self.p ** 3 - 1) // 2) != self.one:
            return None
        if self._z is None:
            c = 1
            while True:  # a quadratic non-residue of the form (c, 1, 0)
                cand = (c, 1, 0)
                if self.pow(cand, (self.p ** 3 - 1) // 2) != self.one:
                    break
                c += 1
            self._z = self.pow(cand, self.t)
        z, v = self._z, self.s
        w = self.pow(a, (self.t - 1) // 2)
        x = self.mul(a, w)
        b = self.mul(x, w)
        while b != self.one:
            k, b2k = 0, b
            while b2k != self.one:
                b2k = self.sqr(b2k)
                k += 1
            wj = z
            for _ in range(v - k - 1):
                wj = self.sqr(wj)
            z = self.sqr(wj)
            b = self.mul(b, z)
            x = self.mul(x, wj)
            v = k
        assert self.sqr(x) == a
        return x

    def gt(self, a, b):
        """Lexicographic, c2 first, then c1, then c0 (ark-ff CubicExtField Ord)."""
        for i in (2, 1, 0):
            if a[i] != b[i]:
                return a[i] > b[i]
        return False

    def size(self, flag_bits=0):
        return 2 * self.b.size(0) + self.b.size(flag_bits)

    def to_bytes(self, a, flags=0, flag_bits=0):
        return self.b.to_bytes(a[0]) + self.b.to_bytes(a[1]) + self.b.to_bytes(a[2], flags, flag_bits)

    def from_bytes(self, b, flag_bits=0):
        n0 = self.b.size(0)
        c0, _ = self.b.from_bytes(b[:n0], 0)
        c1, _ = self.b.from_bytes(b[n0:2 * n0], 0)
        c2, flags = self.b.from_bytes(b[2 * n0:], flag_bits)
        return (c0, c1, c2), flags


# ----------------------------------------------------------------------------------------------
# short-Weierstrass groups (a = 0 for BLS12-377 / BW6-761, a != 0 for the MNT curves)
# ----------------------------------------------------------------------------------------------
FLAG_NEG = 0x80   # SWFlags::YIsNegative
FLAG_INF = 0x40   # SWFlags::PointAtInfinity


class Group:
    """y^2 = x^3 + a x + b over field F (a = 0 unless given). Points: None (identity) or (x, y)."""

    def __init__(self, name, F, b, gen, r, a=None, gen_is_reference=True):
        self.name, self.F, self.b, self.gen, self.r = name, F, b, gen, r
        self.a = F.zero if a is None else a
        self.a_is_zero = F.is_zero(self.a)
        # False when `gen` is only SOME generator of the order-r group (the reference's constant is unknown here)
        self.gen_is_reference = gen_is_reference
        self.usize = 2 * F.size(0) if F.degree == 1 else F.size(0) + F.size(2)
        self.usize = F.size(0) + F.size(2)
        self.csize = F.size(2)

    # -- group law (affine, textbook) --
    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sqr(y) == self.rhs(x)

    def rhs(self, x):
        """x^3 + a x + b"""
        F = self.F
        return F.add(F.mul(F.add(F.sqr(x), self.a), x), self.b)

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        F = self.F
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if P[1] == Q[1]:
                if F.is_zero(P[1]):
                    return None
                x2 = F.sqr(P[0])
                lam = F.mul(F.add(F.add(F.add(x2, x2), x2), self.a), F.inv(F.add(P[1], P[1])))
            else:
                return None
        else:
            lam = F.mul(F.sub(Q[1], P[1]), F.inv(F.sub(Q[0], P[0])))
        x3 = F.sub(F.sub(F.sqr(lam), P[0]), Q[0])
        y3 = F.sub(F.mul(lam, F.sub(P[0], x3)), P[1])
        return (x3, y3)

    def mul(self, P, k):
        """k * P for a non-negative integer k (not reduced: callers decide)."""
        # Jacobian double-and-add for speed; affine result.
        if P is None or k == 0:
            return None
        F = self.F
        X, Y, Z = P[0], P[1], F.one
        inf = True
        AX = AY = AZ = None
        for bit in bin(k)[2:]:
            if not inf:
                AX, AY, AZ = self._jdbl(AX, AY, AZ)
            if bit == "1":
                if inf:
                    AX, AY, AZ, inf = X, Y, Z, False
                else:
                    AX, AY, AZ = self._jadd_affine(AX, AY, AZ, P)
                    if AX is None:
                        inf = True
        if inf:
            return None
        return self._to_affine(AX, AY, AZ)

    def _to_affine(self, X, Y, Z):
        F = self.F
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def _jdbl(self, X, Y, Z):
        F = self.F
        if F.is_zero(Z) or F.is_zero(Y):
            return (F.one, F.one, F.zero)
        if not self.a_is_zero:  # dbl-2007-bl
            XX, YY, ZZ = F.sqr(X), F.sqr(Y), F.sqr(Z)
            YYYY = F.sqr(YY)
            t = F.sub(F.sub(F.sqr(F.add(X, YY)), XX), YYYY)
            S = F.add(t, t)
            M = F.add(F.add(F.add(XX, XX), XX), F.mul(self.a, F.sqr(ZZ)))
            X3 = F.sub(F.sqr(M), F.add(S, S))
            Y8 = F.add(YYYY, YYYY); Y8 = F.add(Y8, Y8); Y8 = F.add(Y8, Y8)
            Y3 = F.sub(F.mul(M, F.sub(S, X3)), Y8)
            Z3 = F.sub(F.sub(F.sqr(F.add(Y, Z)), YY), ZZ)
            return (X3, Y3, Z3)
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        t = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
        D = F.add(t, t)
        E = F.add(F.add(A, A), A)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def _jadd_affine(self, X1, Y1, Z1, Q):
        """Jacobian + affine; returns (None,None,None) for identity result."""
        F = self.F
        if F.is_zero(Z1):
            return (Q[0], Q[1], F.one)
        Z1Z1 = F.sqr(Z1)
        U2 = F.mul(Q[0], Z1Z1)
        S2 = F.mul(F.mul(Q[1], Z1), Z1Z1)
        if U2 == X1:
            if S2 == Y1:
                return self._jdbl(X1, Y1, Z1)
            return (None, None, None)
        H = F.sub(U2, X1)
        R = F.sub(S2, Y1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(X1, HH)
        X3 = F.sub(F.sub(F.sqr(R), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(Y1, HHH))
        Z3 = F.mul(Z1, H)
        return (X3, Y3, Z3)

    def in_subgroup(self, P):
        """p.mul_bigint(r).is_zero()  (setup-utils/src/elements.rs:138-142)."""
        return self.mul(P, self.r) is None

    # -- canonical serialisation (ark-ec 0.4 SWCurveConfig::{serialize,deserialize}_with_mode) --
    def size(self, compressed):
        return self.csize if compressed else self.usize

    def flags_of(self, P):
        if P is None:
            return FLAG_INF
        F = self.F
        y = P[1]
        return FLAG_NEG if F.gt(y, F.neg(y)) else 0

    def encode(self, P, compressed):
        F = self.F
        fl = self.flags_of(P)
        x, y = (F.zero, F.zero) if P is None else P
        if compressed:
            return F.to_bytes(x, fl, 2)
        return F.to_bytes(x) + F.to_bytes(y, fl, 2)

    def decode(self, b, compressed, check=NO):
        """read_element (setup-utils/src/io/read.rs:57-73) for one element."""
        F = self.F
        validate = check in (FULL, ONLY_IN_GROUP)
        if compressed:
            x, fl = F.from_bytes(b, 2)
            if fl == (FLAG_NEG | FLAG_INF):
                raise UnexpectedFlags()
            if fl & FLAG_INF:
                P = None
            else:
                y = F.sqrt(self.rhs(x))
                if y is None:
                    raise InvalidData("x^3+ax+b is not a square")
                ny = F.neg(y)
                lo, hi = (y, ny) if F.gt(ny, y) else (ny, y)
                P = (x, hi if fl & FLAG_NEG else lo)
        else:
            nx = F.size(0)
            x, _ = F.from_bytes(b[:nx], 0)
            y, fl = F.from_bytes(b[nx:], 2)
            if fl == (FLAG_NEG | FLAG_INF):
                raise UnexpectedFlags()
            P = None if fl & FLAG_INF else (x, y)
        if P is not None and validate:
            if not self.on_curve(P) or not self.in_subgroup(P):
                raise InvalidData("point failed Validate::Yes")
        if check in (FULL, ONLY_NON_ZERO) and P is None:
            raise PointAtInfinity()
        return P

    # -- batches --
    def read_batch(self, buf, compressed, check=NO):
        sz = self.size(compressed)
        assert len(buf) % sz == 0
        return [self.decode(buf[i:i + sz], compressed, check) for i in range(0, len(buf), sz)]

    def write_batch(self, pts, compressed):
        return b"".join(self.encode(P, compressed) for P in pts)


class Curve:
    def __init__(self, name, g1: Group, g2: Group, r: int):
        self.name, self.g1, self.g2, self.r = name, g1, g2, r
        self.fr = Fp(r)
        self.fr_size = (r.bit_length() + 7) // 8


# ----------------------------------------------------------------------------------------------
# curve constants (re-verified by tests/test_pyref.py: on-curve, r*G = O)
# ----------------------------------------------------------------------------------------------
BLS12_377_Q = 0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001
BLS12_377_R = 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001
BW6_761_Q = 0x122e824fb83ce0ad187c94004faff3eb926186a81d14688528275ef8087be41707ba638e584e91903cebaff25b423048689c8ed12f9fd9071dcd3dc73ebff2e98a116c25667a8f8160cf8aeeaf0a437e6913e6870000082f49d00000000008b


def _mk_bls12_377():
    fq = Fp(BLS12_377_Q)
    fq2 = Fp2(fq, -5)
    g1 = Group("bls12_377.g1", fq, 1,
               (0x008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef,
                0x01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6),
               BLS12_377_R)
    g2 = Group("bls12_377.g2", fq2,
               (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906),
               ((233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
                 140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118),
                (63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
                 149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491)),
               BLS12_377_R)
    return Curve("bls12_377", g1, g2, BLS12_377_R)


def _mk_bw6_761():
    fq = Fp(BW6_761_Q)
    g1 = Group("bw6_761.g1", fq, BW6_761_Q - 1,
               (0x01075b020ea190c8b277ce98a477beaee6a0cfb7551b27f0ee05c54b85f56fc779017ffac15520ac11dbfcd294c2e746a17a54ce47729b905bd71fa0c9ea097103758f9a280ca27f6750dd0356133e82055928aca6af603f4088f3af66e5b43d,
                0x0058b84e0a6fc574e6fd637b45cc2a420f952589884c9ec61a7348d2a2e573a3265909f1af7e0dbac5b8fa1771b5b806cc685d31717a4c55be3fb90b6fc2cdd49f9df141b3053253b2b08119cad0fb93ad1cb2be0b20d2a1bafc8f2db4e95363),
               BLS12_377_Q)
    g2 = Group("bw6_761.g2", fq, 4,
               (0x0110133241d9b816c852a82e69d660f9d61053aac5a7115f4c06201013890f6d26b41c5dab3da268734ec3f1f09feb58c5bbcae9ac70e7c7963317a300e1b6bace6948cb3cd208d700e96efbc2ad54b06410cf4fe1bf995ba830c194cd025f1c,
                0x0017c3357761369f8179eb10e4b6d2dc26b7cf9acec2181c81a78e2753ffe3160a1d86c80b95a59c94c97eb733293fef64f293dbd2c712b88906c170ffa823003ea96fcd504affc758aa2d3a3c5a02a591ec0594f9eac689eb70a16728c73b61),
               BLS12_377_Q)
    return Curve("bw6_761", g1, g2, BLS12_377_Q)


# MNT4-753 / MNT6-753 (setup-utils/src/converters.rs:18-45 exposes both).  The two 753-bit primes form a cycle:
# MNT4: Fq = MNT753_Q, Fr = MNT753_R;  MNT6: Fq = MNT753_R, Fr = MNT753_Q.  Curve coefficients and the G1 generators are
# the ark-mnt4-753 / ark-mnt6-753 0.4.0 constants as recalled; they are VERIFIED here, not trusted
# (tests/test_oracle_mnt_cpu.py): both curves have prime order = the other prime (r * P = O for random points, so the
# G1 cofactor is 1), the generators lie on their curves, 13 is a quadratic non-residue mod MNT753_Q and 11 a cubic
# non-residue mod MNT753_R, and the twists E'(Fq2) / E'(Fq3) have order divisible by r.  The G2 GENERATOR constants of
# arkworks could not be recalled: `gen` of the two G2 groups is a deterministically derived point of order r
# (gen_is_reference = False) — good for tests, not for Phase1::initialization, whose output must be arkworks' constant.
MNT753_Q = 41898490967918953402344214791240637128170709919953949071783502921025352812571106773058893763790338921418070971888253786114353726529584385201591605722013126468931404347949840543007986327743462853720628051692141265303114721689601
MNT753_R = 41898490967918953402344214791240637128170709919953949071783502921025352812571106773058893763790338921418070971888458477323173057491593855069696241854796396165721416325350064441470418137846398469611935719059908164220784476160001
MNT4_B = 28798803903456388891410036793299405764940372360099938340752576406393880372126970068421383312482853541572780087363938442377933706865252053507077543420534380486492786626556269083255657125025963825610840222568694137138741554679540
MNT6_B = 11625908999541321152027340224010374716841167701783584648338908235410859267060079819722747939267925389062611062156601938166010098747920378738927832658133625454260115409075816187555055859490253375704728027944315501122723426879114


def _derive_generator(group, cofactor):
    """first x = (counter, 1, ..) with a point on the curve, cofactor-cleared: SOME generator of the order-r group"""
    F = group.F
    c = 1
    while True:
        x = c if F.degree == 1 else ((c, 1) if F.degree == 2 else (c, 1, 0))
        y = F.sqrt(group.rhs(x))
        if y is not None:
            P = group.mul((x, y), cofactor)
            if P is not None:
                return P
        c += 1


def _mk_mnt4_753():
    fq = Fp(MNT753_Q)
    fq2 = Fp2(fq, 13)
    g1 = Group("mnt4_753.g1", fq, MNT4_B,
               (7790163481385331313124631546957228376128961350185262705123068027727518350362064426002432450801002268747950550964579198552865939244360469674540925037890082678099826733417900510086646711680891516503232107232083181010099241949569,
                6913648190367314284606685101150155872986263667483624713540251048208073654617802840433842931301128643140890502238233930290161632176167186761333725658542781350626799660920481723757654531036893265359076440986158843531053720994648),
               MNT753_R, a=2)
    # twist by u (u^2 = 13): a' = a u^2 = 26, b' = b u^3 = 13 b u
    g2 = Group("mnt4_753.g2", fq2, (0, 13 * MNT4_B % MNT753_Q), None, MNT753_R, a=(26, 0), gen_is_reference=False)
    t = MNT753_Q + 1 - MNT753_R
    g2.order_full = MNT753_Q ** 2 + 1 + (t * t - 2 * MNT753_Q)  # quadratic twist of E(Fq2)
    g2.cofactor = g2.order_full // MNT753_R
    g2.gen = _derive_generator(g2, g2.cofactor)
    return Curve("mnt4_753", g1, g2, MNT753_R)


def _mk_mnt6_753():
    fq = Fp(MNT753_R)
    fq3 = Fp3(fq, 11)
    g1 = Group("mnt6_753.g1", fq, MNT6_B,
               (3458420969484235708806261200128850544017070333833944116801482064540723268149235477762870414664917360605949659630933184751526227993647030875167687492714052872195770088225183259051403087906158701786758441889742618916006546636728,
                27460508402331965149626600224382137254502975979168371111640924721589127725376473514838234361114855175488242007431439074223827742813911899817930728112297763448010814764117701403540298764970469500339646563344680868495474127850569),
               MNT753_Q, a=11)
    # twist by u (u^3 = 11): a' = a u^2 = (0, 0, 11), b' = b u^3 = 11 b
    g2 = Group("mnt6_753.g2", fq3, (11 * MNT6_B % MNT753_R, 0, 0), None, MNT753_Q, a=(0, 0, 11), gen_is_reference=False)
    q = MNT753_R
    t = q + 1 - MNT753_Q
    t3 = t ** 3 - 3 * q * t
    g2.order_full = q ** 3 + 1 + t3  # quadratic twist of E(Fq3)
    g2.cofactor = g2.order_full // MNT753_Q
    g2.gen = _derive_generator(g2, g2.cofactor)
    return Curve("mnt6_753", g1, g2, MNT753_Q)


BLS12_377 = _mk_bls12_377()
BW6_761 = _mk_bw6_761()
CURVES = {"bls12_377": BLS12_377, "bw6_761": BW6_761}
_LAZY = {"mnt4_753": _mk_mnt4_753, "mnt6_753": _mk_mnt6_753}


def curve_by_name(name):
    """BLS12-377 / BW6-761 are built at import; the MNT curves on first use (their G2 generators take a cofactor
    multiplication to derive)."""
    if name not in CURVES:
        CURVES[name] = _LAZY[name]()
    return CURVES[name]


# ----------------------------------------------------------------------------------------------
# the hot path
# ----------------------------------------------------------------------------------------------
def generate_powers_of_tau(curve: Curve, tau: int, start: int, end: int):
    """setup-utils/src/helpers.rs:32-37 — each power independently as tau.pow([i])."""
    return [pow(tau, i, curve.r) for i in range(start, end)]


def batch_exp(group: Group, bases, exps, coeff=None):
    """setup-utils/src/helpers.rs:75-140: bases[i] <- (exps[i]*coeff?) * bases[i], affine."""
    if len(bases) != len(exps):
        raise InvalidLength(f"expected {len(bases)} got {len(exps)}")
    r = group.r
    out = []
    for P, e in zip(bases, exps):
        if coeff is not None:
            e = e * coeff % r
        out.append(group.mul(P, e % r))
    return out


def batch_mul(group: Group, bases, coeff):
    """setup-utils/src/helpers.rs:56-59."""
    return batch_exp(group, bases, [coeff] * len(bases))


def apply_powers(group: Group, inp, in_compressed, in_check, out_compressed, start, end, powers, coeff=None):
    """phase1/src/helpers/buffers.rs:77-97. Returns the bytes for output[start*out_sz .. end*out_sz]."""
    isz = group.size(in_compressed)
    elems = group.read_batch(inp[start * isz:end * isz], in_compressed, in_check)
    elems = batch_exp(group, elems, powers[:end - start], coeff)
    return group.write_batch(elems, out_compressed)


def msm(group: Group, pts, scalars):
    acc = None
    for P, k in zip(pts, scalars):
        acc = group.add(acc, group.mul(P, k % group.r))
    return acc


def merge_pairs(group: Group, v1, v2, rho):
    """setup-utils/src/helpers.rs:371-384 with the randomness made explicit."""
    assert len(v1) == len(v2) == len(rho)
    return msm(group, v1, rho), msm(group, v2, rho)


def power_pairs(group: Group, v, rho):
    """setup-utils/src/helpers.rs:388-390."""
    return merge_pairs(group, v[:-1], v[1:], rho)


def check_subgroup(group: Group, pts):
    """setup-utils/src/elements.rs:123-150 (all modes but `No` collapse to the direct loop)."""
    if not all(group.in_subgroup(P) for P in pts):
        raise IncorrectSubgroup()


# ----------------------------------------------------------------------------------------------
# phase1 layout + schedule
# ----------------------------------------------------------------------------------------------
FULL_MODE, CHUNKED_MODE = 0, 1
GROTH16, MARLIN = 0, 1


class Phase1Parameters:
    """phase1/src/objects/parameters.rs:115-294."""

    def __init__(self, curve: Curve, power: int, batch_size: int, mode=FULL_MODE, chunk_index=0, chunk_size=0,
                 proving_system=GROTH16):
        self.curve, self.total_size_in_log2, self.batch_size = curve, power, batch_size
        self.contribution_mode, self.chunk_index, self.chunk_size = mode, chunk_index, chunk_size
        self.proving_system = proving_system
        self.hash_size = 64
        self.powers_length = 1 << power
        self.powers_g1_length = (self.powers_length << 1) - 1
        upper = self.powers_g1_length if proving_system == GROTH16 else self.powers_length
        if mode == CHUNKED_MODE:
            start, end = chunk_index * chunk_size, (chunk_index + 1) * chunk_size
        else:
            start, end = 0, upper
        self.g1_chunk_size = upper - start if end > upper else end - start
        if proving_system == GROTH16:
            pl = self.powers_length
            if end > pl and start >= pl:
                self.other_chunk_size = 0
            elif end > pl:
                self.other_chunk_size = pl - start
            else:
                self.other_chunk_size = end - start
        else:
            self.other_chunk_size = 0
        g1u, g2u = curve.g1.usize, curve.g2.usize
        g1c, g2c = curve.g1.csize, curve.g2.csize
        self.public_key_size = 3 * g2c + 6 * g1c
        if proving_system == GROTH16:
            self.accumulator_size = (self.g1_chunk_size * g1u + self.other_chunk_size * (g2u + 2 * g1u) + g2u
                                     + self.hash_size)
            self.contribution_size = (self.g1_chunk_size * g1c + self.other_chunk_size * (g2c + 2 * g1c) + g2c
                                      + self.hash_size + self.public_key_size)
        else:
            extra_u = extra_c = 0
            if chunk_index == 0:
                extra_u = 3 * g1u + 3 * power * g1u + (power + 2) * g2u
                extra_c = 3 * g1c + 3 * power * g1c + (power + 2) * g2c
            self.accumulator_size = self.g1_chunk_size * g1u + extra_u + self.hash_size
            self.contribution_size = self.g1_chunk_size * g1c + extra_c + self.hash_size + self.public_key_size

    def get_length(self, compressed):
        return self.contribution_size - self.public_key_size if compressed else self.accumulator_size

    def split_offsets(self, compressed):
        """(offset, count, element_size) of [TauG1, TauG2, AlphaG1, BetaG1, BetaG2] (buffers.rs:293-341).
        Marlin: [TauG1, TauG2 (k+2), AlphaG1 (3+3k)] on chunk 0, TauG1 only elsewhere; BetaG1/BetaG2 empty."""
        g1, g2 = self.curve.g1.size(compressed), self.curve.g2.size(compressed)
        o = self.hash_size
        out = []
        if self.proving_system == MARLIN:
            k = self.total_size_in_log2
            first = self.chunk_index == 0
            counts = ((self.g1_chunk_size, g1), ((k + 2) if first else 0, g2), ((3 + 3 * k) if first else 0, g1), (0, g1), (0, g2))
        else:
            counts = ((self.g1_chunk_size, g1), (self.other_chunk_size, g2), (self.other_chunk_size, g1),
                      (self.other_chunk_size, g1), (1, g2))
        for cnt, sz in counts:
            out.append((o, cnt, sz))
            o += cnt * sz
        return out


def iter_chunk(params: Phase1Parameters):
    """phase1/src/helpers/buffers.rs:22-73 — list of (start, end) windows, overlapping by one element."""
    upper = params.powers_g1_length if params.proving_system == GROTH16 else params.powers_length
    if params.contribution_mode == CHUNKED_MODE:
        lo, hi = params.chunk_index * params.chunk_size, min((params.chunk_index + 1) * params.chunk_size, upper)
    else:
        lo, hi = 0, upper
    step = params.batch_size - 1
    out = []
    i = lo
    while i < hi:
        chunk = list(range(i, min(i + step, hi)))
        if len(chunk) >= 2:
            start, end = chunk[0], chunk[-1]
            out.append((start, end + 1 if end >= hi - 1 else end + 2))
        else:
            start = chunk[0]
            if start >= hi - 1:
                if hi == lo + 1:
                    out.append((start, start + 1))
            else:
                out.append((start, start + 2))
        i += step
    return out


def phase1_computation_marlin(params: Phase1Parameters, inp: bytes, compressed_in, compressed_out, check_in,
                              tau: int, alpha: int) -> bytearray:
    """Phase1::computation, Marlin branch (phase1/src/computation.rs:195-302)."""
    cv, r = params.curve, params.curve.r
    k, N = params.total_size_in_log2, params.powers_length
    out = bytearray(params.get_length(compressed_out))
    si, so = params.split_offsets(compressed_in), params.split_offsets(compressed_out)
    views = [inp[o:o + c * s] for (o, c, s) in si]

    def put(vec, start, data):
        o, _, s = so[vec]
        out[o + start * s:o + start * s + len(data)] = data

    if params.chunk_index == 0:
        dbp = [pow(tau, N - 1 - (1 << i) + 2, r) for i in range(k)]
        g2_inv = [pow(x, -1, r) for x in dbp]
        put(1, 2, apply_powers(cv.g2, views[1], compressed_in, check_in, compressed_out, 2, k + 2, g2_inv))
        g1_deg = []
        for f in dbp:
            g1_deg += [f, f * tau % r, f * tau * tau % r]
        put(2, 3, apply_powers(cv.g1, views[2], compressed_in, check_in, compressed_out, 3, 3 + 3 * k, g1_deg, alpha))
        put(2, 0, apply_powers(cv.g1, views[2], compressed_in, check_in, compressed_out, 0, 3,
                               generate_powers_of_tau(cv, tau, 0, 3), alpha))
        put(1, 0, apply_powers(cv.g2, views[1], compressed_in, check_in, compressed_out, 0, 2,
                               generate_powers_of_tau(cv, tau, 0, 2)))
    off = params.chunk_index * params.chunk_size if params.contribution_mode == CHUNKED_MODE else 0
    for (start, end) in iter_chunk(params):
        powers = generate_powers_of_tau(cv, tau, start, end)
        put(0, start - off, apply_powers(cv.g1, views[0], compressed_in, check_in, compressed_out, start - off, end - off, powers))
    return out


def phase1_computation(params: Phase1Parameters, inp: bytes, compressed_in, compressed_out, check_in,
                       tau: int, alpha: int, beta: int) -> bytearray:
    """Phase1::computation, Groth16 branch (phase1/src/computation.rs:16-193).

    Returns a buffer of params.get_length(compressed_out) bytes whose first 64 bytes (hash) are left zero:
    the caller writes the hash (phase1-cli/src/contribute.rs:92-96)."""
    if params.proving_system == MARLIN:
        return phase1_computation_marlin(params, inp, compressed_in, compressed_out, check_in, tau, alpha)
    cv = params.curve
    out = bytearray(params.get_length(compressed_out))
    si = params.split_offsets(compressed_in)
    so = params.split_offsets(compressed_out)
    views_in = [inp[o:o + c * s] for (o, c, s) in si]

    def put(vec, start, data):
        o, _, s = so[vec]
        out[o + start * s:o + start * s + len(data)] = data

    # beta_g2 (computation.rs:42-50)
    P = cv.g2.decode(views_in[4], compressed_in, check_in)
    put(4, 0, cv.g2.encode(cv.g2.mul(P, beta % cv.r), compressed_out))
    off = params.chunk_index * params.chunk_size if params.contribution_mode == CHUNKED_MODE else 0
    for (start, end) in iter_chunk(params):
        powers = generate_powers_of_tau(cv, tau, start, end)
        put(0, start - off, apply_powers(cv.g1, views_in[0], compressed_in, check_in, compressed_out,
                                         start - off, end - off, powers))
        if start < params.powers_length:
            if params.contribution_mode == CHUNKED_MODE:
                mx = min((params.chunk_index + 1) * params.chunk_size, params.powers_length)
            else:
                mx = params.powers_length
            e2 = mx if start + params.batch_size > mx else end
            s, e = start - off, e2 - off
            put(1, s, apply_powers(cv.g2, views_in[1], compressed_in, check_in, compressed_out, s, e, powers))
            put(2, s, apply_powers(cv.g1, views_in[2], compressed_in, check_in, compressed_out, s, e, powers, alpha))
            put(3, s, apply_powers(cv.g1, views_in[3], compressed_in, check_in, compressed_out, s, e, powers, beta))
    return out


def phase1_initialization(params: Phase1Parameters, compressed) -> bytearray:
    """Phase1::initialization (phase1/src/initialization.rs:12-57): every slot holds the generator."""
    cv = params.curve
    out = bytearray(params.get_length(compressed))
    for vec, (o, c, s) in enumerate(params.split_offsets(compressed)):
        g = cv.g2 if vec in (1, 4) else cv.g1
        out[o:o + c * s] = g.encode(g.gen, compressed) * c
    return out


# ----------------------------------------------------------------------------------------------
# prepare_phase2: Groth16Params::new / ::write
# ----------------------------------------------------------------------------------------------
# ark-ff 0.4 MontConfig (ark-ff-macros montgomery/mod.rs): TWO_ADIC_ROOT_OF_UNITY = GENERATOR^t with
# r - 1 = 2^s * t; GENERATOR = 22 for ark-bls12-377 Fr, 15 for ark-bls12-377 Fq (= ark-bw6-761 Fr) — the
# smallest primitive roots (checked against the small prime factors of r - 1 in tests/test_oracle_cpu.py).
FR_GENERATOR = {BLS12_377_R: 22, BLS12_377_Q: 15}


def two_adicity(r: int) -> int:
    s, t = 0, r - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    return s


def get_root_of_unity(r: int, n: int) -> int:
    """F::get_root_of_unity(n) (ark-ff fields/mod.rs, FftField): TWO_ADIC_ROOT_OF_UNITY squared down to order n."""
    s = two_adicity(r)
    log_n = n.bit_length() - 1
    if n != 1 << log_n or log_n > s:
        raise ValueError("no radix-2 domain of that size")
    omega = pow(FR_GENERATOR[r], (r - 1) >> s, r)
    for _ in range(log_n, s):
        omega = omega * omega % r
    return omega


def domain_size(phase2_size: int) -> int:
    """setup-utils/src/groth16_utils.rs:65-69: Radix2EvaluationDomain::new(n).size = n.next_power_of_two()."""
    m = 1
    while m < phase2_size:
        m *= 2
    return m


def group_ifft(group: Group, pts):
    """to_coeffs (setup-utils/src/groth16_utils.rs:44-53) by the definition of the inverse DFT:
    out_j = (1/n) sum_i w^(-i*j) P_i.  O(n^2) scalar multiplications — small n only."""
    n, r = len(pts), group.r
    w_inv = pow(get_root_of_unity(r, n), -1, r)
    n_inv = pow(n, -1, r)
    out = []
    for j in range(n):
        acc = None
        for i, P in enumerate(pts):
            acc = group.add(acc, group.mul(P, pow(w_inv, i * j, r) * n_inv % r))
        out.append(acc)
    return out


def group_ifft_fast(group: Group, pts):
    """Same transform by recursive decimation in frequency (an algorithm independent of the device's
    iterative decimation in time): n/2 log n scalar multiplications."""
    n, r = len(pts), group.r
    w_inv = pow(get_root_of_unity(r, n), -1, r)

    def rec(v, w):
        if len(v) == 1:
            return v
        h = len(v) // 2
        even = [group.add(v[i], v[i + h]) for i in range(h)]
        odd = [group.mul(group.add(v[i], group.neg(v[i + h])), pow(w, i, r)) for i in range(h)]
        e, o = rec(even, w * w % r), rec(odd, w * w % r)
        out = [None] * len(v)
        out[0::2], out[1::2] = e, o
        return out

    n_inv = pow(n, -1, r)
    return [group.mul(P, n_inv) for P in rec(list(pts), w_inv)]


def scalar_ifft(r: int, vals):
    """ark-poly domain.ifft on field elements (used to predict group_ifft of points with known discrete logs)."""
    n = len(vals)
    w_inv = pow(get_root_of_unity(r, n), -1, r)

    def rec(v, w):
        if len(v) == 1:
            return v
        h = len(v) // 2
        ww = w * w % r
        e = rec(v[0::2], ww)
        o = rec(v[1::2], ww)
        out = [0] * len(v)
        t = 1
        for i in range(h):
            x = t * o[i] % r
            out[i], out[i + h] = (e[i] + x) % r, (e[i] - x) % r
            t = t * w % r
        return out

    n_inv = pow(n, -1, r)
    return [x * n_inv % r for x in rec(list(vals), w_inv)]


def h_query_groth16(group: Group, powers, degree: int):
    """setup-utils/src/groth16_utils.rs:59-63: powers[i + degree] - powers[i] for i < degree - 1."""
    return [group.add(powers[i + degree], group.neg(powers[i])) for i in range(degree - 1)]


def groth16_params_new(params: "Phase1Parameters", acc: bytes, compressed_in, phase2_size: int, compressed_out,
                       check=NO, ifft=None):
    """Phase1::deserialize + Groth16Params::new + ::write (phase1/src/serialization.rs:23-41,
    setup-utils/src/groth16_utils.rs:81-168; phase2-cli/src/prepare_phase2.rs:16-70)."""
    cv = params.curve
    ifft = ifft or group_ifft_fast
    vecs = []
    for vec, (o, c, s) in enumerate(params.split_offsets(compressed_in)):
        g = cv.g2 if vec in (1, 4) else cv.g1
        vecs.append(g.read_batch(acc[o:o + c * s], compressed_in, check))
    tau_g1, tau_g2, alpha_g1, beta_g1, beta_g2 = vecs
    m = domain_size(phase2_size)
    if m > len(tau_g2):
        raise InvalidLength("phase2 domain larger than the powers of tau")  # slice panic in the reference
    out = bytearray()
    out += cv.g1.write_batch([alpha_g1[0], beta_g1[0]], compressed_out)
    out += cv.g2.write_batch([beta_g2[0]], compressed_out)
    out += cv.g1.write_batch(ifft(cv.g1, tau_g1[:m]), compressed_out)
    out += cv.g2.write_batch(ifft(cv.g2, tau_g2[:m]), compressed_out)
    out += cv.g1.write_batch(ifft(cv.g1, alpha_g1[:m]), compressed_out)
    out += cv.g1.write_batch(ifft(cv.g1, beta_g1[:m]), compressed_out)
    out += cv.g1.write_batch(h_query_groth16(cv.g1, tau_g1, m), compressed_out)
    return bytes(out)


# ----------------------------------------------------------------------------------------------
# same_ratio / check_same_ratio (setup-utils/src/helpers.rs:406-424) — SURVEY.md §8 A12, §8f rank 2
# ----------------------------------------------------------------------------------------------
# The reference compares E::pairing(g1.0, g2.1) with E::pairing(g1.1, g2.0) (arkworks' optimal ate pairings).
# Only the VERDICT crosses the boundary, and every non-degenerate bilinear pairing on G1 x G2 gives the same
# verdict, so the oracle uses the textbook reduced Tate pairing: affine Miller loop over r on E(Fq), lines
# evaluated at the untwisted G2 point in Fq^k = F[w]/(w^6 - xi), and the plain exponentiation by (q^k - 1)/r.
class ExtW6:
    """F[w]/(w^6 - xi) over F = Fp (BW6-761, k = 6) or Fp2 (BLS12-377, k = 12); elements are 6-lists."""

    def __init__(self, F, xi):
        self.F, self.xi = F, xi

    def one(self):
        F = self.F
        return [F.one] + [F.zero] * 5

    def mul(self, a, b):
        F = self.F
        acc = [F.zero] * 11
        for i, ai in enumerate(a):
            if F.is_zero(ai):
                continue
            for j, bj in enumerate(b):
                if not F.is_zero(bj):
                    acc[i + j] = F.add(acc[i + j], F.mul(ai, bj))
        return [F.add(acc[k], F.mul(self.xi, acc[k + 6])) if k < 5 else acc[k] for k in range(6)]

    def pow(self, a, e):
        r = self.one()
        for bit in bin(e)[2:]:
            r = self.mul(r, r)
            if bit == "1":
                r = self.mul(r, a)
        return r


def pairing_tower(curve: Curve):
    """(ExtW6, embedding degree, twist type) — ark-bls12-377: Fq12 = Fq2[v]/(v^3-u)[w]/(w^2-v), D-type twist;
    ark-bw6-761: Fq6 = Fq[v]/(v^3+4)[w]/(w^2-v), M-type twist."""
    F = curve.g2.F
    if curve.name == "bls12_377":
        return ExtW6(F, (0, 1)), 12, "D"
    return ExtW6(F, F.from_int(-4)), 6, "M"


def untwist(curve: Curve, Q):
    """G2 point on the twist -> (x, y) on E(Fq^k) as ExtW6 elements."""
    K, _, tw = pairing_tower(curve)
    F = K.F
    x, y = [F.zero] * 6, [F.zero] * 6
    if tw == "D":      # (x', y') -> (x' w^2, y' w^3)
        x[2], y[3] = Q[0], Q[1]
    else:              # (x', y') -> (x' / w^2, y' / w^3) = (x' w^4 / xi, y' w^3 / xi)
        xi_inv = F.inv(K.xi)
        x[4], y[3] = F.mul(Q[0], xi_inv), F.mul(Q[1], xi_inv)
    return x, y


def tate_miller(curve: Curve, P, Q):
    """f_{r,P}(Q) without vertical lines (denominator elimination); P in G1, Q in G2; identity -> 1."""
    K, _, _ = pairing_tower(curve)
    F, g1 = K.F, curve.g1
    if P is None or Q is None:
        return K.one()
    xq, yq = untwist(curve, Q)
    emb = (lambda v: (v, 0)) if F.degree == 2 else (lambda v: v)
    Fq = g1.F

    def line(T, lam):
        # y_Q - y_T - lam (x_Q - x_T)
        c = [F.sub(yq[k], F.mul(emb(lam), xq[k])) for k in range(6)]
        c[0] = F.add(c[0], emb(Fq.sub(Fq.mul(lam, T[0]), T[1])))
        return c

    f, T = K.one(), P
    bits = bin(curve.r)[3:]
    for n, bit in enumerate(bits):
        lam = Fq.mul(Fq.mul(3, Fq.sqr(T[0])), Fq.inv(Fq.mul(2, T[1])))
        f = K.mul(K.mul(f, f), line(T, lam))
        T = g1.add(T, T)
        if bit == "1" and n != len(bits) - 1:  # the last addition, T = -P, is a vertical line
            lam = Fq.mul(Fq.sub(T[1], P[1]), Fq.inv(Fq.sub(T[0], P[0])))
            f = K.mul(f, line(T, lam))
            T = g1.add(T, P)
    assert T == g1.neg(P)
    return f


def pairing(curve: Curve, P, Q):
    """Reduced Tate pairing t(P, Q) = f_{r,P}(Q)^((q^k - 1)/r)."""
    K, k, _ = pairing_tower(curve)
    q = curve.g1.F.p
    return K.pow(tate_miller(curve, P, Q), (q ** k - 1) // curve.r)


def same_ratio(curve: Curve, g1_pair, g2_pair) -> bool:
    """setup-utils/src/helpers.rs:406-408: e(g1.0, g2.1) == e(g1.1, g2.0), as e(g1.0, g2.1) * e(-g1.1, g2.0) == 1."""
    K, k, _ = pairing_tower(curve)
    q = curve.g1.F.p
    f = K.mul(tate_miller(curve, g1_pair[0], g2_pair[1]), tate_miller(curve, curve.g1.neg(g1_pair[1]), g2_pair[0]))
    return K.pow(f, (q ** k - 1) // curve.r) == K.one()


class InvalidRatio(SetupError):          # VerificationError::InvalidRatio
    pass


def check_same_ratio(curve: Curve, g1_pair, g2_pair):
    """setup-utils/src/helpers.rs:410-424."""
    if any(P is None for P in (*g1_pair, *g2_pair)):
        raise InvalidRatio("zero")
    if not same_ratio(curve, g1_pair, g2_pair):
        raise InvalidRatio("wrong pairing")


# ----------------------------------------------------------------------------------------------
# phase-2 QAP evaluation (phase2/src/polynomial.rs:11-94) — SURVEY.md §8f rank 4
# ----------------------------------------------------------------------------------------------
def dot_product(group: Group, row, coeffs):
    """phase2/src/polynomial.rs:80-94: sum of coeffs[ind].mul(coeff) over the (coeff, ind) entries of one row."""
    acc = None
    for c, ind in row:
        acc = group.add(acc, group.mul(coeffs[ind], c % group.r))
    return acc


def dot_product_vec(group: Group, rows, coeffs):
    """phase2/src/polynomial.rs:75-77 (+ the normalize_batch of eval, :41-45): one affine point per row."""
    return [dot_product(group, row, coeffs) for row in rows]


def process_matrix(xt, num_vars):
    """MPCParameters::process_matrix (phase2/src/parameters.rs:96-105): constraint-major -> variable-major."""
    out = [[] for _ in range(num_vars)]
    for constraint_num, vars_ in enumerate(xt):
        for coeff, var_index in vars_:
            out[var_index].append((coeff, constraint_num))
    return out


def qap_eval(curve: Curve, coeffs_g1, coeffs_g2, alpha_coeffs_g1, beta_coeffs_g1, at, bt, ct, num_inputs):
    """phase2/src/polynomial.rs:11-47 -> (a_g1, b_g1, b_g2, gamma_abc_g1, l)."""
    g1, g2 = curve.g1, curve.g2
    a_g1 = dot_product_vec(g1, at, coeffs_g1)
    b_g1 = dot_product_vec(g1, bt, coeffs_g1)
    b_g2 = dot_product_vec(g2, bt, coeffs_g2)
    ext = [g1.add(g1.add(dot_product(g1, a, beta_coeffs_g1), dot_product(g1, b, alpha_coeffs_g1)),
                  dot_product(g1, c, coeffs_g1)) for a, b, c in zip(at, bt, ct)]
    return a_g1, b_g1, b_g2, ext[:num_inputs], ext[num_inputs:]
