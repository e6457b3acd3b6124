// Canonical point (de)serialisation on the device.
//
// Byte format = ark-serialize 0.4 / ark-ec 0.4 short-Weierstrass:
// little-endian canonical field elements, flags in the two top bits of the last byte
// (0x80 = y is the lexicographically larger root, 0x40 = point at infinity); Fp2 = c0 || c1 with
// the flags on c1; compressed = x+flags, uncompressed = x || y+flags.
// Semantics of the validation modes follow `Deserializer::read_element`
// (setup-utils/src/io/read.rs:57-73) and `CheckForCorrectness -> Validate`
// (setup-utils/src/elements.rs:36-43); encoding follows `Serializer::write_element`
// (setup-utils/src/io/write.rs:30-35).
// All element sizes are multiples of 4 bytes, so elements are moved as little-endian u32 words,
// which are the limbs themselves.
#pragma once
#include "ec.cuh"
#include "glv.cuh"
#include "sqrt_fast.cuh"

namespace ss {

enum : int {
    ERR_OK = 0,
    ERR_INVALID_DATA = 1,       // SerializationError::InvalidData (non-canonical, non-square, failed Validate::Yes)
    ERR_UNEXPECTED_FLAGS = 2,   // SerializationError::UnexpectedFlags (both flag bits set)
    ERR_POINT_AT_INFINITY = 3,  // Error::PointAtInfinity
    ERR_INCORRECT_SUBGROUP = 4  // Error::IncorrectSubgroup
};
enum : int { CHECK_FULL = 0, CHECK_ONLY_NON_ZERO = 1, CHECK_ONLY_IN_GROUP = 2, CHECK_NO = 3 };

constexpr uint32_t FLAG_NEG_W = 0x80000000u;  // in the top word
constexpr uint32_t FLAG_INF_W = 0x40000000u;

template <class F>
struct FieldIO;

template <class P>
struct FieldIO<Fp<P>> {
    static constexpr int WORDS = P::N;
    // raw words -> Montgomery element.  ERR_UNEXPECTED_FLAGS when both flag bits are set (checked
    // first, as ark-ff deserialize_with_flags does), ERR_INVALID_DATA when the integer is >= p.
    SS_HD static int load(const uint32_t* w, bool has_flags, Fp<P>& out, uint32_t& flags) {
        Fp<P> raw;
#pragma unroll
        for (int i = 0; i < P::N; i++) raw.l[i] = w[i];
        flags = 0;
        if (has_flags) {
            flags = raw.l[P::N - 1] & 0xc0000000u;
            raw.l[P::N - 1] &= 0x3fffffffu;
            if (flags == 0xc0000000u) return ERR_UNEXPECTED_FLAGS;
        }
        if (raw_ge_mod<P>(raw.l)) return ERR_INVALID_DATA;
        out = fp_to_mont(raw);
        return ERR_OK;
    }
    SS_HD static void store(uint32_t* w, const Fp<P>& canon, uint32_t flags) {
#pragma unroll
        for (int i = 0; i < P::N - 1; i++) w[i] = canon.l[i];
        w[P::N - 1] = canon.l[P::N - 1] | flags;
    }
    // ---- byte-granular twins: field elements whose serialized size is not a multiple of 4 (MNT4/6-753: 95 bytes)
    static constexpr int BYTES = (P::BITS + 7) / 8;
    SS_HD static int stride(const uint32_t*) { return WORDS; }
    SS_HD static int stride(const uint8_t*) { return BYTES; }
    SS_HD static int load(const uint8_t* b, bool has_flags, Fp<P>& out, uint32_t& flags) {
        Fp<P> raw;
        for (int i = 0; i < P::N; i++) {
            uint32_t w = 0;
            for (int k = 0; k < 4; k++)
                if (4 * i + k < BYTES) w |= (uint32_t)b[4 * i + k] << (8 * k);
            raw.l[i] = w;
        }
        flags = 0;
        if (has_flags) {
            constexpr int TW = (BYTES - 1) / 4, TS = 8 * ((BYTES - 1) % 4);  // word / bit offset of the last byte
            flags = ((raw.l[TW] >> TS) & 0xc0u) << 24;                       // same positions as the word-aligned form
            raw.l[TW] &= ~(0xc0u << TS);
            if (flags == 0xc0000000u) return ERR_UNEXPECTED_FLAGS;
        }
        if (raw_ge_mod<P>(raw.l)) return ERR_INVALID_DATA;
        out = fp_to_mont(raw);
        return ERR_OK;
    }
    SS_HD static void store(uint8_t* b, const Fp<P>& canon, uint32_t flags) {
        for (int i = 0; i < BYTES; i++) b[i] = (uint8_t)(canon.l[i >> 2] >> (8 * (i & 3)));
        b[BYTES - 1] |= (uint8_t)(flags >> 24);
    }
    SS_HD static Fp<P> canonical(const Fp<P>& m) { return fp_from_mont(m); }
    // y > -y on canonical integers; `c` canonical
    SS_HD static bool is_negative(const Fp<P>& c) {
        if (c.is_zero()) return false;
        uint32_t neg[P::N];
        neg[0] = sub_cc(P::mod(0), c.l[0]);
#pragma unroll
        for (int i = 1; i < P::N - 1; i++) neg[i] = subc_cc(P::mod(i), c.l[i]);
        neg[P::N - 1] = subc(P::mod(P::N - 1), c.l[P::N - 1]);
        return raw_gt<P::N>(c.l, neg);
    }
};

template <class P>
struct FieldIO<Fp2<P>> {
    using B = FieldIO<Fp<P>>;
    static constexpr int WORDS = 2 * P::N;
    SS_HD static int load(const uint32_t* w, bool has_flags, Fp2<P>& out, uint32_t& flags) {
        uint32_t f0;
        flags = 0;
        int e = B::load(w, false, out.c0, f0);
        if (e) return e;
        return B::load(w + P::N, has_flags, out.c1, flags);
    }
    SS_HD static void store(uint32_t* w, const Fp2<P>& canon, uint32_t flags) {
        B::store(w, canon.c0, 0);
        B::store(w + P::N, canon.c1, flags);
    }
    static constexpr int BYTES = 2 * B::BYTES;
    SS_HD static int stride(const uint32_t*) { return WORDS; }
    SS_HD static int stride(const uint8_t*) { return BYTES; }
    SS_HD static int load(const uint8_t* b, bool has_flags, Fp2<P>& out, uint32_t& flags) {
        uint32_t f0;
        flags = 0;
        int e = B::load(b, false, out.c0, f0);
        if (e) return e;
        return B::load(b + B::BYTES, has_flags, out.c1, flags);
    }
    SS_HD static void store(uint8_t* b, const Fp2<P>& canon, uint32_t flags) {
        B::store(b, canon.c0, 0);
        B::store(b + B::BYTES, canon.c1, flags);
    }
    SS_HD static Fp2<P> canonical(const Fp2<P>& m) { return Fp2<P>{fp_from_mont(m.c0), fp_from_mont(m.c1)}; }
    // lexicographic, c1 first (ark-ff QuadExtField Ord)
    SS_HD static bool is_negative(const Fp2<P>& c) {
        if (!c.c1.is_zero()) return B::is_negative(c.c1);
        return B::is_negative(c.c0);
    }
};

template <class P>
struct FieldIO<Fp3<P>> {
    using B = FieldIO<Fp<P>>;
    static constexpr int WORDS = 3 * P::N;
    static constexpr int BYTES = 3 * B::BYTES;
    SS_HD static int stride(const uint32_t*) { return WORDS; }
    SS_HD static int stride(const uint8_t*) { return BYTES; }
    template <class T>
    SS_HD static int load(const T* w, bool has_flags, Fp3<P>& out, uint32_t& flags) {
        uint32_t f0;
        flags = 0;
        const int st = B::stride(w);
        int e = B::load(w, false, out.c0, f0);
        if (e) return e;
        if ((e = B::load(w + st, false, out.c1, f0))) return e;
        return B::load(w + 2 * st, has_flags, out.c2, flags);
    }
    template <class T>
    SS_HD static void store(T* w, const Fp3<P>& canon, uint32_t flags) {
        const int st = B::stride((const T*)w);
        B::store(w, canon.c0, 0);
        B::store(w + st, canon.c1, 0);
        B::store(w + 2 * st, canon.c2, flags);
    }
    SS_HD static Fp3<P> canonical(const Fp3<P>& m) { return Fp3<P>{fp_from_mont(m.c0), fp_from_mont(m.c1), fp_from_mont(m.c2)}; }
    // lexicographic, c2 first, then c1, then c0 (ark-ff CubicExtField Ord)
    SS_HD static bool is_negative(const Fp3<P>& c) {
        if (!c.c2.is_zero()) return B::is_negative(c.c2);
        if (!c.c1.is_zero()) return B::is_negative(c.c1);
        return B::is_negative(c.c0);
    }
};

// r * P == O  (the reference's explicit subgroup test, setup-utils/src/elements.rs:138-142)
template <class G>
SS_HD bool in_subgroup_rmul(const Affine<typename G::F>& p) {
    using F = typename G::F;
    Jac<F> t = jac_mul_bits<F>(p, [](int i) { return G::GP::order(i); }, G::GP::ORDER_BITS);
    return t.is_identity();
}

// ---- endomorphism subgroup tests for BLS12-377 -------------------------------------------------------
// The reference tests membership with a full r-multiplication (elements.rs:138-142,
// accumulator.rs:120-137).  For BLS12 curves the same predicate on curve points is
//     G1:  phi'(P) == -[u^2] P  (and not uP == P)      Scott, eprint 2021/1130 §6
//     G2:  psi(P)  ==  [u] P                             ibid. §4
// which is what ark-bls12-377 0.4.0 itself ships as `is_in_correct_subgroup_assuming_on_curve`
// (g1.rs / g2.rs).  u = 0x8508c00000000001 has 7 set bits: 63+6 group operations per [u] instead of
// 253+87, identical for every lane.  tests/test_oracle_cpu.py checks verdict equality with the
// r-multiplication on subgroup points, random curve points, pure cofactor torsion and mixed points.
// BW6-761 G1 and G2 have their own endomorphism tests below; the MNT groups keep the r-multiplication.
// -DSS_SUBGROUP_RMUL forces the reference algorithm everywhere.
constexpr unsigned long long kBls377U = 0x8508c00000000001ull;

template <class F>
SS_HD Jac<F> jac_mul_u(const Affine<F>& p) {
    Jac<F> acc{p.x, p.y, F::one()};
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((kBls377U >> i) & 1) acc = jac_madd(acc, p);
    }
    return acc;
}
template <class F>
SS_HD Jac<F> jac_mul_u(const Jac<F>& p) {
    Jac<F> acc = p;
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((kBls377U >> i) & 1) acc = jac_add(acc, p);
    }
    return acc;
}
// Jacobian J == affine A (both finite or both identity)
template <class F>
SS_HD bool jac_eq_affine(const Jac<F>& j, const F& ax, const F& ay) {
    if (j.Z.is_zero()) return false;
    F z2 = fp_sqr(j.Z);
    if (!(j.X == fp_mul(ax, z2))) return false;
    return j.Y == fp_mul(ay, fp_mul(z2, j.Z));
}

SS_HD bool in_subgroup_endo(const Affine<Fp<Bls377Fq>>& p, Bls377G1*) {
    using F = Fp<Bls377Fq>;
    if (p.inf) return true;
    Jac<F> up = jac_mul_u<F>(p);
    if (jac_eq_affine(up, p.x, p.y)) return false;  // uP == P, P != O
    Jac<F> u2p = jac_mul_u<F>(up);
    // phi'(x, y) = (beta' x, y) with beta' = beta^2 = -1 - beta;  -phi'(P) = (beta' x, -y)
    F beta;
#pragma unroll
    for (int i = 0; i < 12; i++) beta.l[i] = Bls377G1Glv::beta(i);
    F bx = fp_neg(fp_add(p.x, fp_mul(p.x, beta)));
    return jac_eq_affine(u2p, bx, fp_neg(p.y));
}
SS_HD bool in_subgroup_endo(const Affine<Fp2<Bls377Fq>>& p, Bls377G2*) {
    using F = Fp2<Bls377Fq>;
    if (p.inf) return true;
    Jac<F> up = jac_mul_u<F>(p);
    F x = p.x, y = p.y;
    Endo<Bls377G2>::apply(1, x, y);  // psi(P)
    return jac_eq_affine(up, x, y);
}
// ---- endomorphism subgroup tests for BW6-761 ------------------------------------------------------------------------
// phi(x, y) = (beta x, y) acts on the order-r subgroup as lambda, lambda^2 + lambda + 1 = 0 (mod r).  For a lattice
// vector (a, b), a + b lambda = 0 (mod r), of norm a^2 - ab + b^2 EXACTLY r (tools/gen_constants.py, Bw6G{1,2}SubgroupVec;
// the reduced GLV basis consists of such vectors), psi = a + b phi is an endomorphism of degree r that kills the
// subgroup, hence ker psi IS the subgroup: psi(P) = O accepts exactly the order-r points, the same predicate as the
// reference's r-multiplication (elements.rs:138-142), at 189 doublings + ~95 additions (joint sparse form over P, phi(P),
// P + phi(P), P - phi(P); the digits are constants, so control flow is uniform) instead of 376 + 134.
// Round 1 used the sparse vector (u + 1, u^3 - u^2 + 1) for G1, whose norm is 3r: sound there only because the extra
// 3-torsion kernel is irrational when q = 3 (mod 4), and unusable for G2 (on y^2 = x^3 + 4 the points (0, +-2) are
// rational and in that kernel).  The norm-r vector needs no such argument and measures faster on both groups
// (k_subgroup<bw6 g2> 39.4 -> 20.4 ms per 2^16 elements; profiles/r02_ab_variants.md).
template <class Glv, class V>
SS_HD bool in_subgroup_norm_r(const Affine<Fp<Bw6Fq>>& p) {
    using F = Fp<Bw6Fq>;
    if (p.inf) return true;
    F beta;
#pragma unroll
    for (int i = 0; i < 24; i++) beta.l[i] = Glv::beta(i);
    const Affine<F> phip{fp_mul(p.x, beta), p.y, false};
    const Jac<F> jp{p.x, p.y, F::one()};
    const Jac<F> sum = jac_madd(jp, phip);               // P + phi(P)
    const Jac<F> dif = jac_madd(jp, affine_neg(phip));   // P - phi(P)
    Jac<F> acc = Jac<F>::identity();
#pragma unroll 1
    for (int i = 0; i < V::LEN; i++) {
        acc = jac_dbl(acc);
        const int d = (int)V::digit(i), da = d / 3 - 1, db = d % 3 - 1;
        if (da != 0 && db == 0) {
            acc = jac_madd(acc, da > 0 ? p : affine_neg(p));
        } else if (da == 0 && db != 0) {
            acc = jac_madd(acc, db > 0 ? phip : affine_neg(phip));
        } else if (da != 0) {  // both: +-(P + phi P) when the signs agree, +-(P - phi P) otherwise
            Jac<F> t = da == db ? sum : dif;
            if (da < 0) t.Y = fp_neg(t.Y);
            acc = jac_add(acc, t);
        }
    }
    return acc.is_identity();
}
SS_HD bool in_subgroup_endo(const Affine<Fp<Bw6Fq>>& p, Bw6G1*) { return in_subgroup_norm_r<Bw6G1Glv, Bw6G1SubgroupVec>(p); }
SS_HD bool in_subgroup_endo(const Affine<Fp<Bw6Fq>>& p, Bw6G2*) { return in_subgroup_norm_r<Bw6G2Glv, Bw6G2SubgroupVec>(p); }
template <class G>
SS_HD bool in_subgroup_endo(const Affine<typename G::F>& p, G*) {
    return in_subgroup_rmul<G>(p);
}

template <class G>
SS_HD bool in_subgroup(const Affine<typename G::F>& p) {
#if defined(SS_SUBGROUP_RMUL)
    return in_subgroup_rmul<G>(p);
#else
    return in_subgroup_endo(p, (G*)nullptr);
#endif
}

// Decode one element.  Returns ERR_*; `out` is valid when ERR_OK.
template <class G, class T>
SS_HD int decode_point(const T* w, bool compressed, int check, Affine<typename G::F>& out) {
    using F = typename G::F;
    using IO = FieldIO<F>;
    uint32_t flags;
    out.inf = false;
    int e;
    if (compressed) {
        if ((e = IO::load(w, true, out.x, flags)) != ERR_OK) return e;
        if (flags & FLAG_INF_W) {
            out.inf = true;
        } else {
            F rhs = curve_rhs(out.x, G::b());
            F y;
            if (!fp_sqrt_any(rhs, y)) return ERR_INVALID_DATA;
            bool neg = IO::is_negative(IO::canonical(y));
            bool want_neg = (flags & FLAG_NEG_W) != 0;
            out.y = (neg == want_neg) ? y : fp_neg(y);
        }
    } else {
        uint32_t fx;
        if ((e = IO::load(w, false, out.x, fx)) != ERR_OK) return e;
        if ((e = IO::load(w + IO::stride(w), true, out.y, flags)) != ERR_OK) return e;
        if (flags & FLAG_INF_W) out.inf = true;
    }
    if (out.inf) {
        out.x = F::zero();
        out.y = F::zero();
    } else if (check == CHECK_FULL || check == CHECK_ONLY_IN_GROUP) {
        if (!on_curve(out, G::b())) return ERR_INVALID_DATA;
        if (!in_subgroup<G>(out)) return ERR_INVALID_DATA;
    }
    if ((check == CHECK_FULL || check == CHECK_ONLY_NON_ZERO) && out.inf) return ERR_POINT_AT_INFINITY;
    return ERR_OK;
}

// Encode one affine element (Montgomery coordinates) to `w`.
template <class G, class T>
SS_HD void encode_point(T* w, bool compressed, const Affine<typename G::F>& p) {
    using F = typename G::F;
    using IO = FieldIO<F>;
    if (p.inf) {
        const int words = (compressed ? 1 : 2) * IO::stride((const T*)w);
        for (int i = 0; i < words - 1; i++) w[i] = 0;
        w[words - 1] = (T)(FLAG_INF_W >> (8 * (4 - (int)sizeof(T))));  // 0x40 in the last byte
        return;
    }
    F xc = IO::canonical(p.x);
    F yc = IO::canonical(p.y);
    uint32_t flags = IO::is_negative(yc) ? FLAG_NEG_W : 0u;
    if (compressed) {
        IO::store(w, xc, flags);
    } else {
        IO::store(w, xc, 0);
        IO::store(w + IO::stride((const T*)w), yc, flags);
    }
}

}  // namespace ss
