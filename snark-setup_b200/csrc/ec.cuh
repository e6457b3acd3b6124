// Short-Weierstrass (a = 0) group arithmetic in Jacobian coordinates, generic over the
// coordinate field (Fp for BLS12-377 G1 and both BW6-761 groups, Fp2 for BLS12-377 G2).
//
// Device-side replacement for ark-ec 0.4 `short_weierstrass::{Affine, Projective}`:
// `AffineRepr::mul` / `mul_bigint` (setup-utils/src/helpers.rs:64,104; elements.rs:138,142;
// phase1/src/helpers/accumulator.rs:121,131).  Results are exact group elements, so any correct
// formula set yields the bytes the reference produces once normalised and canonically encoded
// (SURVEY.md Appendix A.3).
#pragma once
#include "constants_gen.cuh"
#include "fp2.cuh"
#include "fp2l.cuh"

namespace ss {

template <class F>
struct Affine {
    F x, y;
    bool inf;
};

template <class F>
struct Jac {
    F X, Y, Z;  // Z == 0 <=> identity
    SS_HD static Jac identity() { return Jac{F::one(), F::one(), F::zero()}; }
    SS_HD bool is_identity() const { return Z.is_zero(); }
};

// ---- group descriptors -----------------------------------------------------------------------
struct Bls377G1 {
    using F = Fp<Bls377Fq>;
    using Fr = Fp<Bls377Fr>;
    using GP = Bls377G1Params;
    static constexpr int COORD_LIMBS = 12;
    static constexpr int USIZE = 96, CSIZE = 48;
    static constexpr int SMUL_MINB = 4;  // k_scalar_mul blocks/SM: 128 regs, 16 warps/SM measured best (A/B in profiles/)
    static constexpr bool HAS_ENDO = true, BYTE_IO = false;
    static const char* name() { return "bls12_377.g1"; }
    SS_HD static F b() {
        F r;
#pragma unroll
        for (int i = 0; i < 12; i++) r.l[i] = GP::b(i);
        return r;
    }
    SS_HD static Affine<F> generator() {
        Affine<F> g;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            g.x.l[i] = GP::gx(i);
            g.y.l[i] = GP::gy(i);
        }
        g.inf = false;
        return g;
    }
};

struct Bls377G2 {
    using F = Fp2<Bls377Fq>;
    using Fr = Fp<Bls377Fr>;
    using GP = Bls377G2Params;
    static constexpr int USIZE = 192, CSIZE = 96;
    static constexpr int SMUL_MINB = 1;  // Fq2 needs the full 255 registers
    static constexpr bool HAS_ENDO = true, BYTE_IO = false;
    static const char* name() { return "bls12_377.g2"; }
    SS_HD static F b() {
        F r;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            r.c0.l[i] = GP::b_c0(i);
            r.c1.l[i] = GP::b_c1(i);
        }
        return r;
    }
    SS_HD static Affine<F> generator() {
        Affine<F> g;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            g.x.c0.l[i] = GP::gx_c0(i);
            g.x.c1.l[i] = GP::gx_c1(i);
            g.y.c0.l[i] = GP::gy_c0(i);
            g.y.c1.l[i] = GP::gy_c1(i);
        }
        g.inf = false;
        return g;
    }
};

#if defined(__CUDACC__)
// Lane-split twin of Bls377G2 (fp2l.cuh): the same group, one element per LANE PAIR.  Only the kernels whose time is
// group arithmetic use it (scalar multiplication, subgroup test, bucket accumulation); serialisation, normalisation
// and the reductions keep the per-thread form and the two share every HBM layout.
struct Bls377G2L {
    using F = Fp2L<Bls377Fq>;
    using Fr = Fp<Bls377Fr>;
    using GP = Bls377G2Params;
    using Wide = Bls377G2;
    static constexpr int USIZE = 192, CSIZE = 96;
#ifndef SS_G2L_MINB
#define SS_G2L_MINB 3
#endif
    static constexpr int SMUL_MINB = SS_G2L_MINB;  // 128-thread blocks per SM (168 registers at 3)
    static constexpr bool HAS_ENDO = true, BYTE_IO = false;
    static const char* name() { return "bls12_377.g2"; }
    SS_HD static F b() {
        F r;
        const int odd = lane_odd();
#pragma unroll
        for (int i = 0; i < 12; i++) r.h.l[i] = odd ? GP::b_c1(i) : GP::b_c0(i);
        return r;
    }
};
template <class G>
struct PairTwin {
    using type = void;
};
template <>
struct PairTwin<Bls377G2> {
    using type = Bls377G2L;
};
#endif

template <class GPx>
struct Bw6Group {
    using F = Fp<Bw6Fq>;
    using Fr = Fp<Bls377Fq>;  // BW6-761's scalar field is BLS12-377's base field
    using GP = GPx;
    static constexpr int USIZE = 192, CSIZE = 96;
    static constexpr int SMUL_MINB = 1;
    static constexpr bool HAS_ENDO = true, BYTE_IO = false;
    static const char* name() { return GP::b(0) == Bw6G1Params::b(0) ? "bw6_761.g1" : "bw6_761.g2"; }
    SS_HD static F b() {
        F r;
#pragma unroll
        for (int i = 0; i < 24; i++) r.l[i] = GP::b(i);
        return r;
    }
    SS_HD static Affine<F> generator() {
        Affine<F> g;
#pragma unroll
        for (int i = 0; i < 24; i++) {
            g.x.l[i] = GP::gx(i);
            g.y.l[i] = GP::gy(i);
        }
        g.inf = false;
        return g;
    }
};
using Bw6G1 = Bw6Group<Bw6G1Params>;
using Bw6G2 = Bw6Group<Bw6G2Params>;

// ---- MNT4-753 / MNT6-753 (setup-utils/src/converters.rs:18-45): y^2 = x^3 + a x + b with a != 0 ------------------------
// G1 has prime order (cofactor 1); G2 lives on the twist over Fq2 (MNT4) / Fq3 (MNT6).  Field elements serialize to 95
// bytes, so elements are NOT word-aligned in the reference's packed layout (BYTE_IO); no efficient endomorphism is
// used (HAS_ENDO = false: plain double-and-add).  Constants: tools/gen_constants.py from oracle/pyref.py, where they are
// verified (curve orders, twist orders, non-residues).  The arkworks G2 GENERATOR constants are not known here.
template <class GPx, class FqP, class FrP, int WHICH>
struct MntG1 {
    using F = Fp<FqP>;
    using Fr = Fp<FrP>;
    using GP = GPx;
    static constexpr int USIZE = 190, CSIZE = 95;
    static constexpr int SMUL_MINB = 1;
    static constexpr bool HAS_ENDO = false, BYTE_IO = true, HAS_GENERATOR = true;
    static const char* name() { return WHICH == 4 ? "mnt4_753.g1" : "mnt6_753.g1"; }
    SS_HD static F a() { F r; for (int i = 0; i < 24; i++) r.l[i] = GP::a(i); return r; }
    SS_HD static F b() { F r; for (int i = 0; i < 24; i++) r.l[i] = GP::b(i); return r; }
    SS_HD static Affine<F> generator() {
        Affine<F> g;
        for (int i = 0; i < 24; i++) {
            g.x.l[i] = GP::gx(i);
            g.y.l[i] = GP::gy(i);
        }
        g.inf = false;
        return g;
    }
};
using Mnt4G1 = MntG1<Mnt4G1Params, Mnt753Q, Mnt753R, 4>;
using Mnt6G1 = MntG1<Mnt6G1Params, Mnt753R, Mnt753Q, 6>;

struct Mnt4G2 {
    using F = Fp2<Mnt753Q>;
    using Fr = Fp<Mnt753R>;
    using GP = Mnt4G2Params;
    static constexpr int USIZE = 380, CSIZE = 190;
    static constexpr int SMUL_MINB = 1;
    static constexpr bool HAS_ENDO = false, BYTE_IO = true, HAS_GENERATOR = false;
    static const char* name() { return "mnt4_753.g2"; }
    SS_HD static F a() { F r; for (int i = 0; i < 24; i++) { r.c0.l[i] = GP::a_c0(i); r.c1.l[i] = GP::a_c1(i); } return r; }
    SS_HD static F b() { F r; for (int i = 0; i < 24; i++) { r.c0.l[i] = GP::b_c0(i); r.c1.l[i] = GP::b_c1(i); } return r; }
    SS_HD static Affine<F> generator() { return Affine<F>{F::zero(), F::zero(), true}; }
};
struct Mnt6G2 {
    using F = Fp3<Mnt753R>;
    using Fr = Fp<Mnt753Q>;
    using GP = Mnt6G2Params;
    static constexpr int USIZE = 570, CSIZE = 285;
    static constexpr int SMUL_MINB = 1;
    static constexpr bool HAS_ENDO = false, BYTE_IO = true, HAS_GENERATOR = false;
    static const char* name() { return "mnt6_753.g2"; }
    SS_HD static F a() {
        F r;
        for (int i = 0; i < 24; i++) { r.c0.l[i] = GP::a_c0(i); r.c1.l[i] = GP::a_c1(i); r.c2.l[i] = GP::a_c2(i); }
        return r;
    }
    SS_HD static F b() {
        F r;
        for (int i = 0; i < 24; i++) { r.c0.l[i] = GP::b_c0(i); r.c1.l[i] = GP::b_c1(i); r.c2.l[i] = GP::b_c2(i); }
        return r;
    }
    SS_HD static Affine<F> generator() { return Affine<F>{F::zero(), F::zero(), true}; }
};

// The curve coefficient a, keyed by the COORDINATE FIELD type (every field below belongs to exactly one group, so the
// formula templates — which only know F — can ask for it): zero for the BLS12-377 / BW6-761 fields.
template <class F>
struct CurveCoeffA {
    static constexpr bool ZERO = true;
    SS_HD static F a() { return F::zero(); }
};
template <>
struct CurveCoeffA<Fp<Mnt753Q>> {
    static constexpr bool ZERO = false;
    SS_HD static Fp<Mnt753Q> a() { return Mnt4G1::a(); }
};
template <>
struct CurveCoeffA<Fp<Mnt753R>> {
    static constexpr bool ZERO = false;
    SS_HD static Fp<Mnt753R> a() { return Mnt6G1::a(); }
};
template <>
struct CurveCoeffA<Fp2<Mnt753Q>> {
    static constexpr bool ZERO = false;
    SS_HD static Fp2<Mnt753Q> a() { return Mnt4G2::a(); }
};
template <>
struct CurveCoeffA<Fp3<Mnt753R>> {
    static constexpr bool ZERO = false;
    SS_HD static Fp3<Mnt753R> a() { return Mnt6G2::a(); }
};

// ---- formulas ---------------------------------------------------------------------------------
// Group operations come in three flavours: *_inl (the formula), *_call (one out-of-line copy per
// field, used for the wide fields where inlining every Fp2 / 24-limb add-sub chain makes the NVVM
// front end take >10 minutes and kernels of hundreds of KB) and the dispatching jac_dbl / jac_madd /
// jac_add.  jac_dbl_cold is the always-out-of-line doubling used by the exceptional P+P branches.
template <class F>
SS_HD Jac<F> jac_dbl_cold(const Jac<F>& p);
// dbl-2009-l (a = 0): 2M + 5S
template <class F>
SS_HD Jac<F> jac_dbl_inl(const Jac<F>& p) {
    if (p.Z.is_zero()) return p;
    if constexpr (!CurveCoeffA<F>::ZERO) {
        // dbl-2007-bl (general a): 1M + 8S + 1 multiplication by a
        F XX = fp_sqr(p.X), YY = fp_sqr(p.Y), ZZ = fp_sqr(p.Z);
        F YYYY = fp_sqr(YY);
        F S = fp_dbl(fp_sub(fp_sub(fp_sqr(fp_add(p.X, YY)), XX), YYYY));
        F M = fp_add(fp_add(fp_dbl(XX), XX), fp_mul(CurveCoeffA<F>::a(), fp_sqr(ZZ)));
        Jac<F> r;
        r.X = fp_sub(fp_sqr(M), fp_dbl(S));
        r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_dbl(fp_dbl(fp_dbl(YYYY))));
        r.Z = fp_sub(fp_sub(fp_sqr(fp_add(p.Y, p.Z)), YY), ZZ);
        return r;
    }
    F A = fp_sqr(p.X);
    F B = fp_sqr(p.Y);
    F C = fp_sqr(B);
    F t = fp_sqr(fp_add(p.X, B));
    F D = fp_dbl(fp_sub(fp_sub(t, A), C));
    F E = fp_add(fp_dbl(A), A);
    F Fq = fp_sqr(E);
    Jac<F> r;
    r.X = fp_sub(Fq, fp_dbl(D));
    F C8 = fp_dbl(fp_dbl(fp_dbl(C)));
    r.Z = fp_dbl(fp_mul(p.Y, p.Z));
    r.Y = fp_sub(fp_mul(E, fp_sub(D, r.X)), C8);
    return r;
}

// madd-2007-bl: Jacobian + affine, 7M + 4S, with the exceptional cases handled
template <class F>
SS_HD Jac<F> jac_madd_inl(const Jac<F>& p, const Affine<F>& q) {
    if (q.inf) return p;
    if (p.Z.is_zero()) return Jac<F>{q.x, q.y, F::one()};
    F Z1Z1 = fp_sqr(p.Z);
    F U2 = fp_mul(q.x, Z1Z1);
    F S2 = fp_mul(fp_mul(q.y, p.Z), Z1Z1);
    F H = fp_sub(U2, p.X);
    F rr = fp_sub(S2, p.Y);
    if (H.is_zero()) {
        if (rr.is_zero()) return jac_dbl_cold(p);
        return Jac<F>::identity();
    }
    rr = fp_dbl(rr);
    F HH = fp_sqr(H);
    F I = fp_dbl(fp_dbl(HH));
    F J = fp_mul(H, I);
    F V = fp_mul(p.X, I);
    Jac<F> r;
    r.X = fp_sub(fp_sub(fp_sqr(rr), J), fp_dbl(V));
    r.Y = fp_sub(fp_mul(rr, fp_sub(V, r.X)), fp_dbl(fp_mul(p.Y, J)));
    r.Z = fp_sub(fp_sub(fp_sqr(fp_add(p.Z, H)), Z1Z1), HH);
    return r;
}

// add-2007-bl: Jacobian + Jacobian, 11M + 5S
template <class F>
SS_HD Jac<F> jac_add_inl(const Jac<F>& p, const Jac<F>& q) {
    if (p.Z.is_zero()) return q;
    if (q.Z.is_zero()) return p;
    F Z1Z1 = fp_sqr(p.Z);
    F Z2Z2 = fp_sqr(q.Z);
    F U1 = fp_mul(p.X, Z2Z2);
    F U2 = fp_mul(q.X, Z1Z1);
    F S1 = fp_mul(fp_mul(p.Y, q.Z), Z2Z2);
    F S2 = fp_mul(fp_mul(q.Y, p.Z), Z1Z1);
    F H = fp_sub(U2, U1);
    F rr = fp_sub(S2, S1);
    if (H.is_zero()) {
        if (rr.is_zero()) return jac_dbl_cold(p);
        return Jac<F>::identity();
    }
    rr = fp_dbl(rr);
    F I = fp_sqr(fp_dbl(H));
    F J = fp_mul(H, I);
    F V = fp_mul(U1, I);
    Jac<F> r;
    r.X = fp_sub(fp_sub(fp_sqr(rr), J), fp_dbl(V));
    r.Y = fp_sub(fp_mul(rr, fp_sub(V, r.X)), fp_dbl(fp_mul(S1, J)));
    r.Z = fp_mul(fp_sub(fp_sub(fp_sqr(fp_add(p.Z, q.Z)), Z1Z1), Z2Z2), H);
    return r;
}

#if defined(__CUDACC__)
template <class F>
__device__ __noinline__ Jac<F> jac_dbl_call(Jac<F> p) {
    return jac_dbl_inl(p);
}
template <class F>
__device__ __noinline__ Jac<F> jac_madd_call(Jac<F> p, Affine<F> q) {
    return jac_madd_inl(p, q);
}
template <class F>
__device__ __noinline__ Jac<F> jac_add_call(Jac<F> p, Jac<F> q) {
    return jac_add_inl(p, q);
}
#endif

template <class F>
SS_HD Jac<F> jac_dbl_cold(const Jac<F>& p) {
#if defined(__CUDA_ARCH__)
    return jac_dbl_call<F>(p);
#else
    return jac_dbl_inl(p);
#endif
}
template <class F>
SS_HD Jac<F> jac_dbl(const Jac<F>& p) {
#if defined(__CUDA_ARCH__) && !defined(SS_GROUP_INLINE)
    if constexpr (F::CALL_GROUP_OPS) return jac_dbl_call<F>(p);
#endif
    return jac_dbl_inl(p);
}
template <class F>
SS_HD Jac<F> jac_madd(const Jac<F>& p, const Affine<F>& q) {
#if defined(__CUDA_ARCH__) && !defined(SS_GROUP_INLINE)
    if constexpr (F::CALL_GROUP_OPS) return jac_madd_call<F>(p, q);
#endif
    return jac_madd_inl(p, q);
}
template <class F>
SS_HD Jac<F> jac_add(const Jac<F>& p, const Jac<F>& q) {
#if defined(__CUDA_ARCH__) && !defined(SS_GROUP_INLINE)
    if constexpr (F::CALL_GROUP_OPS) return jac_add_call<F>(p, q);
#endif
    return jac_add_inl(p, q);
}

#if defined(__CUDACC__)
// ---- group law on lane-split elements (fp2l.cuh): converged-warp versions ---------------------------------------
// Same formulas; the exceptional cases are resolved by SELECTION after the common computation, and the one that needs
// more arithmetic (P + P) is entered by the whole warp behind __any_sync, so that every lane executes the same
// sequence of exchanges.  More specialised than the generic templates above, so overload resolution picks them for
// Jac<Fp2L<P>> wherever the generic code calls jac_dbl / jac_madd / jac_add (or their _inl / _cold forms).
template <class P>
SS_D Fp2L<P> f2l_select(bool c, const Fp2L<P>& a, const Fp2L<P>& b) {
    Fp2L<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.h.l[i] = c ? a.h.l[i] : b.h.l[i];
    return r;
}
template <class P>
SS_D Jac<Fp2L<P>> jac_select(bool c, const Jac<Fp2L<P>>& a, const Jac<Fp2L<P>>& b) {
    return Jac<Fp2L<P>>{f2l_select(c, a.X, b.X), f2l_select(c, a.Y, b.Y), f2l_select(c, a.Z, b.Z)};
}

// dbl-2009-l needs no case distinction: Z = 0 (or Y = 0) gives Z3 = 2 Y Z = 0, i.e. the identity
template <class P>
SS_D Jac<Fp2L<P>> jac_dbl_inl(const Jac<Fp2L<P>>& p) {
    using F = Fp2L<P>;
    F A = fp_sqr(p.X);
    F B = fp_sqr(p.Y);
    F C = fp_sqr(B);
    F t = fp_sqr(fp_add(p.X, B));
    F D = fp_dbl(fp_sub(fp_sub(t, A), C));
    F E = fp_add(fp_dbl(A), A);
    F Fq = fp_sqr(E);
    Jac<F> r;
    r.X = fp_sub(Fq, fp_dbl(D));
    F C8 = fp_dbl(fp_dbl(fp_dbl(C)));
    r.Z = fp_dbl(fp_mul(p.Y, p.Z));
    r.Y = fp_sub(fp_mul(E, fp_sub(D, r.X)), C8);
    return r;
}
template <class P>
SS_D Jac<Fp2L<P>> jac_dbl_cold(const Jac<Fp2L<P>>& p) { return jac_dbl_inl(p); }
template <class P>
SS_D Jac<Fp2L<P>> jac_dbl(const Jac<Fp2L<P>>& p) { return jac_dbl_inl(p); }

template <class P>
SS_D Jac<Fp2L<P>> jac_madd_inl(const Jac<Fp2L<P>>& p, const Affine<Fp2L<P>>& q) {
    using F = Fp2L<P>;
    F Z1Z1 = fp_sqr(p.Z);
    F U2 = fp_mul(q.x, Z1Z1);
    F S2 = fp_mul(fp_mul(q.y, p.Z), Z1Z1);
    F H = fp_sub(U2, p.X);
    F rr = fp_sub(S2, p.Y);
    const bool pinf = p.Z.is_zero(), hz = H.is_zero(), rz = rr.is_zero();
    const bool exc = !q.inf && !pinf && hz;  // P = +-Q
    rr = fp_dbl(rr);
    F HH = fp_sqr(H);
    F I = fp_dbl(fp_dbl(HH));
    F J = fp_mul(H, I);
    F V = fp_mul(p.X, I);
    Jac<F> r;
    r.X = fp_sub(fp_sub(fp_sqr(rr), J), fp_dbl(V));
    r.Y = fp_sub(fp_mul(rr, fp_sub(V, r.X)), fp_dbl(fp_mul(p.Y, J)));
    r.Z = fp_sub(fp_sub(fp_sqr(fp_add(p.Z, H)), Z1Z1), HH);
    if (warp_any(exc && rz)) r = jac_select(exc && rz, jac_dbl_inl(p), r);
    r = jac_select(exc && !rz, Jac<F>::identity(), r);
    r = jac_select(pinf, Jac<F>{q.x, q.y, F::one()}, r);
    return jac_select(q.inf, p, r);
}
template <class P>
SS_D Jac<Fp2L<P>> jac_madd(const Jac<Fp2L<P>>& p, const Affine<Fp2L<P>>& q) { return jac_madd_inl(p, q); }

template <class P>
SS_D Jac<Fp2L<P>> jac_add_inl(const Jac<Fp2L<P>>& p, const Jac<Fp2L<P>>& q) {
    using F = Fp2L<P>;
    F Z1Z1 = fp_sqr(p.Z);
    F Z2Z2 = fp_sqr(q.Z);
    F U1 = fp_mul(p.X, Z2Z2);
    F U2 = fp_mul(q.X, Z1Z1);
    F S1 = fp_mul(fp_mul(p.Y, q.Z), Z2Z2);
    F S2 = fp_mul(fp_mul(q.Y, p.Z), Z1Z1);
    F H = fp_sub(U2, U1);
    F rr = fp_sub(S2, S1);
    const bool pinf = p.Z.is_zero(), qinf = q.Z.is_zero(), hz = H.is_zero(), rz = rr.is_zero();
    const bool exc = !pinf && !qinf && hz;
    rr = fp_dbl(rr);
    F I = fp_sqr(fp_dbl(H));
    F J = fp_mul(H, I);
    F V = fp_mul(U1, I);
    Jac<F> r;
    r.X = fp_sub(fp_sub(fp_sqr(rr), J), fp_dbl(V));
    r.Y = fp_sub(fp_mul(rr, fp_sub(V, r.X)), fp_dbl(fp_mul(S1, J)));
    r.Z = fp_mul(fp_sub(fp_sub(fp_sqr(fp_add(p.Z, q.Z)), Z1Z1), Z2Z2), H);
    if (warp_any(exc && rz)) r = jac_select(exc && rz, jac_dbl_inl(p), r);
    r = jac_select(exc && !rz, Jac<F>::identity(), r);
    r = jac_select(qinf, p, r);
    return jac_select(pinf, q, r);
}
template <class P>
SS_D Jac<Fp2L<P>> jac_add(const Jac<Fp2L<P>>& p, const Jac<Fp2L<P>>& q) { return jac_add_inl(p, q); }

// MSB-first double-and-add with the addition computed by every lane and SELECTED by the bit (per-lane scalars must not
// diverge the warp); `base.inf` is resolved at the end.  The slow paths only: strict mode, tiny-order bases, r * P.
template <class P, class LimbFn>
SS_D Jac<Fp2L<P>> jac_mul_bits_pair(const Affine<Fp2L<P>>& base, LimbFn limb, int nbits) {
    using F = Fp2L<P>;
    Affine<F> b = base;
    b.inf = false;
    Jac<F> acc = Jac<F>::identity();
#pragma unroll 1
    for (int i = nbits - 1; i >= 0; i--) {
        acc = jac_dbl_inl(acc);
        const bool bit = ((limb(i >> 5) >> (i & 31)) & 1) != 0;
        acc = jac_select(bit, jac_madd_inl(acc, b), acc);
    }
    return jac_select(base.inf, Jac<F>::identity(), acc);
}

// Jacobian J == affine (ax, ay), all comparisons evaluated by every lane
template <class P>
SS_D bool jac_eq_affine_pair(const Jac<Fp2L<P>>& j, const Fp2L<P>& ax, const Fp2L<P>& ay) {
    using F = Fp2L<P>;
    const bool zz = j.Z.is_zero();
    F z2 = fp_sqr(j.Z);
    const bool ex = j.X == fp_mul(ax, z2);
    const bool ey = j.Y == fp_mul(ay, fp_mul(z2, j.Z));
    return !zz && ex && ey;
}
#endif  // __CUDACC__

template <class F>
SS_HD Affine<F> affine_neg(const Affine<F>& p) {
    return Affine<F>{p.x, fp_neg(p.y), p.inf};
}

// x^3 + a x + b
template <class F>
SS_HD F curve_rhs(const F& x, const F& b) {
    if constexpr (!CurveCoeffA<F>::ZERO) return fp_add(fp_mul(fp_add(fp_sqr(x), CurveCoeffA<F>::a()), x), b);
    else return fp_add(fp_mul(fp_sqr(x), x), b);
}
template <class F>
SS_HD bool on_curve(const Affine<F>& p, const F& b) {
    if (p.inf) return true;
    return fp_sqr(p.y) == curve_rhs(p.x, b);
}

// MSB-first double-and-add over `nbits` bits of a little-endian limb array (the reference
// algorithm, kept as the simple/validated baseline and for the r-multiplication subgroup check).
template <class F, class LimbFn>
SS_HD Jac<F> jac_mul_bits(const Affine<F>& base, LimbFn limb, int nbits) {
    Jac<F> acc = Jac<F>::identity();
    if (base.inf) return acc;
    for (int i = nbits - 1; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((limb(i >> 5) >> (i & 31)) & 1) acc = jac_madd(acc, base);
    }
    return acc;
}

// Jacobian -> affine given 1/Z
template <class F>
SS_HD Affine<F> jac_to_affine_with_zinv(const Jac<F>& p, const F& zinv) {
    F zi2 = fp_sqr(zinv);
    Affine<F> r;
    r.x = fp_mul(p.X, zi2);
    r.y = fp_mul(p.Y, fp_mul(zi2, zinv));
    r.inf = false;
    return r;
}

}  // namespace ss
