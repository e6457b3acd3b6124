// Quadratic extension Fq2 = Fq[u]/(u^2 - NR): NR = -5 for BLS12-377 G2 (ark-bls12-377 0.4.0 Fq2Config), NR = 13 for
// MNT4-753 G2 (ark-mnt4-753 0.4.0).  Device-side stand-in for ark-ff `Fp2<Fq2Config>`.  The BLS12-377 code paths are
// the hot ones and are kept exactly as measured; other non-residues take the generic branches.
#pragma once
#include "constants_gen.cuh"
#include "fp.cuh"

namespace ss {

template <class P>
struct Fp2Config {
    static constexpr int NR = -5;
};
template <>
struct Fp2Config<Mnt753Q> {
    static constexpr int NR = 13;
};

// x * k for a small non-negative constant k (double-and-add over the bits of k, fully reduced)
template <class P>
SS_HD Fp<P> fp_mul_small(const Fp<P>& x, unsigned k) {
    Fp<P> r = Fp<P>::zero();
    bool started = false;
    for (int b = 31; b >= 0; b--) {
        if (started) r = fp_dbl(r);
        if ((k >> b) & 1u) {
            r = started ? fp_add(r, x) : x;
            started = true;
        }
    }
    return r;
}
// x * NR (signed)
template <class P>
SS_HD Fp<P> fp_mul_nr(const Fp<P>& x) {
    constexpr int NR = Fp2Config<P>::NR;
    Fp<P> m = fp_mul_small(x, (unsigned)(NR < 0 ? -NR : NR));
    return NR < 0 ? fp_neg(m) : m;
}

template <class P>
struct Fp2 {
    using Base = Fp<P>;
    using Params = P;
    static constexpr bool CALL_GROUP_OPS = true;
    Base c0, c1;

    SS_HD static Fp2 zero() { return Fp2{Base::zero(), Base::zero()}; }
    SS_HD static Fp2 one() { return Fp2{Base::one(), Base::zero()}; }
    SS_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    SS_HD bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
    SS_HD bool operator!=(const Fp2& o) const { return !(*this == o); }
};

// x * 5
template <class P>
SS_HD Fp<P> fp_mul5(const Fp<P>& x) {
    Fp<P> t = fp_dbl(fp_dbl(x));
    return fp_add(t, x);
}

template <class P>
SS_HD Fp2<P> fp_add(const Fp2<P>& a, const Fp2<P>& b) {
    return Fp2<P>{fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)};
}
template <class P>
SS_HD Fp2<P> fp_sub(const Fp2<P>& a, const Fp2<P>& b) {
    return Fp2<P>{fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)};
}
template <class P>
SS_HD Fp2<P> fp_neg(const Fp2<P>& a) {
    return Fp2<P>{fp_neg(a.c0), fp_neg(a.c1)};
}
template <class P>
SS_HD Fp2<P> fp_dbl(const Fp2<P>& a) {
    return Fp2<P>{fp_dbl(a.c0), fp_dbl(a.c1)};
}

// Out-of-line Fq2 units (device, NR = -5, <= 12 limbs): ONE call per Fq2 product / square instead of 3 / 2 calls of
// the base multiplier, with the base products row-interleaved (fp_mul_xk_inl) so that a single warp has 3 / 2
// independent carry chains in flight.  Measured SLOWER (k_scalar_mul<G2> 83.1 -> 90.6 ms per 2^19 launch,
// profiles/r02_ab_variants.md): the unit needs ~170 registers, so its callers spill their Jacobian state around every
// call.  Off by default; -DSS_FP2_UNITS=1 builds it; tests/emul checks the bodies either way.
#ifndef SS_FP2_UNITS
#define SS_FP2_UNITS 0
#endif
template <class P>
SS_HD Fp2<P> fp2_mul_body(const Fp2<P>& a, const Fp2<P>& b) {
    const Fp<P> x[3] = {a.c0, a.c1, fp_add_nr(a.c0, a.c1)};
    const Fp<P> y[3] = {b.c0, b.c1, fp_add_nr(b.c0, b.c1)};
    Fp<P> v[3];
    fp_mul_xk_inl<P, 3>(x, y, v);
    Fp2<P> r;
    r.c1 = fp_sub(fp_sub(v[2], v[0]), v[1]);
    r.c0 = fp_sub(v[0], fp_mul5(v[1]));
    return r;
}
template <class P>
SS_HD Fp2<P> fp2_sqr_body(const Fp2<P>& a) {
    const Fp<P> x[2] = {a.c0, fp_add_nr(a.c0, a.c1)};
    const Fp<P> y[2] = {a.c1, fp_sub(a.c0, fp_mul5(a.c1))};
    Fp<P> v[2];
    fp_mul_xk_inl<P, 2>(x, y, v);
    const Fp<P> v2 = fp_dbl(v[0]);
    Fp2<P> r;
    r.c0 = fp_add(v[1], fp_dbl(v2));
    r.c1 = v2;
    return r;
}
#if defined(__CUDACC__)
template <class P>
__device__ __noinline__ Fp2<P> fp2_mul_call(Fp2<P> a, Fp2<P> b) {
    return fp2_mul_body(a, b);
}
template <class P>
__device__ __noinline__ Fp2<P> fp2_sqr_call(Fp2<P> a) {
    return fp2_sqr_body(a);
}
#endif
template <class P>
constexpr bool kFp2Units = SS_FP2_UNITS && Fp2Config<P>::NR == -5 && P::N <= 12;

// Karatsuba: 3 base multiplications
template <class P>
SS_HD Fp2<P> fp_mul(const Fp2<P>& a, const Fp2<P>& b) {
#if defined(__CUDA_ARCH__) && !defined(SS_MUL_INLINE)
    if constexpr (kFp2Units<P>) return fp2_mul_call<P>(a, b);
#endif
    Fp<P> v0 = fp_mul(a.c0, b.c0);
    Fp<P> v1 = fp_mul(a.c1, b.c1);
#if defined(SS_FP2_REDUCED_SUMS)
    Fp<P> s = fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
#else
    Fp<P> s = fp_mul(fp_add_nr(a.c0, a.c1), fp_add_nr(b.c0, b.c1));  // operand sums stay below 2p (fp.cuh)
#endif
    Fp2<P> r;
    r.c1 = fp_sub(fp_sub(s, v0), v1);
    if constexpr (Fp2Config<P>::NR == -5) r.c0 = fp_sub(v0, fp_mul5(v1));
    else r.c0 = fp_add(v0, fp_mul_nr(v1));
    return r;
}

// complex squaring: 2 base multiplications.
// c0 = a0^2 - 5 a1^2 = (a0 + a1)(a0 - 5 a1) + 4 a0 a1 ;  c1 = 2 a0 a1
template <class P>
SS_HD Fp2<P> fp_sqr(const Fp2<P>& a) {
#if defined(__CUDA_ARCH__) && !defined(SS_MUL_INLINE)
    if constexpr (kFp2Units<P>) return fp2_sqr_call<P>(a);
#endif
    if constexpr (Fp2Config<P>::NR != -5) {
        // c0 = a0^2 + NR a1^2 = (a0 + a1)(a0 + NR a1) - (NR + 1) a0 a1 ;  c1 = 2 a0 a1
        Fp<P> v = fp_mul(a.c0, a.c1);
        Fp<P> t = fp_mul(fp_add(a.c0, a.c1), fp_add(a.c0, fp_mul_nr(a.c1)));
        constexpr int K = Fp2Config<P>::NR + 1;
        Fp<P> kv = fp_mul_small(v, (unsigned)(K < 0 ? -K : K));
        return Fp2<P>{K < 0 ? fp_add(t, kv) : fp_sub(t, kv), fp_dbl(v)};
    }
    Fp<P> v = fp_mul(a.c0, a.c1);
#if defined(SS_FP2_REDUCED_SUMS)
    Fp<P> t = fp_mul(fp_add(a.c0, a.c1), fp_sub(a.c0, fp_mul5(a.c1)));
#else
    Fp<P> t = fp_mul(fp_add_nr(a.c0, a.c1), fp_sub(a.c0, fp_mul5(a.c1)));
#endif
    Fp<P> v2 = fp_dbl(v);
    Fp2<P> r;
    r.c0 = fp_add(t, fp_dbl(v2));
    r.c1 = v2;
    return r;
}

template <class P>
SS_HD Fp2<P> fp_mul_base(const Fp2<P>& a, const Fp<P>& k) {
    return Fp2<P>{fp_mul(a.c0, k), fp_mul(a.c1, k)};
}

// 1/(a0 + a1 u) = (a0 - a1 u) / (a0^2 - NR a1^2)
template <class P>
SS_HD Fp2<P> fp_inv(const Fp2<P>& a) {
    Fp<P> n = fp_sub(fp_sqr(a.c0), fp_mul_nr(fp_sqr(a.c1)));
    Fp<P> ni = fp_inv(n);
    return Fp2<P>{fp_mul(a.c0, ni), fp_neg(fp_mul(a.c1, ni))};
}

// Square root in Fq2 (any root).  norm = a0^2 + 5 a1^2; alpha = sqrt(norm);
// delta = (a0 +- alpha)/2; c0 = sqrt(delta); c1 = a1 / (2 c0).
template <class P>
SS_HD bool fp_sqrt(const Fp2<P>& a, Fp2<P>& out) {
    if (a.c1.is_zero()) {
        Fp<P> r;
        if (fp_sqrt(a.c0, r)) {
            out = Fp2<P>{r, Fp<P>::zero()};
            return true;
        }
        // a0 is a non-residue in Fq: the root is c1*u with c1^2 * NR = a0
        // NR is a non-residue too, so a0 / NR is a residue.
        Fp<P> q = fp_mul(a.c0, fp_inv(fp_mul_nr(Fp<P>::one())));
        if (!fp_sqrt(q, r)) return false;  // unreachable for a prime field
        out = Fp2<P>{Fp<P>::zero(), r};
        return true;
    }
    Fp<P> norm = fp_sub(fp_sqr(a.c0), fp_mul_nr(fp_sqr(a.c1)));
    Fp<P> alpha;
    if (!fp_sqrt(norm, alpha)) return false;
    Fp<P> half;
#pragma unroll
    for (int i = 0; i < P::N; i++) half.l[i] = P::half(i);
    Fp<P> delta = fp_mul(fp_add(a.c0, alpha), half);
    Fp<P> c0;
    if (!fp_sqrt(delta, c0)) {
        delta = fp_mul(fp_sub(a.c0, alpha), half);
        if (!fp_sqrt(delta, c0)) return false;
    }
    Fp<P> c1 = fp_mul(a.c1, fp_inv(fp_dbl(c0)));
    out = Fp2<P>{c0, c1};
    return true;
}


// ---- cubic extension Fq3 = Fq[u]/(u^3 - 11): MNT6-753 G2 (ark-mnt6-753 0.4.0 Fq3Config, NONRESIDUE = 11) -----------
// Functional, not tuned: the MNT curves are exposed by the reference (setup-utils/src/converters.rs:18-45) but are not
// the ceremony curves; schoolbook products (9 base multiplications), Tonelli-Shanks square roots in Fq3*.
template <class P>
struct Fp3 {
    using Base = Fp<P>;
    using Params = P;
    static constexpr bool CALL_GROUP_OPS = true;
    static constexpr unsigned NR = 11;
    Base c0, c1, c2;

    SS_HD static Fp3 zero() { return Fp3{Base::zero(), Base::zero(), Base::zero()}; }
    SS_HD static Fp3 one() { return Fp3{Base::one(), Base::zero(), Base::zero()}; }
    SS_HD bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
    SS_HD bool operator==(const Fp3& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    SS_HD bool operator!=(const Fp3& o) const { return !(*this == o); }
};
template <class P>
SS_HD Fp3<P> fp_add(const Fp3<P>& a, const Fp3<P>& b) { return Fp3<P>{fp_add(a.c0, b.c0), fp_add(a.c1, b.c1), fp_add(a.c2, b.c2)}; }
template <class P>
SS_HD Fp3<P> fp_sub(const Fp3<P>& a, const Fp3<P>& b) { return Fp3<P>{fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1), fp_sub(a.c2, b.c2)}; }
template <class P>
SS_HD Fp3<P> fp_neg(const Fp3<P>& a) { return Fp3<P>{fp_neg(a.c0), fp_neg(a.c1), fp_neg(a.c2)}; }
template <class P>
SS_HD Fp3<P> fp_dbl(const Fp3<P>& a) { return Fp3<P>{fp_dbl(a.c0), fp_dbl(a.c1), fp_dbl(a.c2)}; }
#if defined(__CUDACC__)
template <class P>
__device__ __host__ __noinline__
#else
template <class P>
inline
#endif
Fp3<P> fp3_mul_impl(const Fp3<P>& a, const Fp3<P>& b) {
    constexpr unsigned NR = Fp3<P>::NR;
    Fp3<P> r;
    r.c0 = fp_add(fp_mul(a.c0, b.c0), fp_mul_small(fp_add(fp_mul(a.c1, b.c2), fp_mul(a.c2, b.c1)), NR));
    r.c1 = fp_add(fp_add(fp_mul(a.c0, b.c1), fp_mul(a.c1, b.c0)), fp_mul_small(fp_mul(a.c2, b.c2), NR));
    r.c2 = fp_add(fp_add(fp_mul(a.c0, b.c2), fp_mul(a.c1, b.c1)), fp_mul(a.c2, b.c0));
    return r;
}
template <class P>
SS_HD Fp3<P> fp_mul(const Fp3<P>& a, const Fp3<P>& b) { return fp3_mul_impl(a, b); }
template <class P>
SS_HD Fp3<P> fp_sqr(const Fp3<P>& a) { return fp3_mul_impl(a, a); }
// standard inverse through the norm: t0 = a0^2 - NR a1 a2, t1 = NR a2^2 - a0 a1, t2 = a1^2 - a0 a2,
// d = a0 t0 + NR (a2 t1 + a1 t2),  a^-1 = (t0, t1, t2) / d
template <class P>
SS_HD Fp3<P> fp_inv(const Fp3<P>& a) {
    constexpr unsigned NR = Fp3<P>::NR;
    Fp<P> t0 = fp_sub(fp_sqr(a.c0), fp_mul_small(fp_mul(a.c1, a.c2), NR));
    Fp<P> t1 = fp_sub(fp_mul_small(fp_sqr(a.c2), NR), fp_mul(a.c0, a.c1));
    Fp<P> t2 = fp_sub(fp_sqr(a.c1), fp_mul(a.c0, a.c2));
    Fp<P> d = fp_add(fp_mul(a.c0, t0), fp_mul_small(fp_add(fp_mul(a.c2, t1), fp_mul(a.c1, t2)), NR));
    Fp<P> di = fp_inv(d);
    return Fp3<P>{fp_mul(t0, di), fp_mul(t1, di), fp_mul(t2, di)};
}
// Tonelli-Shanks in Fq3* (order q^3 - 1 = 2^s t; constants SQ: s, (t - 1)/2, z = non-residue^t)
template <class P, class SQ>
SS_HD bool fp3_sqrt(const Fp3<P>& a, Fp3<P>& out) {
    if (a.is_zero()) {
        out = a;
        return true;
    }
    Fp3<P> w = Fp3<P>::one();
    bool started = false;
    for (int i = SQ::TM1H_LIMBS - 1; i >= 0; i--) {
        const uint32_t e = SQ::tm1h(i);
        for (int b = 31; b >= 0; b--) {
            if (started) w = fp_sqr(w);
            if ((e >> b) & 1u) {
                w = started ? fp_mul(w, a) : a;
                started = true;
            }
        }
    }
    if (!started) w = Fp3<P>::one();
    Fp3<P> x = fp_mul(a, w);
    Fp3<P> b = fp_mul(x, w);
    Fp3<P> z;
#pragma unroll
    for (int i = 0; i < P::N; i++) {
        z.c0.l[i] = SQ::z_c0(i);
        z.c1.l[i] = SQ::z_c1(i);
        z.c2.l[i] = SQ::z_c2(i);
    }
    const Fp3<P> one = Fp3<P>::one();
    int v = SQ::TWO_ADICITY;
    while (!(b == one)) {
        int k = 0;
        Fp3<P> b2k = b;
        while (!(b2k == one)) {
            b2k = fp_sqr(b2k);
            k++;
            if (k == v) return false;  // non-residue
        }
        Fp3<P> wj = z;
        for (int j = 0; j < v - k - 1; j++) wj = fp_sqr(wj);
        z = fp_sqr(wj);
        b = fp_mul(b, z);
        x = fp_mul(x, wj);
        v = k;
    }
    out = x;
    return true;
}

}  // namespace ss
