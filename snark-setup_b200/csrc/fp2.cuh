// Quadratic extension Fq2 = Fq[u]/(u^2 + 5) used by BLS12-377 G2.
// Device-side stand-in for ark-ff `Fp2<Fq2Config>` of ark-bls12-377 0.4.0 (NONRESIDUE = -5).
#pragma once
#include "fp.cuh"

namespace ss {

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

// Karatsuba: 3 base multiplications
template <class P>
SS_HD Fp2<P> fp_mul(const Fp2<P>& a, const Fp2<P>& b) {
    Fp<P> v0 = fp_mul(a.c0, b.c0);
    Fp<P> v1 = fp_mul(a.c1, b.c1);
#if defined(SS_FP2_REDUCED_SUMS)
    Fp<P> s = fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
#else
    Fp<P> s = fp_mul(fp_add_nr(a.c0, a.c1), fp_add_nr(b.c0, b.c1));  // operand sums stay below 2p (fp.cuh)
#endif
    Fp2<P> r;
    r.c1 = fp_sub(fp_sub(s, v0), v1);
    r.c0 = fp_sub(v0, fp_mul5(v1));
    return r;
}

// complex squaring: 2 base multiplications.
// c0 = a0^2 - 5 a1^2 = (a0 + a1)(a0 - 5 a1) + 4 a0 a1 ;  c1 = 2 a0 a1
template <class P>
SS_HD Fp2<P> fp_sqr(const Fp2<P>& a) {
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

// 1/(a0 + a1 u) = (a0 - a1 u) / (a0^2 + 5 a1^2)
template <class P>
SS_HD Fp2<P> fp_inv(const Fp2<P>& a) {
    Fp<P> n = fp_add(fp_sqr(a.c0), fp_mul5(fp_sqr(a.c1)));
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
        // a0 is a non-residue in Fq: the root is c1*u with c1^2 * (-5) = a0
        // -5 is a non-residue too, so a0 / -5 is a residue.
        Fp<P> five = fp_mul5(Fp<P>::one());
        Fp<P> q = fp_neg(fp_mul(a.c0, fp_inv(five)));
        if (!fp_sqrt(q, r)) return false;  // unreachable for a prime field
        out = Fp2<P>{Fp<P>::zero(), r};
        return true;
    }
    Fp<P> norm = fp_add(fp_sqr(a.c0), fp_mul5(fp_sqr(a.c1)));
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

}  // namespace ss
