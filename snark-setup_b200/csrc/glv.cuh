// Endomorphism-accelerated variable-base scalar multiplication.
//
// The reference multiplies every element with arkworks' MSB-first double-and-add
// (setup-utils/src/helpers.rs:95-106 -> ark-ec `mul_bigint`): ~lambda doublings + lambda/2 additions,
// and on a SIMT machine the addition executes for the whole warp whenever any lane has a 1 bit.
// The result of batch_exp is a canonical affine point, so any exact algorithm gives the same bytes
// (SURVEY.md App. A.3).  This file computes k*P as
//     2-dim GLV   k = k1 + k2*lambda,           phi(x,y)  = (beta*x, y)        (BLS12-377 G1, BW6-761 G1/G2)
//     4-dim GLS   k = k0 + k1*u + k2*u^2 + k3*u^3, psi(x,y) = (cx*conj(x), cy*conj(y))   (BLS12-377 G2)
// with all sub-scalars sharing one doubling chain (127 / 190 / 64 doublings instead of 253 / 377 /
// 253), signed 4-bit fixed windows (uniform control flow across a warp), and a per-thread table
// {1..8}P brought to a COMMON Z without any inversion: on an a = 0 curve the doubling and mixed-addition
// formulas do not involve b, so the rescaled table entries (X_j f_j^2, Y_j f_j^3), f_j = Zc/Z_j, are
// affine points of the isomorphic curve y^2 = x^3 + b*Zc^6 and every table addition is a 7M+4S mixed
// addition; the final Z is multiplied by Zc to come back.  phi and psi commute with that isomorphism
// (for psi the common Z is made real: Zc*conj(Zc)).
#pragma once
#include "ec.cuh"

namespace ss {

// ---- tiny multiword helpers (little-endian u32 words, runtime sizes; one-off per element) ----------
SS_HD void mp_mul(const uint32_t* a, int na, const uint32_t* b, int nb, uint32_t* out /*na+nb*/) {
    for (int i = 0; i < na + nb; i++) out[i] = 0;
    for (int i = 0; i < na; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < nb; j++) {
            uint64_t t = (uint64_t)a[i] * b[j] + out[i + j] + carry;
            out[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        out[i + nb] = (uint32_t)carry;
    }
}
// acc[0..n) += sign * v[0..nv)   (two's complement, wraps mod 2^(32n))
SS_HD void mp_addsub(uint32_t* acc, int n, const uint32_t* v, int nv, int sign) {
    if (sign > 0) {
        uint64_t c = 0;
        for (int i = 0; i < n; i++) {
            uint64_t t = (uint64_t)acc[i] + (i < nv ? v[i] : 0u) + c;
            acc[i] = (uint32_t)t;
            c = t >> 32;
        }
    } else {
        uint64_t b = 0;
        for (int i = 0; i < n; i++) {
            uint64_t t = (uint64_t)acc[i] - (i < nv ? v[i] : 0u) - b;
            acc[i] = (uint32_t)t;
            b = (t >> 32) & 1;
        }
    }
}
// two's complement -> (magnitude, negative?)
SS_HD bool mp_abs(uint32_t* a, int n) {
    bool neg = (a[n - 1] >> 31) != 0;
    if (neg) {
        uint64_t c = 1;
        for (int i = 0; i < n; i++) {
            uint64_t t = (uint64_t)(~a[i]) + c;
            a[i] = (uint32_t)t;
            c = t >> 32;
        }
    }
    return neg;
}
SS_HD bool mp_geq(const uint32_t* a, const uint32_t* b, int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}

// signed 4-bit fixed-window recoding of a magnitude: digits in [-7, 8], sum d_i 16^i = m
SS_HD void recode_w4(const uint32_t* m, int nwords, int8_t* dig, int nd, bool negate) {
    int carry = 0;
    for (int i = 0; i < nd; i++) {
        int w = (4 * i) >> 5, s = (4 * i) & 31;
        int d = (w < nwords ? (int)((m[w] >> s) & 15u) : 0) + carry;
        if (d > 8) {
            d -= 16;
            carry = 1;
        } else {
            carry = 0;
        }
        dig[i] = (int8_t)(negate ? -d : d);
    }
}

// ---- 2-dim GLV decomposition ----------------------------------------------------------------------
// P provides KW, HW, GW, ND, g1/g2/a1/a2/b1/b2 magnitudes and the sign pattern S11,S12,S21,S22 with
//   C_i = (k*g_i) >> 32*KW ;  k1 = k - S11*C1*|a1| - S12*C2*|a2| ;  k2 = -S21*C1*|b1| - S22*C2*|b2|
// (Babai rounding against a short basis of {(x,y): x + y*lambda = 0 mod r}; k1 + k2*lambda = k mod r holds
// for ANY C_i, the rounding only keeps |k1|,|k2| < 2^(4*(ND-1)).)
template <class P>
SS_HD void glv_decompose(const uint32_t* k, int8_t* dig1, int8_t* dig2) {
    constexpr int KW = P::KW, HW = P::HW, GW = P::GW, L = KW + 1;
    uint32_t g[GW], prod[KW + GW], C1[HW], C2[HW], t[2 * HW], acc[L], a[HW];
    for (int i = 0; i < GW; i++) g[i] = P::g1(i);
    mp_mul(k, KW, g, GW, prod);
    for (int i = 0; i < HW; i++) C1[i] = prod[KW + i];
    for (int i = 0; i < GW; i++) g[i] = P::g2(i);
    mp_mul(k, KW, g, GW, prod);
    for (int i = 0; i < HW; i++) C2[i] = prod[KW + i];
    // k1
    for (int i = 0; i < KW; i++) acc[i] = k[i];
    acc[KW] = 0;
    for (int i = 0; i < HW; i++) a[i] = P::a1(i);
    mp_mul(C1, HW, a, HW, t);
    mp_addsub(acc, L, t, 2 * HW, -P::S11);
    for (int i = 0; i < HW; i++) a[i] = P::a2(i);
    mp_mul(C2, HW, a, HW, t);
    mp_addsub(acc, L, t, 2 * HW, -P::S12);
    bool neg = mp_abs(acc, L);
    recode_w4(acc, L, dig1, P::ND, neg);
    // k2
    for (int i = 0; i < L; i++) acc[i] = 0;
    for (int i = 0; i < HW; i++) a[i] = P::b1(i);
    mp_mul(C1, HW, a, HW, t);
    mp_addsub(acc, L, t, 2 * HW, -P::S21);
    for (int i = 0; i < HW; i++) a[i] = P::b2(i);
    mp_mul(C2, HW, a, HW, t);
    mp_addsub(acc, L, t, 2 * HW, -P::S22);
    neg = mp_abs(acc, L);
    recode_w4(acc, L, dig2, P::ND, neg);
}

// ---- 4-dim base-u decomposition (BLS12-377 G2: psi = [u] on G2, r = u^4 - u^2 + 1) -----------------
// q = floor(k / d) via a precomputed m = floor(2^(32*kw) / d): q' = (k*m) >> 32*kw, one correction.
SS_HD void mp_divrem_const(const uint32_t* k, int kw, const uint32_t* d, int dw, const uint32_t* m, int mw, uint32_t* q /*2*/,
                           uint32_t* rem /*dw*/) {
    uint32_t prod[24], t[24], r[13];
    mp_mul(k, kw, m, mw, prod);
    q[0] = prod[kw];
    q[1] = (kw + 1 < kw + mw) ? prod[kw + 1] : 0u;
    mp_mul(q, 2, d, dw, t);  // dw + 2 words
    for (int i = 0; i < dw + 1; i++) r[i] = i < kw ? k[i] : 0u;
    mp_addsub(r, dw + 1, t, dw + 1, -1);
    uint32_t dd[13];
    for (int i = 0; i < dw; i++) dd[i] = d[i];
    dd[dw] = 0;
    if (mp_geq(r, dd, dw + 1)) {
        mp_addsub(r, dw + 1, dd, dw + 1, -1);
        if (++q[0] == 0) q[1]++;
    }
    for (int i = 0; i < dw; i++) rem[i] = r[i];
}

template <class P>
SS_HD void gls4_decompose(const uint32_t* k, int8_t* dig /*[4][ND]*/) {
    uint32_t d[6], m[6], q3[2], q2[2], q1[2], r3[6], r2[4], r1[2];
    for (int i = 0; i < 6; i++) d[i] = P::u3(i);
    for (int i = 0; i < 3; i++) m[i] = P::m3(i);
    mp_divrem_const(k, 8, d, 6, m, 3, q3, r3);
    for (int i = 0; i < 4; i++) d[i] = P::u2(i);
    for (int i = 0; i < 3; i++) m[i] = P::m2(i);
    mp_divrem_const(r3, 6, d, 4, m, 3, q2, r2);
    for (int i = 0; i < 2; i++) d[i] = P::u1(i);
    for (int i = 0; i < 3; i++) m[i] = P::m1(i);
    mp_divrem_const(r2, 4, d, 2, m, 3, q1, r1);
    recode_w4(r1, 2, dig + 0 * P::ND, P::ND, false);
    recode_w4(q1, 2, dig + 1 * P::ND, P::ND, false);
    recode_w4(q2, 2, dig + 2 * P::ND, P::ND, false);
    recode_w4(q3, 2, dig + 3 * P::ND, P::ND, false);
}

// ---- per-group endomorphism traits ----------------------------------------------------------------
template <class G>
struct Endo;  // DIMS, ND, decompose(k, dig), apply(j, x, y), real_factor(Zc, out) -> bool

template <class Glv, class FP>
struct Endo2 {  // phi(x, y) = (beta x, y) over a prime field
    using F = Fp<FP>;
    static constexpr int DIMS = 2, ND = Glv::ND;
    SS_HD static void decompose(const uint32_t* k, int8_t* dig) { glv_decompose<Glv>(k, dig, dig + ND); }
    SS_HD static void apply(int j, F& x, F& y) {
        if (j == 1) {
            F b;
#pragma unroll
            for (int i = 0; i < FP::N; i++) b.l[i] = Glv::beta(i);
            x = fp_mul(x, b);
        }
        (void)y;
    }
    SS_HD static bool real_factor(const F&, F&) { return false; }
};

template <>
struct Endo<Bls377G1> : Endo2<Bls377G1Glv, Bls377Fq> {};
template <>
struct Endo<Bw6G1> : Endo2<Bw6G1Glv, Bw6Fq> {};
template <>
struct Endo<Bw6G2> : Endo2<Bw6G2Glv, Bw6Fq> {};

template <>
struct Endo<Bls377G2> {
    using B = Fp<Bls377Fq>;
    using F = Fp2<Bls377Fq>;
    using P = Bls377G2Gls;
    static constexpr int DIMS = 4, ND = P::ND;
    SS_HD static void decompose(const uint32_t* k, int8_t* dig) { gls4_decompose<P>(k, dig); }
    SS_HD static B cst(int which) {
        B c;
#pragma unroll
        for (int i = 0; i < 12; i++) c.l[i] = which == 0 ? P::cx(i) : which == 1 ? P::cy(i) : P::cx2(i);
        return c;
    }
    // psi(x,y) = (cx conj x, cy conj y); psi^2 = (cx^2 x, -y); psi^3 = (-conj x, -cy conj y)
    SS_HD static void apply(int j, F& x, F& y) {
        if (j == 1) {
            B cx = cst(0), cy = cst(1);
            x = F{fp_mul(x.c0, cx), fp_neg(fp_mul(x.c1, cx))};
            y = F{fp_mul(y.c0, cy), fp_neg(fp_mul(y.c1, cy))};
        } else if (j == 2) {
            B c = cst(2);
            x = F{fp_mul(x.c0, c), fp_mul(x.c1, c)};
            y = fp_neg(y);
        } else if (j == 3) {
            B cy = cst(1);
            x = F{fp_neg(x.c0), x.c1};
            y = F{fp_neg(fp_mul(y.c0, cy)), fp_mul(y.c1, cy)};
        }
    }
    // psi needs the common Z in Fq: multiply everything by conj(Zc)
    SS_HD static bool real_factor(const F& zc, F& out) {
        out = F{zc.c0, fp_neg(zc.c1)};
        return true;
    }
};

#if defined(__CUDACC__)
// the same endomorphism on the lane-split representation: every map is per-lane (a base-field multiplication of the
// own half, a sign on the odd lane), no exchange needed
template <>
struct Endo<Bls377G2L> {
    using B = Fp<Bls377Fq>;
    using F = Fp2L<Bls377Fq>;
    using P = Bls377G2Gls;
    static constexpr int DIMS = 4, ND = P::ND;
    SS_HD static void decompose(const uint32_t* k, int8_t* dig) { gls4_decompose<P>(k, dig); }
    SS_HD static void apply(int j, F& x, F& y) {
        if (j == 1) {
            x = fp_conj(fp_mul_base(x, Endo<Bls377G2>::cst(0)));
            y = fp_conj(fp_mul_base(y, Endo<Bls377G2>::cst(1)));
        } else if (j == 2) {
            x = fp_mul_base(x, Endo<Bls377G2>::cst(2));
            y = fp_neg(y);
        } else if (j == 3) {
            x = fp_conj(fp_neg(x));
            y = fp_neg(fp_conj(fp_mul_base(y, Endo<Bls377G2>::cst(1))));
        }
    }
    SS_HD static bool real_factor(const F& zc, F& out) {
        out = fp_conj(zc);
        return true;
    }
};
#endif

// out-of-line plain ladder for the (never taken on subgroup inputs) tiny-order fallback
template <class G>
#if defined(__CUDACC__)
__device__ __host__ __noinline__
#endif
Jac<typename G::F> jac_mul_ladder_cold(Affine<typename G::F> base, const uint32_t* k) {
    return jac_mul_bits<typename G::F>(base, [&](int i) { return k[i]; }, G::Fr::Params::BITS);
}

// One window of the main loop: 4 doublings, then for each sub-scalar one mixed addition of the (endomorphism
// image of the) table entry selected by its signed digit.  For the wide fields this is ONE out-of-line function
// (with the doubling and the addition inlined once each) so the accumulator crosses a call boundary once per
// window instead of once per group operation: as separate calls the by-value Jacobian arguments cost
// ~1300 words of local-memory traffic per window and k_scalar_mul<G2> stalled on them (3100 local
// loads/stores per thread, fmaheavy pipe 75 % active; profiles/r01_ncu_full_v3.json).
template <class G>
SS_HD void endo_window_inl(Jac<typename G::F>& acc, const typename G::F* tx, const typename G::F* ty, const int8_t* dig,
                           int i, bool first) {
    using F = typename G::F;
    using E = Endo<G>;
    if (!first) {
#pragma unroll 1
        for (int d = 0; d < 4; d++) acc = jac_dbl_inl(acc);
    }
#pragma unroll 1
    for (int j = 0; j < E::DIMS; j++) {
        const int d = dig[j * E::ND + i];
        if (d != 0) {
            const int a = d < 0 ? -d : d;
            Affine<F> q;
            q.x = tx[a - 1];
            q.y = ty[a - 1];
            q.inf = false;
            E::apply(j, q.x, q.y);
            if (d < 0) q.y = fp_neg(q.y);
            acc = jac_madd_inl(acc, q);
        }
    }
}
#if defined(__CUDACC__)
template <class G>
__device__ __noinline__ void endo_window_call(Jac<typename G::F>* acc, const typename G::F* tx, const typename G::F* ty,
                                              const int8_t* dig, int i, int first) {
    Jac<typename G::F> a = *acc;
    endo_window_inl<G>(a, tx, ty, dig, i, first != 0);
    *acc = a;
}
#endif
template <class G>
SS_HD void endo_window(Jac<typename G::F>& acc, const typename G::F* tx, const typename G::F* ty, const int8_t* dig, int i,
                       bool first) {
#if defined(__CUDA_ARCH__) && !defined(SS_GROUP_INLINE)
    if constexpr (G::F::CALL_GROUP_OPS) {
        endo_window_call<G>(&acc, tx, ty, dig, i, first ? 1 : 0);
        return;
    }
#endif
    endo_window_inl<G>(acc, tx, ty, dig, i, first);
}

// ---- k * P ----------------------------------------------------------------------------------------
// `k` canonical little-endian words (G::Fr::N of them), k < r.
// PRECONDITION: `base` lies in the order-r subgroup (phi / psi act as the scalars lambda / u only there).  That is
// what a ceremony's challenge holds — the coordinator verified it, which is why the reference's contribute reads it
// with CheckForCorrectness::No (setup-utils/src/helpers.rs:550) — and what CHECK_FULL / CHECK_ONLY_IN_GROUP have
// tested.  `plain` = true selects the reference's own MSB-first double-and-add instead (ss_set_strict_unchecked_inputs),
// which reproduces ark-ec's mul_bigint on ANY input, off-subgroup or off-curve (the a = 0 formulas do not involve b).
template <class G>
SS_HD Jac<typename G::F> scalar_mul_endo(const Affine<typename G::F>& base, const uint32_t* k, bool plain = false) {
    using F = typename G::F;
    using E = Endo<G>;
    constexpr int ND = E::ND, DIMS = E::DIMS, TS = 8;
    if (base.inf) return Jac<F>::identity();
    if (plain) return jac_mul_ladder_cold<G>(base, k);
    // table j*P, j = 1..8, Jacobian
    F tx[TS], ty[TS], tz[TS];
    {
        Jac<F> t{base.x, base.y, F::one()};
        tx[0] = t.X; ty[0] = t.Y; tz[0] = t.Z;
        t = jac_dbl(t);
        tx[1] = t.X; ty[1] = t.Y; tz[1] = t.Z;
#pragma unroll 1
        for (int j = 2; j < TS; j++) {
            t = jac_madd(t, base);
            tx[j] = t.X; ty[j] = t.Y; tz[j] = t.Z;
        }
    }
    // common Z: f_j = prod_{i != j} Z_i (Z_1 = 1 is skipped), Zc = prod Z_i
    F pre[TS];  // pre[j] = prod_{1 <= i < j} tz[i]
    pre[1] = F::one();
    for (int j = 2; j < TS; j++) pre[j] = fp_mul(pre[j - 1], tz[j - 1]);
    F zc = fp_mul(pre[TS - 1], tz[TS - 1]);
    if (zc.is_zero()) {
        // some j*P (j <= 8) is the identity: P has tiny order — take the plain ladder
        return jac_mul_ladder_cold<G>(base, k);
    }
    F extra;
    const bool has_extra = E::real_factor(zc, extra);
    F zfinal = has_extra ? fp_mul(zc, extra) : zc;
    {
        F suf = has_extra ? extra : F::one();  // running prod_{i > j} tz[i] (* extra)
        for (int j = TS - 1; j >= 1; j--) {
            F f = fp_mul(pre[j], suf);
            suf = fp_mul(suf, tz[j]);
            F f2 = fp_sqr(f);
            tx[j] = fp_mul(tx[j], f2);
            ty[j] = fp_mul(ty[j], fp_mul(f2, f));
        }
        // j = 0 (Z = 1): f = Zc (* extra) = suf
        F f2 = fp_sqr(suf);
        tx[0] = fp_mul(tx[0], f2);
        ty[0] = fp_mul(ty[0], fp_mul(f2, suf));
    }
    int8_t dig[DIMS * ND];
    E::decompose(k, dig);
    Jac<F> acc = Jac<F>::identity();
    for (int i = ND - 1; i >= 0; i--) endo_window<G>(acc, tx, ty, dig, i, i == ND - 1);
    acc.Z = fp_mul(acc.Z, zfinal);
    return acc;
}


#if defined(__CUDACC__)
// ---- k * P on the lane-split group: scalar_mul_endo under the converged-warp discipline of fp2l.cuh -----------------
// Same algorithm (common-Z table of {1..8} P, 4-dim GLS digits, signed 4-bit windows); what changes is control flow:
// a zero digit adds a dummy entry and keeps the old accumulator by selection, an infinite base runs on a harmless
// stand-in and is resolved at the end, and the tiny-order fallback (some j P = O) is entered by the whole warp.
template <class GL>
SS_D Jac<typename GL::F> scalar_mul_endo_pair(const Affine<typename GL::F>& base_in, const uint32_t* k, bool plain) {
    using F = typename GL::F;
    using E = Endo<GL>;
    using FrP = typename GL::Fr::Params;
    constexpr int ND = E::ND, DIMS = E::DIMS, TS = 8;
    Affine<F> base = base_in;
    if (plain)  // kernel argument: uniform
        return jac_mul_bits_pair<typename F::Params>(base, [&](int i) { return k[i]; }, FrP::BITS);
    base.inf = false;
    // an infinite base carries x = y = 0: the table arithmetic below runs on it without harm (Z products become 0,
    // which would request the ladder) — its result is replaced by the identity at the end and it never votes
    F tx[TS], ty[TS], tz[TS];
    {
        Jac<F> t{base.x, base.y, F::one()};
        tx[0] = t.X; ty[0] = t.Y; tz[0] = t.Z;
        t = jac_dbl_inl(t);
        tx[1] = t.X; ty[1] = t.Y; tz[1] = t.Z;
#pragma unroll 1
        for (int j = 2; j < TS; j++) {
            t = jac_madd_inl(t, base);
            tx[j] = t.X; ty[j] = t.Y; tz[j] = t.Z;
        }
    }
    F pre[TS];
    pre[1] = F::one();
#pragma unroll 1
    for (int j = 2; j < TS; j++) pre[j] = fp_mul(pre[j - 1], tz[j - 1]);
    F zc = fp_mul(pre[TS - 1], tz[TS - 1]);
    const bool need_ladder = !base_in.inf && zc.is_zero();
    F extra;
    E::real_factor(zc, extra);
    F zfinal = fp_mul(zc, extra);
    {
        F suf = extra;
#pragma unroll 1
        for (int j = TS - 1; j >= 1; j--) {
            F f = fp_mul(pre[j], suf);
            suf = fp_mul(suf, tz[j]);
            F f2 = fp_sqr(f);
            tx[j] = fp_mul(tx[j], f2);
            ty[j] = fp_mul(ty[j], fp_mul(f2, f));
        }
        F f2 = fp_sqr(suf);
        tx[0] = fp_mul(tx[0], f2);
        ty[0] = fp_mul(ty[0], fp_mul(f2, suf));
    }
    int8_t dig[DIMS * ND];
    E::decompose(k, dig);
    Jac<F> acc = Jac<F>::identity();
#pragma unroll 1
    for (int i = ND - 1; i >= 0; i--) {
        if (i != ND - 1) {
#pragma unroll 1
            for (int d = 0; d < 4; d++) acc = jac_dbl_inl(acc);
        }
#pragma unroll 1
        for (int j = 0; j < DIMS; j++) {
            const int d = dig[j * ND + i];
            const int a = d < 0 ? -d : (d == 0 ? 1 : d);
            Affine<F> q;
            q.x = tx[a - 1];
            q.y = ty[a - 1];
            q.inf = false;
            E::apply(j, q.x, q.y);
            q.y = f2l_select(d < 0, fp_neg(q.y), q.y);
            acc = jac_select(d != 0, jac_madd_inl(acc, q), acc);
        }
    }
    acc.Z = fp_mul(acc.Z, zfinal);
    if (warp_any(need_ladder)) {
        Jac<F> lad = jac_mul_bits_pair<typename F::Params>(base, [&](int i) { return k[i]; }, FrP::BITS);
        acc = jac_select(need_ladder, lad, acc);
    }
    return jac_select(base_in.inf, Jac<F>::identity(), acc);
}
#endif

}  // namespace ss
