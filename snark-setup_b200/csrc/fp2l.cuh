// Lane-split Fq2: one Fq2 element lives in TWO adjacent lanes of a warp — the even lane holds c0, the odd lane c1.
//
// Why: the per-thread Fq2 kernels need 255 registers (a Jacobian accumulator alone is 72), so only 8 warps fit on an
// SM and the serial IMAD.WIDE carry chains of two warps per scheduler cannot keep the multiplier pipe full (ncu,
// round 1: fmaheavy 78 % busy, 65 % useful, top stall = fixed-latency "wait", 12 % warps active).  Split over a lane
// pair every thread carries half the state — the register footprint of the G1 kernels — twice as many warps are
// resident, and both lanes run ONE uniform instruction stream:
//     mul:  even lane  c0 = a0 b0 + a1 (-5 b1)        odd lane  c1 = a0 b1 + a1 b0     one fp_mul2 each (3 N^2 MAC)
//     sqr:  even lane  t = (a0 + a1)(a0 - 5 a1)       odd lane  v = a0 a1              one fp_mul each  (2 N^2 MAC)
//           c0 = t + 4 v,  c1 = 2 v
// i.e. 6 N^2 / 4 N^2 multiply-accumulates per Fq2 product / square in total, the same as the 3- / 2-multiplication
// Karatsuba forms of fp2.cuh.  The partner's half travels by full-mask __shfl_xor_sync (see the converged-warp
// discipline below); every predicate (is_zero, ==) is evaluated pair-wide.
//
// The arithmetic is factored into pure functions of (lane parity, own half, partner half) so that the host emulation
// (tests/emul) can check it against the big-integer oracle without a GPU; only the exchange itself is device code.
#pragma once
#include "fp2.cuh"

namespace ss {

// ---- pure per-lane arithmetic (host-testable) -------------------------------------------------------------------
// product: own/oth = this lane's / the partner's half of a and b; returns this lane's half of a*b
template <class P>
SS_HD Fp<P> fp2l_mul_lane(int odd, const Fp<P>& a_own, const Fp<P>& a_oth, const Fp<P>& b_own, const Fp<P>& b_oth) {
    // even: a_own b_own + a_oth (-5 b_oth)      odd: a_oth b_own + a_own b_oth
    const Fp<P> n5 = fp_neg(fp_mul5(b_oth));
    Fp<P> x1, x2, y2;
#pragma unroll
    for (int i = 0; i < P::N; i++) {
        x1.l[i] = odd ? a_oth.l[i] : a_own.l[i];
        x2.l[i] = odd ? a_own.l[i] : a_oth.l[i];
        y2.l[i] = odd ? b_oth.l[i] : n5.l[i];
    }
    return fp_mul2(x1, b_own, x2, y2);
}

// square, phase 1: even lane (a0 + a1)(a0 - 5 a1), odd lane a0 a1
template <class P>
SS_HD Fp<P> fp2l_sqr_lane1(int odd, const Fp<P>& own, const Fp<P>& oth) {
    const Fp<P> s = fp_add_nr(own, oth);              // < 2p, multiplier operand only
    const Fp<P> d = fp_sub(own, fp_mul5(oth));
    Fp<P> x, y;
#pragma unroll
    for (int i = 0; i < P::N; i++) {
        x.l[i] = odd ? oth.l[i] : s.l[i];
        y.l[i] = odd ? own.l[i] : d.l[i];
    }
    return fp_mul(x, y);
}
// square, phase 2: even lane t + 4 v (v = partner's product), odd lane 2 v (v = own product)
template <class P>
SS_HD Fp<P> fp2l_sqr_lane2(int odd, const Fp<P>& prod_own, const Fp<P>& prod_oth) {
    const Fp<P> e = fp_add(prod_own, fp_dbl(fp_dbl(prod_oth)));
    const Fp<P> o = fp_dbl(prod_own);
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.l[i] = odd ? o.l[i] : e.l[i];
    return r;
}

#if defined(__CUDACC__)
// ---- the lane-split element (device only; the host bodies exist so the generic __host__ __device__ templates of
//      ec.cuh / glv.cuh compile, they are never executed) ---------------------------------------------------------
// CONVERGED-WARP DISCIPLINE: every exchange is a full-mask __shfl_xor_sync, so all 32 lanes of a warp must execute the
// same sequence of lane-split operations.  (A first version exchanged over the 2-lane pair mask, which let pairs
// diverge, but ptxas brackets every such shuffle with WARPSYNC.COLLECTIVE / BSSY / BSYNC — 1 814 of them in
// k_scalar_mul_pair — and the kernel ran 44 % SLOWER than the per-thread one; profiles/r02_ab_variants.md.)  Kernels
// therefore keep out-of-range lanes alive on a clamped index, replace data-dependent branches around group operations
// by compute-and-select, and enter the rare exceptional paths (P + P, tiny-order bases, off-curve inputs) warp-wide
// behind __any_sync.
constexpr unsigned kFullMask = 0xffffffffu;

SS_HD int lane_odd() {
#if defined(__CUDA_ARCH__)
    return (int)(threadIdx.x & 1u);
#else
    return 0;
#endif
}

template <class P>
SS_HD Fp<P> pair_swap(const Fp<P>& v) {
#if defined(__CUDA_ARCH__)
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.l[i] = __shfl_xor_sync(kFullMask, v.l[i], 1);
    return r;
#else
    return v;
#endif
}
SS_HD bool pair_all(bool f) {
#if defined(__CUDA_ARCH__)
    return (__shfl_xor_sync(kFullMask, f ? 1 : 0, 1) != 0) && f;
#else
    return f;
#endif
}
SS_HD bool warp_any(bool f) {
#if defined(__CUDA_ARCH__)
    return __any_sync(kFullMask, f) != 0;
#else
    return f;
#endif
}

// One out-of-line function per operation and field, operands in registers like fp_mul_call: the exchange, the
// operand selection and the multiplier body live INSIDE the call, so a lane-split product marshals 24 + 12 registers
// (as a G1 multiplication does) instead of the 48 + 12 of a bare fp_mul2 call.
template <class P>
__device__ __noinline__ Fp<P> fp2l_mul_call(Fp<P> a, Fp<P> b) {
    const int odd = lane_odd();
    const Fp<P> ao = pair_swap(a), bo = pair_swap(b);
    const Fp<P> n5 = fp_neg(fp_mul5(bo));
    Fp<P> x1, x2, y2;
#pragma unroll
    for (int i = 0; i < P::N; i++) {
        x1.l[i] = odd ? ao.l[i] : a.l[i];
        x2.l[i] = odd ? a.l[i] : ao.l[i];
        y2.l[i] = odd ? bo.l[i] : n5.l[i];
    }
    return fp_mul2_inl(x1, b, x2, y2);
}
template <class P>
__device__ __noinline__ Fp<P> fp2l_sqr_call(Fp<P> a) {
    const int odd = lane_odd();
    const Fp<P> oth = pair_swap(a);
    const Fp<P> s = fp_add_nr(a, oth);
    const Fp<P> d = fp_sub(a, fp_mul5(oth));
    Fp<P> x, y;
#pragma unroll
    for (int i = 0; i < P::N; i++) {
        x.l[i] = odd ? oth.l[i] : s.l[i];
        y.l[i] = odd ? a.l[i] : d.l[i];
    }
    const Fp<P> prod = fp_mul_inl(x, y);
    return fp2l_sqr_lane2<P>(odd, prod, pair_swap(prod));
}

template <class P>
struct Fp2L {
    using Base = Fp<P>;
    using Params = P;
    // group operations stay inlined like the 12-limb G1 kernels: per lane the state IS a 12-limb field element
    static constexpr bool CALL_GROUP_OPS = false;
    Base h;  // c0 in the even lane, c1 in the odd lane

    SS_HD static Fp2L zero() { return Fp2L{Base::zero()}; }
    SS_HD static Fp2L one() {
        Fp2L r;
        const Base o = Base::one();
        const int odd = lane_odd();
#pragma unroll
        for (int i = 0; i < P::N; i++) r.h.l[i] = odd ? 0u : o.l[i];
        return r;
    }
    SS_HD bool is_zero() const { return pair_all(h.is_zero()); }
    SS_HD bool operator==(const Fp2L& o) const { return pair_all(h == o.h); }
    SS_HD bool operator!=(const Fp2L& o) const { return !(*this == o); }
};

template <class P>
SS_HD Fp2L<P> fp_add(const Fp2L<P>& a, const Fp2L<P>& b) { return Fp2L<P>{fp_add(a.h, b.h)}; }
template <class P>
SS_HD Fp2L<P> fp_sub(const Fp2L<P>& a, const Fp2L<P>& b) { return Fp2L<P>{fp_sub(a.h, b.h)}; }
template <class P>
SS_HD Fp2L<P> fp_neg(const Fp2L<P>& a) { return Fp2L<P>{fp_neg(a.h)}; }
template <class P>
SS_HD Fp2L<P> fp_dbl(const Fp2L<P>& a) { return Fp2L<P>{fp_dbl(a.h)}; }
template <class P>
SS_HD Fp2L<P> fp_mul(const Fp2L<P>& a, const Fp2L<P>& b) {
#if defined(__CUDA_ARCH__)
    return Fp2L<P>{fp2l_mul_call<P>(a.h, b.h)};
#else
    return Fp2L<P>{fp2l_mul_lane<P>(lane_odd(), a.h, pair_swap(a.h), b.h, pair_swap(b.h))};
#endif
}
template <class P>
SS_HD Fp2L<P> fp_sqr(const Fp2L<P>& a) {
#if defined(__CUDA_ARCH__)
    return Fp2L<P>{fp2l_sqr_call<P>(a.h)};
#else
    const int odd = lane_odd();
    const Fp<P> prod = fp2l_sqr_lane1<P>(odd, a.h, pair_swap(a.h));
    return Fp2L<P>{fp2l_sqr_lane2<P>(odd, prod, pair_swap(prod))};
#endif
}
template <class P>
SS_HD Fp2L<P> fp_mul_base(const Fp2L<P>& a, const Fp<P>& k) { return Fp2L<P>{fp_mul(a.h, k)}; }
// conjugate: c1 -> -c1
template <class P>
SS_HD Fp2L<P> fp_conj(const Fp2L<P>& a) {
    const Fp<P> n = fp_neg(a.h);
    Fp2L<P> r;
    const int odd = lane_odd();
#pragma unroll
    for (int i = 0; i < P::N; i++) r.h.l[i] = odd ? n.l[i] : a.h.l[i];
    return r;
}
// 1/(a0 + a1 u) = (a0 - a1 u) / (a0^2 + 5 a1^2): both lanes compute the same norm inverse (one Fermat power each)
template <class P>
SS_HD Fp2L<P> fp_inv(const Fp2L<P>& a) {
    const Fp<P> own2 = fp_sqr(a.h);
    const Fp<P> oth2 = pair_swap(own2);
    const int odd = lane_odd();
    // norm = c0^2 + 5 c1^2
    Fp<P> c0s, c1s;
#pragma unroll
    for (int i = 0; i < P::N; i++) {
        c0s.l[i] = odd ? oth2.l[i] : own2.l[i];
        c1s.l[i] = odd ? own2.l[i] : oth2.l[i];
    }
    const Fp<P> ni = fp_inv(fp_add(c0s, fp_mul5(c1s)));
    return fp_conj(Fp2L<P>{fp_mul(a.h, ni)});
}
#endif  // __CUDACC__

}  // namespace ss
