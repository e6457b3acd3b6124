// Montgomery prime-field arithmetic on N 32-bit limbs held in registers (N even).
//
// Replaces, on the device, what the reference gets from ark-ff 0.4 `Fp<MontBackend<_, N>>`
// (call sites: setup-utils/src/helpers.rs:36,101,104; setup-utils/src/io/read.rs:63).
//
// Multiplication is operand scanning with interleaved reduction, written so that every
// 32x32->64 product lands on an aligned (even, odd) register pair: the accumulator is kept as two
// arrays, X holding the tiles at limb positions (0,1),(2,3).. and Y the tiles at (1,2),(3,4)..,
// i.e. value = X + Y * 2^32.  Each tile update is `mad.lo.cc` + `madc.hi.cc` on one pair, which
// ptxas turns into one IMAD.WIDE.U32.X; the carry runs along the whole array in one chain.  The
// per-iteration division by 2^32 is free: X and Y swap roles (see mont_step).
// Cost per multiplication: 2*N*N wide multiply-adds (N = 12: 279 IMAD.WIDE + 26 IMAD in SASS), DESIGN.md §3.
#pragma once
#include "ptx.cuh"

namespace ss {

template <class P>
struct Fp {
    static constexpr int N = P::N;
    using Params = P;
#ifndef SS_CALL_GROUP_OPS_LIMBS
#define SS_CALL_GROUP_OPS_LIMBS 12  // group ops are out-of-line for fields wider than this
#endif
    static constexpr bool CALL_GROUP_OPS = (N > SS_CALL_GROUP_OPS_LIMBS);
    uint32_t l[N];

    SS_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    SS_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r;
    }
    SS_HD bool is_zero() const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < N; i++) t |= l[i];
        return t == 0;
    }
    SS_HD bool operator==(const Fp& o) const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < N; i++) t |= l[i] ^ o.l[i];
        return t == 0;
    }
    SS_HD bool operator!=(const Fp& o) const { return !(*this == o); }
};

// -x computed as (x ^ 0xffffffff) + 1 with the mask read from constant memory: when ptxas can prove that a
// multiplier operand is a negation it folds the sign into the low product and stops fusing the
// mad.lo/madc.hi pairs into IMAD.WIDE (144 instead of 267 fused products per 12-limb multiplication).
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t kOpaqueOnes = 0xffffffffu;
#else
static const uint32_t kOpaqueOnes = 0xffffffffu;
#endif
SS_HD uint32_t opaque_neg(uint32_t x) { return (x ^ kOpaqueOnes) + 1u; }
#if defined(SS_NO_P0_TRICK)  // A/B switch
#define SS_P0_TRICK false
#else
#define SS_P0_TRICK true
#endif

// r = (a >= p) ? a - p : a      (a < 2p)
template <class P>
SS_HD void fp_final_sub(uint32_t* a) {
    constexpr int N = P::N;
    uint32_t t[N];
    t[0] = sub_cc(a[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = subc_cc(a[i], P::mod(i));
    uint32_t borrow = subc(0, 0);  // 0xffffffff when a < p
#pragma unroll
    for (int i = 0; i < N; i++) a[i] = borrow ? a[i] : t[i];
}

template <class P>
SS_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    r.l[N - 1] = addc(a.l[N - 1], b.l[N - 1]);  // p has spare top bits: no carry out
    fp_final_sub<P>(r.l);
    return r;
}

// a + b WITHOUT the conditional subtraction: the result is < 2p and may only feed fp_mul / fp_sqr, whose
// Montgomery reduction tolerates operands below 2p (the product before the final subtraction is below
// (4p^2 + R p)/R = p (1 + 4p/R) < 1.04 p for the moduli here, which leave >= 7 spare bits in R = 2^(32 N)).
template <class P>
SS_HD Fp<P> fp_add_nr(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    r.l[N - 1] = addc(a.l[N - 1], b.l[N - 1]);
    return r;
}

template <class P>
SS_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = subc(0, 0);
    // add back p under mask
    r.l[0] = add_cc(r.l[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], P::mod(i) & borrow);
    r.l[N - 1] = addc(r.l[N - 1], P::mod(N - 1) & borrow);
    return r;
}

template <class P>
SS_HD Fp<P> fp_neg(const Fp<P>& a) {
    constexpr int N = P::N;
    Fp<P> r;
    uint32_t nz = a.is_zero() ? 0u : 0xffffffffu;
    r.l[0] = sub_cc(P::mod(0), a.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(P::mod(i), a.l[i]);
    r.l[N - 1] = subc(P::mod(N - 1), a.l[N - 1]);
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] &= nz;
    return r;
}

template <class P>
SS_HD Fp<P> fp_dbl(const Fp<P>& a) {
    return fp_add(a, a);
}

// ---- Montgomery multiplication -------------------------------------------------------------
// X[k] sits at limb position k, Y[k] at position k+1.
// reduce: m = X[0]*INV ; (X,Y) += m*p  => X[0] == 0
template <class P>
SS_HD void mont_reduce_step(uint32_t* X, uint32_t* Y) {
    constexpr int N = P::N;
    // p = 1 mod 2^32 (both BLS12-377 moduli): -p^-1 = -1, so m = -X[0] and the limb-0 product m*p[0] = m
    // only turns X[0] into 0 with carry (X[0] != 0) — no multiplier instruction for either.
    constexpr bool P0_ONE = SS_P0_TRICK && (P::mod(0) == 1u);
    uint32_t m = P0_ONE ? opaque_neg(X[0]) : mul_lo(X[0], P::inv());
    // odd limbs of p into Y
    Y[0] = mad_lo_cc(P::mod(1), m, Y[0]);
    Y[1] = madc_hi_cc(P::mod(1), m, Y[1]);
#pragma unroll
    for (int j = 2; j < N - 2; j += 2) {
        Y[j] = madc_lo_cc(P::mod(j + 1), m, Y[j]);
        Y[j + 1] = madc_hi_cc(P::mod(j + 1), m, Y[j + 1]);
    }
    Y[N - 2] = madc_lo_cc(P::mod(N - 1), m, Y[N - 2]);
    Y[N - 1] = madc_hi(P::mod(N - 1), m, Y[N - 1]);
    // even limbs of p into X
    if (P0_ONE) {
        X[0] = add_cc(X[0], m);
        X[1] = addc_cc(X[1], 0);
    } else {
        X[0] = mad_lo_cc(P::mod(0), m, X[0]);
        X[1] = madc_hi_cc(P::mod(0), m, X[1]);
    }
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        X[j] = madc_lo_cc(P::mod(j), m, X[j]);
        X[j + 1] = madc_hi_cc(P::mod(j), m, X[j + 1]);
    }
    Y[N - 1] = addc(Y[N - 1], 0);
}

// One operand-scanning step for bi, entered with the previous step's X (now E, aligned k-1 after
// the implicit >>32, E[0] == 0) and Y (now aligned k).  On exit roles are (X=oldY, Y=oldX).
template <class P>
SS_HD void mont_step(uint32_t* X /*old Y*/, uint32_t* Y /*old X*/, const uint32_t* a, uint32_t bi) {
    constexpr int N = P::N;
    X[0] = add_cc(X[0], Y[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        Y[j] = madc_lo_cc(a[j + 1], bi, Y[j + 2]);
        Y[j + 1] = madc_hi_cc(a[j + 1], bi, Y[j + 3]);
    }
    Y[N - 2] = madc_lo_cc(a[N - 1], bi, 0);
    Y[N - 1] = madc_hi(a[N - 1], bi, 0);
    X[0] = mad_lo_cc(a[0], bi, X[0]);
    X[1] = madc_hi_cc(a[0], bi, X[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        X[j] = madc_lo_cc(a[j], bi, X[j]);
        X[j + 1] = madc_hi_cc(a[j], bi, X[j + 1]);
    }
    Y[N - 1] = addc(Y[N - 1], 0);
    mont_reduce_step<P>(X, Y);
}

template <class P>
SS_HD Fp<P> fp_mul_inl(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    uint32_t X[N], Y[N];
    // first step: plain products
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        X[j] = mul_lo(a.l[j], b.l[0]);
        X[j + 1] = mul_hi(a.l[j], b.l[0]);
        Y[j] = mul_lo(a.l[j + 1], b.l[0]);
        Y[j + 1] = mul_hi(a.l[j + 1], b.l[0]);
    }
    mont_reduce_step<P>(X, Y);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_step<P>(Y, X, a.l, b.l[i]);
        if (i + 1 < N) mont_step<P>(X, Y, a.l, b.l[i + 1]);
    }
    // after an odd number of role swaps: X-role = Y, Y-role = X.  result = Yrole + (Xrole >> 32)
    Fp<P> r;
    r.l[0] = add_cc(X[0], Y[1]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(X[j], Y[j + 1]);
    r.l[N - 1] = addc(X[N - 1], 0);
    fp_final_sub<P>(r.l);
    return r;
}

// ---- K independent products, row-interleaved --------------------------------------------------------
// r[k] = a[k] * b[k] / R for k < K, with the operand-scanning rows of the K products issued in turn.  One product's
// rows are one long dependency chain (carry into carry, reduction factor from limb 0); a warp that is alone on its
// scheduler (the Fq2 kernels: 255 registers, 2 warps per scheduler) issues 26 % of the cycles on it.  K chains side by
// side give ptxas independent work to fill the multiplier pipe from ONE warp.  Used by the out-of-line Fq2 units of
// fp2.cuh (K = 3: Karatsuba, K = 2: complex squaring).  Same operand bounds as fp_mul_inl.
template <class P, int K>
SS_HD void fp_mul_xk_inl(const Fp<P> (&a)[K], const Fp<P> (&b)[K], Fp<P> (&r)[K]) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    uint32_t X[K][N], Y[K][N];
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            X[k][j] = mul_lo(a[k].l[j], b[k].l[0]);
            X[k][j + 1] = mul_hi(a[k].l[j], b[k].l[0]);
            Y[k][j] = mul_lo(a[k].l[j + 1], b[k].l[0]);
            Y[k][j + 1] = mul_hi(a[k].l[j + 1], b[k].l[0]);
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) mont_reduce_step<P>(X[k], Y[k]);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
#pragma unroll
        for (int k = 0; k < K; k++) mont_step<P>(Y[k], X[k], a[k].l, b[k].l[i]);
        if (i + 1 < N) {
#pragma unroll
            for (int k = 0; k < K; k++) mont_step<P>(X[k], Y[k], a[k].l, b[k].l[i + 1]);
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        r[k].l[0] = add_cc(X[k][0], Y[k][1]);
#pragma unroll
        for (int j = 1; j < N - 1; j++) r[k].l[j] = addc_cc(X[k][j], Y[k][j + 1]);
        r[k].l[N - 1] = addc(X[k][N - 1], 0);
        fp_final_sub<P>(r[k].l);
    }
}

// ---- sum of two products ----------------------------------------------------------------------------
// (a*b + c*d) / R mod p with ONE interleaved reduction: 2 N^2 products + N^2 reduction instead of the 4 N^2 of two
// multiplications.  Used by the lane-split Fq2 arithmetic (fp2l.cuh), where one lane computes a0 b0 + a1 (-5 b1) and
// the other a0 b1 + a1 b0.  Operand bounds: a, c < 2p and b, d < 2^(32 N) with a*b + c*d < p*R (callers keep
// b, d <= 5p, i.e. a*b + c*d <= 20 p^2 < p*R since R/p > 2^7); every partial sum then fits the N+1 limbs of (X, Y):
// T + a*b_i + c*d_i + m*p < 2^32 (2p + 2p + p) + 2p, far below 2^(32 (N+1)).  The result is < 1.2 p before the final
// conditional subtraction.
// accumulate one more product row c * di into (X, Y) WITHOUT shifting (same shape as mont_reduce_step with p := c)
template <class P>
SS_HD void mont_row_acc(uint32_t* X, uint32_t* Y, const uint32_t* c, uint32_t di) {
    constexpr int N = P::N;
    Y[0] = mad_lo_cc(c[1], di, Y[0]);
    Y[1] = madc_hi_cc(c[1], di, Y[1]);
#pragma unroll
    for (int j = 2; j < N - 2; j += 2) {
        Y[j] = madc_lo_cc(c[j + 1], di, Y[j]);
        Y[j + 1] = madc_hi_cc(c[j + 1], di, Y[j + 1]);
    }
    Y[N - 2] = madc_lo_cc(c[N - 1], di, Y[N - 2]);
    Y[N - 1] = madc_hi(c[N - 1], di, Y[N - 1]);
    X[0] = mad_lo_cc(c[0], di, X[0]);
    X[1] = madc_hi_cc(c[0], di, X[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        X[j] = madc_lo_cc(c[j], di, X[j]);
        X[j + 1] = madc_hi_cc(c[j], di, X[j + 1]);
    }
    Y[N - 1] = addc(Y[N - 1], 0);
}

// the shifting product row of mont_step without its reduction
template <class P>
SS_HD void mont_row_shift(uint32_t* X /*old Y*/, uint32_t* Y /*old X*/, const uint32_t* a, uint32_t bi) {
    constexpr int N = P::N;
    X[0] = add_cc(X[0], Y[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        Y[j] = madc_lo_cc(a[j + 1], bi, Y[j + 2]);
        Y[j + 1] = madc_hi_cc(a[j + 1], bi, Y[j + 3]);
    }
    Y[N - 2] = madc_lo_cc(a[N - 1], bi, 0);
    Y[N - 1] = madc_hi(a[N - 1], bi, 0);
    X[0] = mad_lo_cc(a[0], bi, X[0]);
    X[1] = madc_hi_cc(a[0], bi, X[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        X[j] = madc_lo_cc(a[j], bi, X[j]);
        X[j + 1] = madc_hi_cc(a[j], bi, X[j + 1]);
    }
    Y[N - 1] = addc(Y[N - 1], 0);
}

template <class P>
SS_HD Fp<P> fp_mul2_inl(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    constexpr int N = P::N;
    uint32_t X[N], Y[N];
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        X[j] = mul_lo(a.l[j], b.l[0]);
        X[j + 1] = mul_hi(a.l[j], b.l[0]);
        Y[j] = mul_lo(a.l[j + 1], b.l[0]);
        Y[j + 1] = mul_hi(a.l[j + 1], b.l[0]);
    }
    mont_row_acc<P>(X, Y, c.l, d.l[0]);
    mont_reduce_step<P>(X, Y);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_row_shift<P>(Y, X, a.l, b.l[i]);
        mont_row_acc<P>(Y, X, c.l, d.l[i]);
        mont_reduce_step<P>(Y, X);
        if (i + 1 < N) {
            mont_row_shift<P>(X, Y, a.l, b.l[i + 1]);
            mont_row_acc<P>(X, Y, c.l, d.l[i + 1]);
            mont_reduce_step<P>(X, Y);
        }
    }
    Fp<P> r;
    r.l[0] = add_cc(X[0], Y[1]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(X[j], Y[j + 1]);
    r.l[N - 1] = addc(X[N - 1], 0);
    fp_final_sub<P>(r.l);
    return r;
}

#if defined(__CUDACC__)
template <class P>
__device__ __noinline__ Fp<P> fp_mul2_call(Fp<P> a, Fp<P> b, Fp<P> c, Fp<P> d) {
    return fp_mul2_inl(a, b, c, d);
}
#endif

// ---- dedicated squaring --------------------------------------------------------------------------
// a^2 = 2 * sum_{i<j} a_i a_j 2^(32(i+j)) + sum_i a_i^2 2^(64 i): N(N+1)/2 wide products instead of N^2,
// followed by a separate Montgomery reduction of the 2N-limb square (N^2 wide products): 222 instead
// of 288 IMAD.WIDE for N = 12.  Off-diagonal products are accumulated in two 2N-limb arrays so that
// every product again lands on an aligned register pair: E[k] sits at limb position k and takes the
// products with even i+j, O[k] sits at position k+1 and takes the odd ones.  Rows are processed in
// increasing i, so the limb after the end of each chain has only ever received carry bits and one
// `addc` closes the chain.

// reduction iteration with the >>32 folded into the odd chain (mont_step with a := p, b_i := m)
template <class P>
SS_HD void mont_redc_shift_step(uint32_t* X /*old Y*/, uint32_t* Y /*old X*/) {
    constexpr int N = P::N;
    constexpr bool P0_ONE = SS_P0_TRICK && (P::mod(0) == 1u);  // see mont_reduce_step
    uint32_t m = P0_ONE ? opaque_neg(X[0] + Y[1]) : mul_lo(X[0] + Y[1], P::inv());
    X[0] = add_cc(X[0], Y[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        Y[j] = madc_lo_cc(P::mod(j + 1), m, Y[j + 2]);
        Y[j + 1] = madc_hi_cc(P::mod(j + 1), m, Y[j + 3]);
    }
    Y[N - 2] = madc_lo_cc(P::mod(N - 1), m, 0);
    Y[N - 1] = madc_hi(P::mod(N - 1), m, 0);
    if (P0_ONE) {
        X[0] = add_cc(X[0], m);
        X[1] = addc_cc(X[1], 0);
    } else {
        X[0] = mad_lo_cc(P::mod(0), m, X[0]);
        X[1] = madc_hi_cc(P::mod(0), m, X[1]);
    }
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        X[j] = madc_lo_cc(P::mod(j), m, X[j]);
        X[j + 1] = madc_hi_cc(P::mod(j), m, X[j + 1]);
    }
    Y[N - 1] = addc(Y[N - 1], 0);
}

// Montgomery reduction of a 2N-limb value T < p * 2^(32N) (consumed in place): T / 2^(32N) mod p
template <class P>
SS_HD Fp<P> mont_redc_wide(uint32_t* E) {
    constexpr int N = P::N;
    // X aligned at 0 = T[0..N), Y = 0
    uint32_t* X = E;  // low half is consumed in place
    uint32_t Y[N];
#pragma unroll
    for (int k = 0; k < N; k++) Y[k] = 0;
    mont_reduce_step<P>(X, Y);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_redc_shift_step<P>(Y, X);
        if (i + 1 < N) mont_redc_shift_step<P>(X, Y);
    }
    // W = Yrole + (Xrole >> 32) with Xrole = Y, Yrole = X (odd number of swaps), then + T_hi
    Fp<P> r;
    r.l[0] = add_cc(X[0], Y[1]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(X[j], Y[j + 1]);
    r.l[N - 1] = addc(X[N - 1], 0);
    r.l[0] = add_cc(r.l[0], E[N]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(r.l[j], E[N + j]);
    r.l[N - 1] = addc(r.l[N - 1], E[2 * N - 1]);
    fp_final_sub<P>(r.l);
    return r;
}

// By default the multiplier is ONE out-of-line function per field and kernel: operands travel in
// registers (no local memory), the call costs ~50 MOVs that issue in the shadow of the IMAD pipe,
// and kernels stay a few thousand instructions long instead of hundreds of KB of inlined
// carry chains.  -DSS_MUL_INLINE restores full inlining (for A/B measurements).
#if defined(__CUDACC__)
template <class P>
__device__ __noinline__ Fp<P> fp_mul_call(Fp<P> a, Fp<P> b) {
    return fp_mul_inl(a, b);
}
#endif

// Host-emulation builds (tests/emul) can count the field multiplications / squarings an algorithm executes:
// ss_op_count[0][N] multiplications, [1][N] squarings on N-limb fields.  bench.py's "executed work" figure is
// pinned by tests/test_device_algos_emul.py::test_executed_work_per_scalar_mul with these counters.
#if defined(SS_COUNT_OPS) && !defined(__CUDACC__)
inline unsigned long long ss_op_count[3][32] = {};
#define SS_COUNT_OP(kind, n) (++ss_op_count[kind][n])
#else
#define SS_COUNT_OP(kind, n) ((void)0)
#endif

template <class P>
SS_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__) && !defined(SS_MUL_INLINE) && !defined(SS_MULONLY_INLINE)
    return fp_mul_call<P>(a, b);
#else
    SS_COUNT_OP(0, P::N);
    return fp_mul_inl(a, b);
#endif
}

template <class P>
SS_HD Fp<P> fp_mul2(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
#if defined(__CUDA_ARCH__) && !defined(SS_MUL_INLINE)
    return fp_mul2_call<P>(a, b, c, d);
#else
    SS_COUNT_OP(0, P::N);
    SS_COUNT_OP(2, P::N);  // [2][N]: how many of the counted multiplications were the 1.5x sum-of-two-products form
    return fp_mul2_inl(a, b, c, d);
#endif
}

template <class P>
SS_HD Fp<P> fp_sqr_inl(const Fp<P>& av) {
    constexpr int N = P::N;
    const uint32_t* a = av.l;
    uint32_t E[2 * N], O[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) E[k] = O[k] = 0;
#pragma unroll
    for (int i = 0; i < N - 1; i++) {
        // odd i+j: j = i+1, i+3, ...  -> O[p-1], O[p]
        {
            int last = -1;
#pragma unroll
            for (int j = i + 1; j < N; j += 2) {
                const int p = i + j;
                O[p - 1] = (j == i + 1) ? mad_lo_cc(a[i], a[j], O[p - 1]) : madc_lo_cc(a[i], a[j], O[p - 1]);
                O[p] = madc_hi_cc(a[i], a[j], O[p]);
                last = p;
            }
            if (last >= 0) O[last + 1] = addc(O[last + 1], 0);
        }
        // even i+j: j = i+2, i+4, ... -> E[p], E[p+1]
        {
            int last = -1;
#pragma unroll
            for (int j = i + 2; j < N; j += 2) {
                const int p = i + j;
                E[p] = (j == i + 2) ? mad_lo_cc(a[i], a[j], E[p]) : madc_lo_cc(a[i], a[j], E[p]);
                E[p + 1] = madc_hi_cc(a[i], a[j], E[p + 1]);
                last = p;
            }
            if (last >= 0) E[last + 2] = addc(E[last + 2], 0);
        }
    }
    // D = E + O * 2^32  (into E)
    E[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 2 * N - 1; k++) E[k] = addc_cc(E[k], O[k - 1]);
    E[2 * N - 1] = addc(E[2 * N - 1], O[2 * N - 2]);
    // T = 2 D
    E[0] = add_cc(E[0], E[0]);
#pragma unroll
    for (int k = 1; k < 2 * N - 1; k++) E[k] = addc_cc(E[k], E[k]);
    E[2 * N - 1] = addc(E[2 * N - 1], E[2 * N - 1]);
    // T += sum a_i^2 2^(64 i)
    E[0] = mad_lo_cc(a[0], a[0], E[0]);
    E[1] = madc_hi_cc(a[0], a[0], E[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) {
        E[2 * i] = madc_lo_cc(a[i], a[i], E[2 * i]);
        E[2 * i + 1] = madc_hi_cc(a[i], a[i], E[2 * i + 1]);
    }
    E[2 * N - 2] = madc_lo_cc(a[N - 1], a[N - 1], E[2 * N - 2]);
    E[2 * N - 1] = madc_hi(a[N - 1], a[N - 1], E[2 * N - 1]);
    return mont_redc_wide<P>(E);
}

#if defined(__CUDACC__)
template <class P>
__device__ __noinline__ Fp<P> fp_sqr_call(Fp<P> a) {
    return fp_sqr_inl(a);
}
#endif

template <class P>
SS_HD Fp<P> fp_sqr(const Fp<P>& a) {
#if defined(SS_NO_DEDICATED_SQR)
    return fp_mul(a, a);
#elif defined(__CUDA_ARCH__) && !defined(SS_MUL_INLINE) && !defined(SS_SQR_INLINE)
    return fp_sqr_call<P>(a);
#else
    SS_COUNT_OP(1, P::N);
    return fp_sqr_inl(a);
#endif
}

// Montgomery form <-> canonical integer
template <class P>
SS_HD Fp<P> fp_to_mont(const Fp<P>& a) {
    Fp<P> r2;
#pragma unroll
    for (int i = 0; i < P::N; i++) r2.l[i] = P::r2(i);
    return fp_mul(a, r2);
}

template <class P>
SS_HD Fp<P> fp_from_mont(const Fp<P>& a) {
    Fp<P> o = Fp<P>::zero();
    o.l[0] = 1;
    return fp_mul(a, o);
}

// a^e, e given as a limb accessor (runtime-indexed constant table).  Sliding windows of W bits over the exponent
// (odd powers a, a^3, ..., a^(2^W - 1) in a table): nbits squarings + ~nbits / (W + 1) + 2^(W-1) multiplications instead
// of the nbits / 2 of square-and-multiply — 377 + 83 instead of 377 + 188 for a Fermat inversion in Fq377.  The exponent
// is the same for every thread, so the window boundaries are too: no divergence.  0^e = 0, a^0 = 1.
template <class P, class E, int W = 4>
SS_HD Fp<P> fp_pow(const Fp<P>& a, E exp_limb, int nlimbs) {
    constexpr int T = 1 << (W - 1);
    auto bit = [&](int i) -> uint32_t { return (exp_limb(i >> 5) >> (i & 31)) & 1u; };
    int i = 32 * nlimbs - 1;
    while (i >= 0 && !bit(i)) i--;
    if (i < 0) return Fp<P>::one();
    Fp<P> tab[T];  // tab[k] = a^(2k + 1)
    tab[0] = a;
    {
        const Fp<P> a2 = fp_sqr(a);
#pragma unroll 1
        for (int k = 1; k < T; k++) tab[k] = fp_mul(tab[k - 1], a2);
    }
    Fp<P> r = a;
    bool started = false;
    while (i >= 0) {
        if (!bit(i)) {  // only reached once started: the scan begins on the leading one
            r = fp_sqr(r);
            i--;
            continue;
        }
        int j = i - W + 1;
        if (j < 0) j = 0;
        while (!bit(j)) j++;  // window [i .. j], odd value
        uint32_t v = 0;
        for (int k = i; k >= j; k--) v = (v << 1) | bit(k);
        if (started) {
#pragma unroll 1
            for (int k = 0; k < i - j + 1; k++) r = fp_sqr(r);
            r = fp_mul(r, tab[v >> 1]);
        } else {
            r = tab[v >> 1];
            started = true;
        }
        i = j - 1;
    }
    return r;
}

template <int W, class P, class E>
SS_HD Fp<P> fp_pow_win(const Fp<P>& a, E exp_limb, int nlimbs) {
    return fp_pow<P, E, W>(a, exp_limb, nlimbs);
}

template <class P>
SS_HD Fp<P> fp_inv(const Fp<P>& a) {  // Fermat; 0 -> 0
    return fp_pow<P>(a, [](int i) { return P::pm2(i); }, P::N);
}

// canonical-integer comparison of two RAW (non-Montgomery) limb arrays: a > b
template <int N>
SS_HD bool raw_gt(const uint32_t* a, const uint32_t* b) {
    // b - a borrows  <=>  a > b
    uint32_t t = sub_cc(b[0], a[0]);
#pragma unroll
    for (int i = 1; i < N; i++) t = subc_cc(b[i], a[i]);
    (void)t;
    return subc(0, 0) != 0;
}

template <class P>
SS_HD bool raw_ge_mod(const uint32_t* a) {  // a >= p
    uint32_t t = sub_cc(a[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < P::N; i++) t = subc_cc(a[i], P::mod(i));
    (void)t;
    return subc(0, 0) == 0;
}

// Square root.  Returns false when a is a non-residue.  Any root is acceptable: callers order
// (y, -y) canonically afterwards (ark-ec get_ys_from_x_unchecked).
template <class P>
SS_HD bool fp_sqrt(const Fp<P>& a, Fp<P>& out) {
    if (a.is_zero()) {
        out = a;
        return true;
    }
    if (P::P3MOD4) {
        Fp<P> r = fp_pow<P>(a, [](int i) { return P::pp1q(i); }, P::N);
        out = r;
        return fp_sqr(r) == a;
    } else {
        // Tonelli-Shanks, p - 1 = 2^s * t
        Fp<P> w = fp_pow<P>(a, [](int i) { return P::tm1h(i); }, P::N);  // a^((t-1)/2)
        Fp<P> x = fp_mul(a, w);                                           // a^((t+1)/2)
        Fp<P> b = fp_mul(x, w);                                           // a^t
        Fp<P> z;
#pragma unroll
        for (int i = 0; i < P::N; i++) z.l[i] = P::root_of_unity(i);
        const Fp<P> one = Fp<P>::one();
        int v = P::TWO_ADICITY;
        while (!(b == one)) {
            int k = 0;
            Fp<P> b2k = b;
            while (!(b2k == one)) {
                b2k = fp_sqr(b2k);
                k++;
                if (k == v) return false;  // non-residue
            }
            Fp<P> wj = z;
            for (int j = 0; j < v - k - 1; j++) wj = fp_sqr(wj);
            z = fp_sqr(wj);
            b = fp_mul(b, z);
            x = fp_mul(x, wj);
            v = k;
        }
        out = x;
        return true;
    }
}

}  // namespace ss
