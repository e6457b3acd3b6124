// Square roots for point decompression (ark-ec `get_ys_from_x_unchecked` at
// setup-utils/src/io/read.rs:63 when the input is compressed).
//
// BLS12-377 Fq has 2-adicity 46, the worst case for Tonelli-Shanks: the textbook loop costs ~1000
// data-dependent squarings per element and every lane of a warp waits for the slowest one.  Here
//   x = a^((t+1)/2), b = a^t = zeta^e                       one fixed-exponent power, 5-bit sliding windows
//   e recovered by Pohlig-Hellman in chunks [6,8,8,8,8,8]   40 squarings + 15 table multiplications,
//                                                           chunk lookups through a 1024-slot hash
//   sqrt(a) = x * zeta^(-e/2)                               6 table multiplications
// with uniform control flow.  a is a residue iff e is even.
// The Fq2 root (BLS12-377 G2) uses the norm method with ONE extra power: with alpha = sqrt(N(a)),
// delta = (a0+alpha)/2 and (w, x, e) the power data of delta:
//   e even:  c0 = x zeta^(-e/2),            c1 = a1/(2 c0) with 1/c0 = c0 * w^2 zeta^(-e)
//   e odd :  delta is a non-residue, so is -5, hence s = sqrt(-5 delta) = K x zeta^(-(e+f)/2) exists
//            (K = (-5)^((t+1)/2), zeta^f = (-5)^t) and the root of a is
//            c0 = (a1/2) s / delta,  c1 = -s/5          ((a0-alpha)/2 = -5 a1^2 / (4 delta)).
// Any root is acceptable: the caller orders (y, -y) canonically afterwards.
#pragma once
#include "constants_gen.cuh"
#include "fp2.cuh"
#include "sqrt_tables_gen.cuh"

namespace ss {

using FqB = Fp<Bls377Fq>;

SS_HD FqB sqrt_tab(const uint32_t* t, uint32_t j) {
    FqB r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = t[12 * j + i];
    return r;
}

// index of c in the order-256 subgroup <zeta^(2^38)>; -1 when c is not in it
SS_HD int sqrt_dlog8(const FqB& c) {
    const uint32_t key = c.l[0];
    uint32_t h = (key * kSqrtHashMul) >> (32 - kSqrtHashBits);
    for (int probe = 0; probe < (1 << kSqrtHashBits); probe++) {
        const uint32_t k = kSqrtDlogKey[h];
        if (k == key) return (int)kSqrtDlogVal[h];
        if (k == 0xffffffffu) return -1;
        h = (h + 1) & ((1u << kSqrtHashBits) - 1);
    }
    return -1;
}

// e with b = zeta^e (46 bits); false when b is not a 2^46-th root of unity power (never for b = a^t)
SS_HD bool sqrt_dlog46(const FqB& b, uint64_t& e_out) {
    FqB pw[5];  // pw[k] = b^(2^(8(k+1))): 2^8, 2^16, 2^24, 2^32, 2^40
    FqB c = b;
#pragma unroll 1
    for (int k = 0; k < 5; k++) {
#pragma unroll 1
        for (int s = 0; s < 8; s++) c = fp_sqr(c);
        pw[k] = c;
    }
    uint32_t d[6];
    int idx = sqrt_dlog8(pw[4]);
    if (idx < 0 || (idx & 3)) return false;
    d[0] = (uint32_t)idx >> 2;
    const uint32_t* W0[5] = {kSqrtW0_1, kSqrtW0_2, kSqrtW0_3, kSqrtW0_4, kSqrtW0_5};
    const uint32_t* V[4] = {kSqrtV1, kSqrtV2, kSqrtV3, kSqrtV4};
#pragma unroll 1
    for (int i = 1; i <= 5; i++) {
        c = i < 5 ? pw[4 - i] : b;  // b^(2^(40-8i))
        c = fp_mul(c, sqrt_tab(W0[i - 1], d[0]));
#pragma unroll 1
        for (int j = 1; j < i; j++) c = fp_mul(c, sqrt_tab(V[i - j - 1], d[j]));
        idx = sqrt_dlog8(c);
        if (idx < 0) return false;
        d[i] = (uint32_t)idx;
    }
    e_out = (uint64_t)d[0] | ((uint64_t)d[1] << 6) | ((uint64_t)d[2] << 14) | ((uint64_t)d[3] << 22) |
            ((uint64_t)d[4] << 30) | ((uint64_t)d[5] << 38);
    return true;
}

// zeta^(-E/2) for even E (taken mod 2^46; the result is defined up to sign)
SS_HD FqB sqrt_zeta_neg_half(uint64_t E) {
    E &= (1ull << 46) - 1;
    const uint32_t* S[5] = {kSqrtS1, kSqrtS2, kSqrtS3, kSqrtS4, kSqrtS5};
    FqB r = sqrt_tab(kSqrtS0, (uint32_t)(E & 63));
#pragma unroll 1
    for (int i = 1; i <= 5; i++) {
        const uint32_t di = (uint32_t)(E >> (6 + 8 * (i - 1))) & 255u;
        r = fp_mul(r, sqrt_tab(S[i - 1], di));
    }
    return r;
}

// w = a^((t-1)/2), x = a w, e = dlog(a^t).  a != 0.
SS_HD bool sqrt_parts(const FqB& a, FqB& w, FqB& x, uint64_t& e) {
    w = fp_pow_win<5>(a, [](int i) { return Bls377Fq::tm1h(i); }, 12);  // 5-bit sliding windows: 331 squarings + 71 multiplications
    x = fp_mul(a, w);
    FqB b = fp_mul(x, w);
    return sqrt_dlog46(b, e);
}

SS_HD bool fp_sqrt_fast(const FqB& a, FqB& out) {
    if (a.is_zero()) {
        out = a;
        return true;
    }
    FqB w, x;
    uint64_t e;
    if (!sqrt_parts(a, w, x, e)) return false;
    if (e & 1) return false;
    out = fp_mul(x, sqrt_zeta_neg_half(e));
    // the root comes out of lookup tables: one squaring proves it (a corrupted table entry or a hash collision
    // must surface as "no root", never as an off-curve point that CheckForCorrectness::No would let through)
    return fp_sqr(out) == a;
}

SS_HD bool fp2_sqrt_fast(const Fp2<Bls377Fq>& a, Fp2<Bls377Fq>& out) {
    using F2 = Fp2<Bls377Fq>;
    if (a.c1.is_zero()) {
        FqB r;
        if (fp_sqrt_fast(a.c0, r)) {
            out = F2{r, FqB::zero()};
            return true;
        }
        // a0 is a non-residue: root = c1 u with c1^2 = a0 / -5
        FqB inv5 = sqrt_tab(kSqrtInv5, 0);
        if (!fp_sqrt_fast(fp_neg(fp_mul(a.c0, inv5)), r)) return false;
        out = F2{FqB::zero(), r};
        return true;
    }
    FqB norm = fp_add(fp_sqr(a.c0), fp_mul5(fp_sqr(a.c1)));
    FqB alpha;
    if (!fp_sqrt_fast(norm, alpha)) return false;
    FqB half;
#pragma unroll
    for (int i = 0; i < 12; i++) half.l[i] = Bls377Fq::half(i);
    FqB delta = fp_mul(fp_add(a.c0, alpha), half);
    if (delta.is_zero()) delta = fp_mul(fp_sub(a.c0, alpha), half);  // a0 = -alpha cannot happen with a1 != 0
    FqB w, x;
    uint64_t e;
    if (!sqrt_parts(delta, w, x, e)) return false;
    const FqB a1h = fp_mul(a.c1, half);
    if ((e & 1) == 0) {
        FqB h = sqrt_zeta_neg_half(e);
        FqB c0 = fp_mul(x, h);
        FqB dinv = fp_mul(fp_sqr(w), fp_sqr(h));  // 1/delta = w^2 zeta^(-e)
        out = F2{c0, fp_mul(a1h, fp_mul(c0, dinv))};
    } else {
        FqB h = sqrt_zeta_neg_half(e + kSqrtM5Dlog);
        FqB s = fp_mul(fp_mul(sqrt_tab(kSqrtM5K, 0), x), h);                       // sqrt(-5 delta)
        FqB dinv = fp_mul(fp_mul(fp_sqr(w), fp_sqr(h)), sqrt_tab(kSqrtM5B, 0));    // 1/delta
        out = F2{fp_mul(a1h, fp_mul(s, dinv)), fp_neg(fp_mul(s, sqrt_tab(kSqrtInv5, 0)))};
    }
    return fp_sqr(out) == a;  // see fp_sqrt_fast
}

// dispatch used by the codec: table-driven for BLS12-377, windowed power for p = 3 mod 4
template <class P>
SS_HD bool fp_sqrt_any(const Fp<P>& a, Fp<P>& out) {
    if (a.is_zero()) {
        out = a;
        return true;
    }
    if (P::P3MOD4) {
        Fp<P> r = fp_pow_win<5>(a, [](int i) { return P::pp1q(i); }, P::N);
        out = r;
        return fp_sqr(r) == a;
    }
    return fp_sqrt(a, out);
}
SS_HD bool fp_sqrt_any(const FqB& a, FqB& out) { return fp_sqrt_fast(a, out); }
// MNT4-753 G2 (Fq2 = Fq[u]/(u^2 - 13)): generic norm method; MNT6-753 G2 (Fq3): Tonelli-Shanks in Fq3*
SS_HD bool fp_sqrt_any(const Fp2<Mnt753Q>& a, Fp2<Mnt753Q>& out) { return fp_sqrt(a, out) && fp_sqr(out) == a; }
SS_HD bool fp_sqrt_any(const Fp3<Mnt753R>& a, Fp3<Mnt753R>& out) {
    return fp3_sqrt<Mnt753R, Mnt6Fq3Sqrt>(a, out) && fp_sqr(out) == a;
}
SS_HD bool fp_sqrt_any(const Fp2<Bls377Fq>& a, Fp2<Bls377Fq>& out) { return fp2_sqrt_fast(a, out); }

}  // namespace ss
