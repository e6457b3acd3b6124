// same_ratio / check_same_ratio on the device (setup-utils/src/helpers.rs:406-424).
//
// The reference compares two optimal-ate pairings, E::pairing(g1.0, g2.1) == E::pairing(g1.1, g2.0).  Only the
// verdict crosses the boundary and every non-degenerate bilinear pairing on G1 x G2 gives the same one, so this
// kernel evaluates the reduced TATE pairing product
//        ( f_{r,g1.0}(psi(g2.1)) * f_{r,-g1.1}(psi(g2.0)) ) ^ ((q^k - 1)/r)  ==  1
// which needs nothing but the G1 group law already in ec.cuh and one extension-field multiplication:
//   * Fq^k = F[w]/(w^6 - xi) with F = Fq2, xi = u (BLS12-377: ark's Fq12 = Fq2[v]/(v^3-u)[w]/(w^2-v)) or F = Fq,
//     xi = -4 (BW6-761: Fq6 = Fq[v]/(v^3+4)[w]/(w^2-v)); psi = untwist, (x', y') -> (x' w^2, y' w^3) for the D-type
//     twist of BLS12-377 and (x' w^4/xi, y' w^3/xi) for the M-type twist of BW6-761.
//   * One WARP per ratio check, lane k (mod 6) owns coefficient k of the running value: a product is 6
//     F-multiplications per lane with the operands exchanged by warp shuffles, the sparse line multiplication 3.
//     The G1 Miller points (Jacobian, shared doubling/addition formulas) are computed redundantly by every lane.
//   * Lines are scaled by Fq factors (killed by the final exponentiation, as are the vertical lines):
//       tangent at T = (X, Y, Z):      (3X^3 - 2Y^2)  -  3X^2 Z^2 * x_Q  +  2YZ^3 * y_Q
//       chord through T and P:         (N x_P - D y_P) -  N * x_Q        +  D * y_Q,   N = Y - y_P Z^3, D = Z (X - x_P Z^2)
//   * Final exponentiation: f^(q^(k/2)-1) = conj(f)/f (norm to the cubic subfield, one F inversion), then
//     f^(Q+1) with Q = q^2 (k = 12) / q (k = 6) through the Frobenius multipliers zeta^k (in Fq), then the
//     hard part Phi_k(q)/r by square-and-multiply (1255 / 1144 bits; exponent from tools/gen_constants.py).
//   * BLS12-377 takes two shortcuts with the same verdict (Bls377Pairing::ATE / FROB4): the Miller loop is the ate loop
//     over u (64 bits, Miller points on the twist over Fq2, lines at positions 1, w, w^3 — see miller_ate) instead of
//     the Tate loop over r (253 bits), and the hard part is a 4-way simultaneous exponentiation over the base-q
//     digits of the exponent (314 bits: f^(q^i) by the q-power Frobenius c_k -> conj(c_k) gamma^k).  BW6-761 keeps
//     the Tate loop and shortens the hard part to a 2-way exponentiation f^a0 (f^q)^a1 (572 bits, FROB2).
// O(1) work per verification (4 checks per response): latency, not throughput, is what matters here.
#pragma once
#include "codec.cuh"

namespace ss {

template <class P>
SS_D Fp<P> lane_get(const Fp<P>& a, int src) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.l[i] = __shfl_sync(0xffffffffu, a.l[i], src);
    return r;
}
template <class P>
SS_D Fp2<P> lane_get(const Fp2<P>& a, int src) {
    return Fp2<P>{lane_get(a.c0, src), lane_get(a.c1, src)};
}
template <class P>
SS_D Fp<P> fscale(const Fp<P>& a, const Fp<P>& s) {
    return fp_mul(a, s);
}
template <class P>
SS_D Fp2<P> fscale(const Fp2<P>& a, const Fp<P>& s) {
    return Fp2<P>{fp_mul(a.c0, s), fp_mul(a.c1, s)};
}

struct Bls377Pairing {
    using G1 = Bls377G1;
    using G2 = Bls377G2;
    using Fq = Fp<Bls377Fq>;
    using F = Fp2<Bls377Fq>;
    using PP = Bls377PairingParams;
    static constexpr int XPOS = 2;               // D-type: x_Q = x' w^2, y_Q = y' w^3
    static constexpr bool XI_ON_CONST = false;
    static constexpr bool ATE = true;            // Miller loop over u on the twist (64 bits) instead of r on G1 (253)
#if defined(SS_PAIRING_FROB4)
    static constexpr bool UCHAIN = false, FROB4 = true;  // A/B: the round-1 hard part (4-way simultaneous exponentiation)
#else
    static constexpr bool UCHAIN = true, FROB4 = false;  // hard part along the u-chain, see k_same_ratio
#endif
    static constexpr bool FROB2 = false;
    SS_D static F mul_xi(const F& a) { return F{fp_neg(fp_mul5(a.c1)), a.c0}; }  // xi = u, u^2 = -5
};
struct Bw6Pairing {
    using G1 = Bw6G1;
    using G2 = Bw6G2;
    using Fq = Fp<Bw6Fq>;
    using F = Fp<Bw6Fq>;
    using PP = Bw6PairingParams;
    static constexpr int XPOS = 4;               // M-type: xi x_Q = x' w^4, xi y_Q = y' w^3 (whole line scaled by xi)
    static constexpr bool XI_ON_CONST = true;
    static constexpr bool ATE = false;
    static constexpr bool UCHAIN = false, FROB4 = false;
    static constexpr bool FROB2 = true;  // hard part as f^a0 * (f^q)^a1 with the Gauss-reduced (a0, a1), 572 bits
    SS_D static F mul_xi(const F& a) { return fp_neg(fp_dbl(fp_dbl(a))); }  // xi = -4
};

// Lane layout (one warp per check): lane = 6 s + k (mod 18) — every lane holds coefficient k = lane % 6 of the running
// value, and the THREE lanes (k, s = 0..2) share the work of that coefficient: a product needs 6 F-multiplications per
// coefficient, each lane does 2 and the partial sums are added over lanes k, k + 6, k + 12 (every lane adds the same
// three values in the same order, so the copies stay identical); a sparse line multiplication is 1 F-multiplication per
// lane.  (Round 1 had every lane do all 6: the check is pure latency — one warp on an SM — so the fewer multiplications
// in sequence, the better: 18.7 -> see profiles/r02_pairing_latency.jsonl.)
SS_D int lane_k() { return (int)(threadIdx.x % 6u); }
SS_D int lane_s() { return (int)((threadIdx.x / 6u) % 3u); }

template <class F>
SS_D F ext_reduce3(const F& part, int k) {
    return fp_add(fp_add(lane_get(part, k), lane_get(part, k + 6)), lane_get(part, k + 12));
}

// c = a * b in F[w]/(w^6 - xi); every lane passes its own coefficient (lane k <-> w^k)
template <class C>
SS_D typename C::F ext_mul(const typename C::F& a, const typename C::F& b, int k) {
    using F = typename C::F;
    const int s = lane_s();
    F lo = F::zero(), hi = F::zero();  // hi collects the terms with i + j >= 6 (one multiplication by xi at the end)
#pragma unroll 1
    for (int ii = 0; ii < 2; ii++) {
        const int i = 2 * s + ii;
        int j = k - i;
        const bool wrap = j < 0;
        if (wrap) j += 6;
        F t = fp_mul(lane_get(a, i), lane_get(b, j));
        // per-lane selection instead of a branch: the lanes of a warp take different (i, j)
        F z = F::zero();
        hi = fp_add(hi, wrap ? t : z);
        lo = fp_add(lo, wrap ? z : t);
    }
    return ext_reduce3(fp_add(lo, C::mul_xi(hi)), k);
}

// f * (s0 + lx w^XPOS + ly w^3), s0 in Fq (times xi for the M-type twist), lx, ly in F: one term per lane group s
template <class C>
SS_D typename C::F ext_mul_line(const typename C::F& f, const typename C::Fq& s0, const typename C::F& lx,
                                const typename C::F& ly, int k) {
    using F = typename C::F;
    const int s = lane_s();
    // s = 0: f_k * s0 (* xi)   s = 1: f_{k-3} * ly   s = 2: f_{k-XPOS} * lx
    const int off = s == 1 ? 3 : (s == 2 ? C::XPOS : 0);
    int j = k - off;
    const bool wrap = j < 0;
    if (wrap) j += 6;
    const F fj = lane_get(f, j);
    F t0 = fscale(fj, s0);
    if (C::XI_ON_CONST) t0 = C::mul_xi(t0);
    F m = s == 1 ? ly : lx;
    F t12 = fp_mul(fj, m);
    F t12x = C::mul_xi(t12);
    F part;
    if (s == 0) part = t0;
    else part = wrap ? t12x : t12;
    return ext_reduce3(part, k);
}

// f * (c0 + c1 w + c3 w^3), all three in F — the lines of the ate Miller loop on a D-type twist; one term per lane group
template <class C>
SS_D typename C::F ext_mul_013(const typename C::F& f, const typename C::F& c0, const typename C::F& c1,
                               const typename C::F& c3, int k) {
    using F = typename C::F;
    const int s = lane_s();
    const int off = s == 1 ? 1 : (s == 2 ? 3 : 0);
    int j = k - off;
    const bool wrap = j < 0;
    if (wrap) j += 6;
    const F c = s == 0 ? c0 : (s == 1 ? c1 : c3);
    F t = fp_mul(lane_get(f, j), c);
    F tx = C::mul_xi(t);
    return ext_reduce3(wrap ? tx : t, k);
}

// The two pairings of a check are advanced by different lane groups: lanes with s == 1 carry the Miller point of
// pairing 1, all others that of pairing 0; each lane computes the line of ITS pairing only and the coefficients are
// broadcast from lane 0 (pairing 0) and lane 6 (pairing 1) for the two sparse multiplications every lane takes part in.
SS_D int lane_pairing() { return lane_s() == 1 ? 1 : 0; }
SS_D int lane_of_pairing(int j) { return 6 * j; }

// Tate: f_{r,P0}(psi Q0) * f_{r,P1}(psi Q1), Miller points on G1 (see the header of this file for the lines)
template <class C>
SS_D typename C::F miller_tate(const Affine<typename C::Fq>* P, const Affine<typename C::F>* Q, int k) {
    using F = typename C::F;
    using Fq = typename C::Fq;
    using G1 = typename C::G1;
    const int my = lane_pairing();
    const Affine<Fq> Pm = P[my];
    const Affine<F> Qm = Q[my];
    Jac<Fq> T = Jac<Fq>{Pm.x, Pm.y, Fq::one()};
    F f = k == 0 ? F::one() : F::zero();
#pragma unroll 1
    for (int i = G1::GP::ORDER_BITS - 2; i >= 0; i--) {
        f = ext_mul<C>(f, f, k);
        {
            Fq X2 = fp_sqr(T.X), Y2 = fp_sqr(T.Y), Z2 = fp_sqr(T.Z);
            Fq X2_3 = fp_add(fp_dbl(X2), X2);
            Fq s0 = fp_sub(fp_mul(X2_3, T.X), fp_dbl(Y2));
            Fq sx = fp_neg(fp_mul(X2_3, Z2));
            Fq sy = fp_dbl(fp_mul(fp_mul(T.Y, T.Z), Z2));
            F lx = fscale(Qm.x, sx), ly = fscale(Qm.y, sy);
            T = jac_dbl(T);
#pragma unroll 1
            for (int j = 0; j < 2; j++) {
                const int src = lane_of_pairing(j);
                f = ext_mul_line<C>(f, lane_get(s0, src), lane_get(lx, src), lane_get(ly, src), k);
            }
        }
        const bool bit = (G1::GP::order(i >> 5) >> (i & 31)) & 1;
        if (bit && i != 0) {  // the last addition (T = -P) is a vertical line
            Fq Z2 = fp_sqr(T.Z);
            Fq N = fp_sub(T.Y, fp_mul(Pm.y, fp_mul(Z2, T.Z)));
            Fq D = fp_mul(T.Z, fp_sub(T.X, fp_mul(Pm.x, Z2)));
            Fq s0 = fp_sub(fp_mul(N, Pm.x), fp_mul(D, Pm.y));
            F lx = fscale(Qm.x, fp_neg(N)), ly = fscale(Qm.y, D);
            T = jac_madd(T, Pm);
#pragma unroll 1
            for (int j = 0; j < 2; j++) {
                const int src = lane_of_pairing(j);
                f = ext_mul_line<C>(f, lane_get(s0, src), lane_get(lx, src), lane_get(ly, src), k);
            }
        }
    }
    return f;
}

// ate (BLS12, D-type twist): f_{u,Q0}(P0) * f_{u,Q1}(P1), Miller points T on the twist E'(Fq2), lines evaluated
// at P in E(Fq) and scaled by Fq2 factors:
//   tangent at T = (X, Y, Z):  2YZ^3 y_P  -  3X^2 Z^2 x_P * w  +  (3X^3 - 2Y^2) * w^3
//   chord through T and Q:     D y_P      -  N x_P * w         +  (N x_Q - D y_Q) * w^3,  N = Y - y_Q Z^3, D = Z (X - x_Q Z^2)
template <class C>
SS_D typename C::F miller_ate(const Affine<typename C::Fq>* P, const Affine<typename C::F>* Q, int k) {
    using F = typename C::F;
    using Fq = typename C::Fq;
    using PP = typename C::PP;
    const int my = lane_pairing();
    const Affine<Fq> Pm = P[my];
    const Affine<F> Qm = Q[my];
    Jac<F> T = Jac<F>{Qm.x, Qm.y, F::one()};
    F f = k == 0 ? F::one() : F::zero();
#pragma unroll 1
    for (int i = PP::ATE_BITS - 2; i >= 0; i--) {
        f = ext_mul<C>(f, f, k);
        {
            F X2 = fp_sqr(T.X), Y2 = fp_sqr(T.Y), Z2 = fp_sqr(T.Z);
            F X2_3 = fp_add(fp_dbl(X2), X2);
            F c3 = fp_sub(fp_mul(X2_3, T.X), fp_dbl(Y2));
            F c1 = fscale(fp_neg(fp_mul(X2_3, Z2)), Pm.x);
            F c0 = fscale(fp_dbl(fp_mul(fp_mul(T.Y, T.Z), Z2)), Pm.y);
            T = jac_dbl(T);
#pragma unroll 1
            for (int j = 0; j < 2; j++) {
                const int src = lane_of_pairing(j);
                f = ext_mul_013<C>(f, lane_get(c0, src), lane_get(c1, src), lane_get(c3, src), k);
            }
        }
        if ((PP::ate(i >> 5) >> (i & 31)) & 1) {
            F Z2 = fp_sqr(T.Z);
            F N = fp_sub(T.Y, fp_mul(Qm.y, fp_mul(Z2, T.Z)));
            F D = fp_mul(T.Z, fp_sub(T.X, fp_mul(Qm.x, Z2)));
            F c3 = fp_sub(fp_mul(N, Qm.x), fp_mul(D, Qm.y));
            F c0 = fscale(D, Pm.y), c1 = fscale(fp_neg(N), Pm.x);
            T = jac_madd(T, Qm);
#pragma unroll 1
            for (int j = 0; j < 2; j++) {
                const int src = lane_of_pairing(j);
                f = ext_mul_013<C>(f, lane_get(c0, src), lane_get(c1, src), lane_get(c3, src), k);
            }
        }
    }
    return f;
}

// q-power Frobenius on the w-basis over Fq2: c_k -> conj(c_k) * gamma^k
template <class C, class P>
SS_D Fp2<P> frobenius_q(const Fp2<P>& c, int k) {
    using PP = typename C::PP;
    Fp2<P> g;
#pragma unroll
    for (int i = 0; i < P::N; i++) {
        g.c0.l[i] = PP::frobq(k, i);
        g.c1.l[i] = PP::frobq(k, P::N + i);
    }
    return fp_mul(Fp2<P>{c.c0, fp_neg(c.c1)}, g);
}
template <class C, class P>
SS_D Fp<P> frobenius_q(const Fp<P>& c, int) {
    return c;  // not used (FROB4 is an Fq2-tower option)
}

// verdict bits: 1 = same ratio, 2 = one of the four points is the identity (check_same_ratio rejects those)
template <class C>
__global__ void __launch_bounds__(32) k_same_ratio(const uint32_t* g1_pairs, const uint32_t* g2_pairs, int count,
                                                   int* verdict) {
    using F = typename C::F;
    using Fq = typename C::Fq;
    using G1 = typename C::G1;
    using G2 = typename C::G2;
    using PP = typename C::PP;
    const int b = blockIdx.x;
    if (b >= count) return;
    const int lane = threadIdx.x, k = lane % 6;
    Affine<Fq> P[2];
    Affine<F> Q[2];
    {
        Affine<Fq> a0, a1;
        Affine<F> b0, b1;
        const uint32_t* p1 = g1_pairs + (size_t)b * 2 * (G1::USIZE / 4);
        const uint32_t* p2 = g2_pairs + (size_t)b * 2 * (G2::USIZE / 4);
        int e = decode_point<G1>(p1, false, CHECK_NO, a0) | decode_point<G1>(p1 + G1::USIZE / 4, false, CHECK_NO, a1) |
                decode_point<G2>(p2, false, CHECK_NO, b0) | decode_point<G2>(p2 + G2::USIZE / 4, false, CHECK_NO, b1);
        if (e != ERR_OK) {  // non-canonical field bytes
            if (lane == 0) verdict[b] = -e;
            return;
        }
        const bool lhs_one = a0.inf || b1.inf, rhs_one = a1.inf || b0.inf;
        if (lhs_one || rhs_one) {  // e(O, .) = 1 and e(P, Q) != 1 for non-zero subgroup points
            if (lane == 0) verdict[b] = 2 | ((lhs_one && rhs_one) ? 1 : 0);
            return;
        }
        P[0] = a0;
        Q[0] = b1;
        P[1] = affine_neg(a1);
        Q[1] = b0;
    }
    F f;
    if constexpr (C::ATE) f = miller_ate<C>(P, Q, k);
    else f = miller_tate<C>(P, Q, k);
    // ---- final exponentiation ----
    F fc = (k & 1) ? fp_neg(f) : f;  // f^(q^(k/2)): w -> -w
    F nrm = ext_mul<C>(f, fc, k);    // in the cubic subfield F[v]/(v^3 - xi), v = w^2: coefficients 0, 2, 4
    F n0 = lane_get(nrm, 0), n1 = lane_get(nrm, 2), n2 = lane_get(nrm, 4);
    F t0 = fp_sub(fp_sqr(n0), C::mul_xi(fp_mul(n1, n2)));
    F t1 = fp_sub(C::mul_xi(fp_sqr(n2)), fp_mul(n0, n1));
    F t2 = fp_sub(fp_sqr(n1), fp_mul(n0, n2));
    F d = fp_add(fp_mul(n0, t0), C::mul_xi(fp_add(fp_mul(n2, t1), fp_mul(n1, t2))));
    F di = fp_inv(d);
    F ninv = k == 0 ? fp_mul(t0, di) : k == 2 ? fp_mul(t1, di) : k == 4 ? fp_mul(t2, di) : F::zero();
    F finv = ext_mul<C>(fc, ninv, k);
    F f1 = ext_mul<C>(fc, finv, k);  // f^(q^(k/2) - 1)
    Fq z;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) z.l[i] = PP::frob(k, i);
    F f2 = ext_mul<C>(fscale(f1, z), f1, k);  // f1^(Q + 1): (sum c_k w^k)^Q = sum c_k zeta^k w^k
    F acc;
    if constexpr (C::UCHAIN) {
        // BLS12 hard part along the seed (Hayashida-Hayasaka-Teruya):  3 (q^4 - q^2 + 1)/r = (u-1)^2 (u+q) (u^2+q^2-1) + 3
        // (checked numerically for BLS12-377 in tools/gen_constants.py), i.e. five exponentiations by the 64-bit, weight-7
        // seed u — 315 squarings + 30 multiplications instead of the 314 + 295 of the 4-way simultaneous form.  f2 lies in
        // the cyclotomic subgroup, where x^-1 = x^(q^6) (w -> -w); the value obtained is f2^(3 hard), and gcd(3, r) = 1,
        // so "== 1" is the same verdict.
        auto inv = [&](const F& x) { return (k & 1) ? fp_neg(x) : x; };
        auto exp_u = [&](const F& x) {
            F r = x;
#pragma unroll 1
            for (int i = PP::ATE_BITS - 2; i >= 0; i--) {
                r = ext_mul<C>(r, r, k);
                if ((PP::ate(i >> 5) >> (i & 31)) & 1) r = ext_mul<C>(r, x, k);
            }
            return r;
        };
        F a = ext_mul<C>(exp_u(f2), inv(f2), k);                                  // f2^(u-1)
        a = ext_mul<C>(exp_u(a), inv(a), k);                                      // ^(u-1)
        F b2 = ext_mul<C>(exp_u(a), frobenius_q<C>(a, k), k);                     // ^(u+q)
        F c = ext_mul<C>(exp_u(exp_u(b2)), frobenius_q<C>(frobenius_q<C>(b2, k), k), k);
        c = ext_mul<C>(c, inv(b2), k);                                            // ^(u^2+q^2-1)
        acc = ext_mul<C>(c, ext_mul<C>(ext_mul<C>(f2, f2, k), f2, k), k);         // * f2^3
    } else if constexpr (C::FROB4) {
        // hard = sum_i h_i q^i (i < 4): f2^hard = prod_i (f2^(q^i))^(h_i), one squaring and at most one table
        // multiplication per bit of the (<= 314-bit) digits; tab[m] = prod_{i in m} f2^(q^i)
        F tab[16];
        tab[1] = f2;
        tab[2] = frobenius_q<C>(tab[1], k);
        tab[4] = frobenius_q<C>(tab[2], k);
        tab[8] = frobenius_q<C>(tab[4], k);
        tab[3] = ext_mul<C>(tab[1], tab[2], k);
#pragma unroll 1
        for (int m = 5; m < 8; m++) tab[m] = ext_mul<C>(tab[4], tab[m - 4], k);
#pragma unroll 1
        for (int m = 9; m < 16; m++) tab[m] = ext_mul<C>(tab[8], tab[m - 8], k);
        acc = k == 0 ? F::one() : F::zero();
#pragma unroll 1
        for (int i = PP::HARDQ_BITS - 1; i >= 0; i--) {
            acc = ext_mul<C>(acc, acc, k);
            int m = 0;
#pragma unroll
            for (int d = 0; d < 4; d++) m |= (int)((PP::hardq(d, i >> 5) >> (i & 31)) & 1) << d;
            if (m) acc = ext_mul<C>(acc, tab[m], k);
        }
    } else if constexpr (C::FROB2) {
        // a0 + a1 q = c * hard with gcd(c, r) = 1 (tools/gen_constants.py): f2^(a0 + a1 q) = 1 <=> f2^hard = 1, and
        // f2^q costs one Frobenius (the multipliers zeta^k already loaded in z: for k = 6 the Q of the easy part is q)
        F tab[4];
        tab[1] = f2;
        tab[2] = fscale(f2, z);
        tab[3] = ext_mul<C>(tab[1], tab[2], k);
        acc = k == 0 ? F::one() : F::zero();
#pragma unroll 1
        for (int i = PP::HARD2_BITS - 1; i >= 0; i--) {
            acc = ext_mul<C>(acc, acc, k);
            const int m = (int)((PP::hard2(0, i >> 5) >> (i & 31)) & 1) | ((int)((PP::hard2(1, i >> 5) >> (i & 31)) & 1) << 1);
            if (m) acc = ext_mul<C>(acc, tab[m], k);
        }
    } else {
        acc = f2;
#pragma unroll 1
        for (int i = PP::HARD_BITS - 2; i >= 0; i--) {
            acc = ext_mul<C>(acc, acc, k);
            if ((PP::hard(i >> 5) >> (i & 31)) & 1) acc = ext_mul<C>(acc, f2, k);
        }
    }
    const bool ok = k == 0 ? (acc == F::one()) : acc.is_zero();
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) verdict[b] = all_ok ? 1 : 0;
}

struct PairingOps {
    void (*same_ratio)(const uint32_t* g1_pairs, const uint32_t* g2_pairs, int count, int* verdict, cudaStream_t);
};
template <class C>
struct PairingLaunch {
    static void same_ratio(const uint32_t* g1_pairs, const uint32_t* g2_pairs, int count, int* verdict, cudaStream_t s) {
        if (count > 0) k_same_ratio<C><<<count, 32, 0, s>>>(g1_pairs, g2_pairs, count, verdict);
    }
    static PairingOps ops() { return PairingOps{&same_ratio}; }
};

const PairingOps& pairing_ops_bls377();
const PairingOps& pairing_ops_bw6();

}  // namespace ss
