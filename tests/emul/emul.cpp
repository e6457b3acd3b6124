// TEST INFRASTRUCTURE: host build of the device math headers (ptx.cuh's carry-flag emulation), so
// the limb-level algorithms can be checked against oracle/pyref.py on a box without a GPU.
// Never linked into the product library.
#define SS_COUNT_OPS 1  // fp.cuh: count multiplications / squarings (host emulation only)
#include <cstring>
#include "../../snark-setup_b200/csrc/codec.cuh"
#include "../../snark-setup_b200/csrc/glv.cuh"
#include "../../snark-setup_b200/csrc/fp2l.cuh"

using namespace ss;

template <class P>
static Fp<P> load_raw(const uint8_t* b) {
    Fp<P> r;
    memcpy(r.l, b, 4 * P::N);
    return r;
}
template <class P>
static void store_raw(uint8_t* b, const Fp<P>& a) {
    memcpy(b, a.l, 4 * P::N);
}

// op: 0 mul 1 add 2 sub 3 neg 4 inv 5 sqrt (returns 1 if no root) 6 sqr
template <class P>
static int fp_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    Fp<P> x = fp_to_mont(load_raw<P>(a)), y = fp_to_mont(load_raw<P>(b)), r;
    int rc = 0;
    switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_neg(x); break;
        case 4: r = fp_inv(x); break;
        case 5: rc = fp_sqrt_any(x, r) ? 0 : 1; if (rc) r = Fp<P>::zero(); break;
        case 6: r = fp_sqr(x); break;
        default: return -1;
    }
    store_raw<P>(out, fp_from_mont(r));
    return rc;
}

template <class P>
static int fp2_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    Fp2<P> x{fp_to_mont(load_raw<P>(a)), fp_to_mont(load_raw<P>(a + 4 * P::N))};
    Fp2<P> y{fp_to_mont(load_raw<P>(b)), fp_to_mont(load_raw<P>(b + 4 * P::N))};
    Fp2<P> r;
    int rc = 0;
    switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_neg(x); break;
        case 4: r = fp_inv(x); break;
        case 5: rc = fp_sqrt_any(x, r) ? 0 : 1; if (rc) r = Fp2<P>::zero(); break;
        case 6: r = fp_sqr(x); break;
        default: return -1;
    }
    store_raw<P>(out, fp_from_mont(r.c0));
    store_raw<P>(out + 4 * P::N, fp_from_mont(r.c1));
    return rc;
}

template <class G>
static int point_mul(const uint8_t* in, int in_compressed, int check, const uint8_t* scalar, int nbits,
                     uint8_t* out, int out_compressed) {
    using F = typename G::F;
    Affine<F> p;
    uint32_t w[64];
    memcpy(w, in, in_compressed ? G::CSIZE : G::USIZE);
    int e = decode_point<G>(w, in_compressed != 0, check, p);
    if (e) return e;
    uint32_t s[16] = {0};
    memcpy(s, scalar, (nbits + 7) / 8);
    Jac<F> r = jac_mul_bits<F>(p, [&](int i) { return s[i]; }, nbits);
    Affine<F> a;
    if (r.is_identity()) {
        a.inf = true;
        a.x = F::zero();
        a.y = F::zero();
    } else {
        a = jac_to_affine_with_zinv(r, fp_inv(r.Z));
    }
    uint32_t o[64];
    encode_point<G>(o, out_compressed != 0, a);
    memcpy(out, o, out_compressed ? G::CSIZE : G::USIZE);
    return 0;
}

// out = P + Q via jac_add / jac_madd (which: 0 madd, 1 add(jac,jac) with randomised Z)
template <class G>
static int point_add(const uint8_t* pa, const uint8_t* pb, int which, uint8_t* out) {
    using F = typename G::F;
    Affine<F> p, q;
    uint32_t w[64];
    memcpy(w, pa, G::USIZE);
    if (int e = decode_point<G>(w, false, CHECK_NO, p)) return e;
    memcpy(w, pb, G::USIZE);
    if (int e = decode_point<G>(w, false, CHECK_NO, q)) return e;
    Jac<F> jp = p.inf ? Jac<F>::identity() : Jac<F>{p.x, p.y, F::one()};
    Jac<F> r;
    if (which == 0) {
        // give jp a non-trivial Z first: (X*4, Y*8, 2)
        if (!p.inf) {
            F two = fp_dbl(F::one());
            jp.X = fp_mul(p.x, fp_sqr(two));
            jp.Y = fp_mul(p.y, fp_mul(fp_sqr(two), two));
            jp.Z = two;
        }
        r = jac_madd(jp, q);
    } else {
        Jac<F> jq = q.inf ? Jac<F>::identity() : Jac<F>{q.x, q.y, F::one()};
        if (!q.inf) {
            F three = fp_add(fp_dbl(F::one()), F::one());
            jq.X = fp_mul(q.x, fp_sqr(three));
            jq.Y = fp_mul(q.y, fp_mul(fp_sqr(three), three));
            jq.Z = three;
        }
        r = jac_add(jp, jq);
    }
    Affine<F> a;
    if (r.is_identity()) {
        a.inf = true;
        a.x = F::zero();
        a.y = F::zero();
    } else {
        a = jac_to_affine_with_zinv(r, fp_inv(r.Z));
    }
    uint32_t o[64];
    encode_point<G>(o, false, a);
    memcpy(out, o, G::USIZE);
    return 0;
}

// k*P through the endomorphism path (glv.cuh), uncompressed in/out, scalar = canonical LE bytes
template <class G>
static int point_mul_endo(const uint8_t* in, const uint8_t* scalar, uint8_t* out) {
    using F = typename G::F;
    Affine<F> p;
    uint32_t w[64];
    memcpy(w, in, G::USIZE);
    int e = decode_point<G>(w, false, CHECK_NO, p);
    if (e) return e;
    uint32_t k[12] = {0};
    memcpy(k, scalar, 4 * G::Fr::N);
    Jac<F> r = scalar_mul_endo<G>(p, k);
    Affine<F> a;
    if (r.is_identity()) {
        a.inf = true;
        a.x = F::zero();
        a.y = F::zero();
    } else {
        a = jac_to_affine_with_zinv(r, fp_inv(r.Z));
    }
    uint32_t o[64];
    encode_point<G>(o, false, a);
    memcpy(out, o, G::USIZE);
    return 0;
}

template <class G>
static int subgroup_both(const uint8_t* in) {
    using F = typename G::F;
    Affine<F> p;
    uint32_t w[64];
    memcpy(w, in, G::USIZE);
    int e = decode_point<G>(w, false, CHECK_NO, p);
    if (e) return -e;
    return (in_subgroup<G>(p) ? 1 : 0) | (in_subgroup_rmul<G>(p) ? 2 : 0);
}

template <class G>
static int count_scalar_mul(const uint8_t* in, const uint8_t* scalar, unsigned long long* counts) {
    using F = typename G::F;
    Affine<F> p;
    uint32_t w[64];
    memcpy(w, in, G::USIZE);
    int e = decode_point<G>(w, false, CHECK_NO, p);
    if (e) return e;
    uint32_t k[12] = {0};
    memcpy(k, scalar, 4 * G::Fr::N);
    memset(ss_op_count, 0, sizeof(ss_op_count));
    Jac<F> r = scalar_mul_endo<G>(p, k);
    memcpy(counts, ss_op_count, 2 * 32 * sizeof(unsigned long long));  // [0] multiplications, [1] squarings (the caller holds 2 x 32 counters)
    return r.is_identity() ? 1 : 0;
}

// ---- MNT4/6-753 (a != 0, Fq2 / Fq3 twists, byte-granular codec): point arithmetic through the device headers --------
// op 0: transcode (decode with `check`, re-encode)   op 1: scalar multiplication by the 95-byte canonical scalar
// op 2: P + Q through jac_madd (first operand given a non-trivial Z)   op 3: r * P == O ? (returns 1 / 0)
template <class G>
static int mnt_op(int op, const uint8_t* in, int in_c, int check, const uint8_t* arg, uint8_t* out, int out_c) {
    using F = typename G::F;
    Affine<F> p;
    int e = decode_point<G>(in, in_c != 0, check, p);
    if (e) return -e;
    Jac<F> r = p.inf ? Jac<F>::identity() : Jac<F>{p.x, p.y, F::one()};
    if (op == 1) {
        uint32_t k[24] = {0};
        memcpy(k, arg, 95);
        r = jac_mul_bits<F>(p, [&](int i) { return k[i]; }, 753);
    } else if (op == 2) {
        Affine<F> q;
        if ((e = decode_point<G>(arg, false, CHECK_NO, q))) return -e;
        if (!p.inf) {
            F two = fp_dbl(F::one());
            r.X = fp_mul(p.x, fp_sqr(two));
            r.Y = fp_mul(p.y, fp_mul(fp_sqr(two), two));
            r.Z = two;
        }
        r = jac_madd(r, q);
    } else if (op == 3) {
        return in_subgroup_rmul<G>(p) ? 1 : 0;
    }
    Affine<F> a;
    if (r.is_identity()) {
        a.inf = true;
        a.x = F::zero();
        a.y = F::zero();
    } else {
        a = jac_to_affine_with_zinv(r, fp_inv(r.Z));
    }
    encode_point<G>(out, out_c != 0, a);
    return 0;
}
extern "C" {
// counts[kind][limbs] (kind 0 = multiplications, 1 = squarings) executed by ONE scalar_mul_endo<G> (glv.cuh)
int emul_count_scalar_mul(int group, const uint8_t* in, const uint8_t* scalar, unsigned long long* counts) {
    switch (group) {
        case 0: return count_scalar_mul<Bls377G1>(in, scalar, counts);
        case 1: return count_scalar_mul<Bls377G2>(in, scalar, counts);
        case 2: return count_scalar_mul<Bw6G1>(in, scalar, counts);
        case 3: return count_scalar_mul<Bw6G2>(in, scalar, counts);
    }
    return -1;
}
// bit0: endomorphism test, bit1: r-multiplication
int emul_in_subgroup(int group, const uint8_t* in) {
    switch (group) {
        case 0: return subgroup_both<Bls377G1>(in);
        case 1: return subgroup_both<Bls377G2>(in);
        case 2: return subgroup_both<Bw6G1>(in);
        case 3: return subgroup_both<Bw6G2>(in);
    }
    return -100;
}
int emul_point_mul_endo(int group, const uint8_t* in, const uint8_t* scalar, uint8_t* out) {
    switch (group) {
        case 0: return point_mul_endo<Bls377G1>(in, scalar, out);
        case 1: return point_mul_endo<Bls377G2>(in, scalar, out);
        case 2: return point_mul_endo<Bw6G1>(in, scalar, out);
        case 3: return point_mul_endo<Bw6G2>(in, scalar, out);
    }
    return -1;
}
// field: 0 Bls377Fq, 1 Bls377Fr, 2 Bw6Fq, 3 Bls377Fq2
int emul_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    switch (field) {
        case 0: return fp_op<Bls377Fq>(op, a, b, out);
        case 1: return fp_op<Bls377Fr>(op, a, b, out);
        case 2: return fp_op<Bw6Fq>(op, a, b, out);
        case 3: return fp2_op<Bls377Fq>(op, a, b, out);
    }
    return -1;
}
// group: 0 bls g1, 1 bls g2, 2 bw6 g1, 3 bw6 g2
int emul_point_mul(int group, const uint8_t* in, int in_compressed, int check, const uint8_t* scalar, int nbits,
                   uint8_t* out, int out_compressed) {
    switch (group) {
        case 0: return point_mul<Bls377G1>(in, in_compressed, check, scalar, nbits, out, out_compressed);
        case 1: return point_mul<Bls377G2>(in, in_compressed, check, scalar, nbits, out, out_compressed);
        case 2: return point_mul<Bw6G1>(in, in_compressed, check, scalar, nbits, out, out_compressed);
        case 3: return point_mul<Bw6G2>(in, in_compressed, check, scalar, nbits, out, out_compressed);
    }
    return -1;
}
int emul_point_add(int group, const uint8_t* a, const uint8_t* b, int which, uint8_t* out) {
    switch (group) {
        case 0: return point_add<Bls377G1>(a, b, which, out);
        case 1: return point_add<Bls377G2>(a, b, which, out);
        case 2: return point_add<Bw6G1>(a, b, which, out);
        case 3: return point_add<Bw6G2>(a, b, which, out);
    }
    return -1;
}

// fp_mul2: (a*b + c*d) / R mod p on canonical inputs scaled as (a, c) < scale_ac * p and (b, d) < scale_bd * p is NOT
// needed here: the inputs are plain residues; `a2`/`c2` flags add p to a / c first (operands below 2p are allowed)
int emul_fp_mul2(const uint8_t* a, const uint8_t* b, const uint8_t* c, const uint8_t* d, int a_plus_p, int d_times5, uint8_t* out) {
    typedef Bls377Fq P;
    Fp<P> x = fp_to_mont(load_raw<P>(a)), y = fp_to_mont(load_raw<P>(b)), z = fp_to_mont(load_raw<P>(c)), w = fp_to_mont(load_raw<P>(d));
    if (a_plus_p) {  // unreduced representative x + p (< 2p): same residue, exercises the operand bound
        Fp<P> m;
        for (int i = 0; i < P::N; i++) m.l[i] = P::mod(i);
        x = fp_add_nr(x, m);
        z = fp_add_nr(z, m);
    }
    if (d_times5) {  // unreduced 5*w (< 5p) as the scanned operand
        Fp<P> t = fp_add_nr(w, w);
        t = fp_add_nr(t, t);
        w = fp_add_nr(t, w);
    }
    Fp<P> r = fp_mul2(x, y, z, w);
    store_raw<P>(out, fp_from_mont(r));
    return 0;
}

// out-of-line Fq2 units (fp2.cuh: row-interleaved Karatsuba product / complex square); op 0 = product, 1 = square
int emul_fp2_unit(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    typedef Bls377Fq P;
    Fp2<P> x{fp_to_mont(load_raw<P>(a)), fp_to_mont(load_raw<P>(a + 48))};
    Fp2<P> y{fp_to_mont(load_raw<P>(b)), fp_to_mont(load_raw<P>(b + 48))};
    Fp2<P> r = op == 0 ? fp2_mul_body(x, y) : fp2_sqr_body(x);
    store_raw<P>(out, fp_from_mont(r.c0));
    store_raw<P>(out + 48, fp_from_mont(r.c1));
    return 0;
}

// lane-split Fq2 product / square through the pure per-lane functions of fp2l.cuh: both lanes evaluated in turn
int emul_fp2l_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    typedef Bls377Fq P;
    Fp<P> a0 = fp_to_mont(load_raw<P>(a)), a1 = fp_to_mont(load_raw<P>(a + 48));
    Fp<P> b0 = fp_to_mont(load_raw<P>(b)), b1 = fp_to_mont(load_raw<P>(b + 48));
    Fp<P> r0, r1;
    if (op == 0) {
        r0 = fp2l_mul_lane<P>(0, a0, a1, b0, b1);
        r1 = fp2l_mul_lane<P>(1, a1, a0, b1, b0);
    } else {
        Fp<P> p0 = fp2l_sqr_lane1<P>(0, a0, a1), p1 = fp2l_sqr_lane1<P>(1, a1, a0);
        r0 = fp2l_sqr_lane2<P>(0, p0, p1);
        r1 = fp2l_sqr_lane2<P>(1, p1, p0);
    }
    store_raw<P>(out, fp_from_mont(r0));
    store_raw<P>(out + 48, fp_from_mont(r1));
    return 0;
}

int emul_mnt_op(int group, int op, const uint8_t* in, int in_c, int check, const uint8_t* arg, uint8_t* out, int out_c) {
    switch (group) {
        case 0: return mnt_op<Mnt4G1>(op, in, in_c, check, arg, out, out_c);
        case 1: return mnt_op<Mnt4G2>(op, in, in_c, check, arg, out, out_c);
        case 2: return mnt_op<Mnt6G1>(op, in, in_c, check, arg, out, out_c);
        case 3: return mnt_op<Mnt6G2>(op, in, in_c, check, arg, out, out_c);
    }
    return -100;
}
}
