// CPU oracle for the snark-setup batch-exponentiation hot path.
//
// TEST INFRASTRUCTURE ONLY — not part of the product.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product path
// (snark-setup_b200/) never does and has no CPU fallback.
//
// PARITY UNPINNED against the real reference binary: nimiq/snark-setup holds no golden vectors for
// this path, its arithmetic lives in un-vendored arkworks 0.4 (Cargo.lock:80-82,103-105,151-153,
// 187-189,459-461) and no Rust toolchain exists in the build container.  This file restates the
// published arkworks algorithms at the reference's call sites and is pinned (a) against the
// independent big-int restatement oracle/pyref.py and (b) against tests/golden/*.json produced from
// pyref by tests/golden/make_golden.py.
//
// What it follows (the *reference algorithm*, which is also what the CPU baseline times):
//   generate_powers_of_tau   setup-utils/src/helpers.rs:32-37   (tau.pow([i]) per element)
//   batch_exp                setup-utils/src/helpers.rs:75-140  (MSB-first double-and-add per
//                            element = ark-ec mul_bigint, then normalize_batch)
//   read_batch/write_batch   setup-utils/src/io/read.rs:110-135, io/write.rs:57-66
//   check_subgroup           setup-utils/src/elements.rs:123-150 (p.mul_bigint(r).is_zero())
//   merge_pairs/power_pairs  setup-utils/src/helpers.rs:371-390 (msm_bigint: Pippenger)
//   apply_powers             phase1/src/helpers/buffers.rs:77-97
//   Phase1::computation      phase1/src/computation.rs:40-193 (Groth16)
//   to_coeffs                setup-utils/src/groth16_utils.rs:44-53 (group IFFT, double-and-add twiddles)
// Arithmetic: Montgomery CIOS on 64-bit limbs with unsigned __int128 (ark-ff MontBackend uses the
// same limb size), Jacobian a=0 formulas dbl-2009-l / madd-2007-bl as ark-ec 0.4 does.
#include <omp.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "constants_gen.h"
#include "mont_asm_gen.h"  // mulx / adcx / adox Montgomery product for 4- and 6-limb moduli (tools/gen_mont_asm.py)

typedef unsigned __int128 u128;

#if defined(ORACLE_MONT_ASM)
static const bool kHaveAdx = __builtin_cpu_supports("adx") && __builtin_cpu_supports("bmi2");
#endif

namespace {

enum { E_OK = 0, E_INVALID_DATA = 1, E_UNEXPECTED_FLAGS = 2, E_POINT_AT_INFINITY = 3, E_INCORRECT_SUBGROUP = 4, E_INVALID_LENGTH = 5 };
enum { CHK_FULL = 0, CHK_NONZERO = 1, CHK_INGROUP = 2, CHK_NO = 3 };

// ------------------------------------------------------------------------------------------------
static bool g_force_portable = false;
static inline bool force_portable() { return g_force_portable; }

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint64_t v[N];

    static Fp zero() { Fp r; memset(r.v, 0, sizeof r.v); return r; }
    static Fp one() { Fp r; memcpy(r.v, P::ONE, sizeof r.v); return r; }
    bool is_zero() const { uint64_t t = 0; for (int i = 0; i < N; i++) t |= v[i]; return t == 0; }
    bool operator==(const Fp& o) const { return memcmp(v, o.v, sizeof v) == 0; }
    bool operator!=(const Fp& o) const { return !(*this == o); }

    static bool geq_mod(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) {
            if (a[i] > P::MOD[i]) return true;
            if (a[i] < P::MOD[i]) return false;
        }
        return true;
    }
    static void sub_mod(uint64_t* a) {
        u128 br = 0;
        for (int i = 0; i < N; i++) {
            u128 t = (u128)a[i] - P::MOD[i] - br;
            a[i] = (uint64_t)t;
            br = (t >> 64) & 1;
        }
    }
    Fp operator+(const Fp& o) const {
        Fp r; u128 c = 0;
        for (int i = 0; i < N; i++) { c += (u128)v[i] + o.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
        if (geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r; u128 br = 0;
        for (int i = 0; i < N; i++) { u128 t = (u128)v[i] - o.v[i] - br; r.v[i] = (uint64_t)t; br = (t >> 64) & 1; }
        if (br) { u128 c = 0; for (int i = 0; i < N; i++) { c += (u128)r.v[i] + P::MOD[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
        return r;
    }
    Fp neg() const { return is_zero() ? *this : zero() - *this; }
    Fp dbl() const { return *this + *this; }
    // CIOS Montgomery product, "no-carry" variant (valid because every modulus here leaves the top
    // bit of the top limb clear): the accumulator never exceeds N limbs.
    Fp operator*(const Fp& o) const {
#if defined(ORACLE_MONT_ASM)
        // same CIOS product through the two ADX carry chains — what LLVM makes of ark-ff's MontBackend (and what its
        // `asm` feature hand-writes); the portable loop below is the fallback and the cross-check (tests)
        if ((N == 4 || N == 6) && kHaveAdx && !force_portable()) {
            Fp r;
            if (N == 6) mont_mul_asm_6(r.v, v, o.v, P::MOD, P::INV);
            else mont_mul_asm_4(r.v, v, o.v, P::MOD, P::INV);
            if (geq_mod(r.v)) sub_mod(r.v);
            return r;
        }
#endif
        uint64_t t[N];
#pragma GCC unroll 16
        for (int j = 0; j < N; j++) t[j] = 0;
#pragma GCC unroll 16
        for (int i = 0; i < N; i++) {
            const uint64_t bi = o.v[i];
            u128 A = (u128)v[0] * bi + t[0];
            const uint64_t m = (uint64_t)A * P::INV;
            u128 C = (u128)m * P::MOD[0] + (uint64_t)A;
            A >>= 64;
            C >>= 64;
#pragma GCC unroll 16
            for (int j = 1; j < N; j++) {
                A += (u128)v[j] * bi + t[j];
                C += (u128)m * P::MOD[j] + (uint64_t)A;
                t[j - 1] = (uint64_t)C;
                A >>= 64;
                C >>= 64;
            }
            t[N - 1] = (uint64_t)(C + A);
        }
        Fp r;
#pragma GCC unroll 16
        for (int j = 0; j < N; j++) r.v[j] = t[j];
        if (geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    Fp sqr() const { return *this * *this; }
    static constexpr int TWO_ADICITY_ = P::TWO_ADICITY;
    static constexpr int BITS_ = P::BITS;
    static Fp fft_root() { Fp r; memcpy(r.v, P::FFT_ROOT, sizeof r.v); return r; }  // ark-ff TWO_ADIC_ROOT_OF_UNITY
    static Fp from_raw(const uint64_t* raw) { Fp a, r2; memcpy(a.v, raw, sizeof a.v); memcpy(r2.v, P::R2, sizeof r2.v); return a * r2; }
    void to_raw(uint64_t* out) const { Fp o = zero(); o.v[0] = 1; Fp c = *this * o; memcpy(out, c.v, sizeof c.v); }
    Fp pow(const uint64_t* e, int n) const {
        Fp r = one();
        for (int i = n - 1; i >= 0; i--)
            for (int b = 63; b >= 0; b--) { r = r.sqr(); if ((e[i] >> b) & 1) r = r * *this; }
        return r;
    }
    Fp inv() const { return pow(P::PM2, N); }
    bool sqrt(Fp& out) const {
        if (is_zero()) { out = *this; return true; }
        if (pow(P::PM1H, N) != one()) return false;
        if (P::P3MOD4) { out = pow(P::PP1Q, N); return true; }
        Fp w = pow(P::TM1H, N), x = *this * w, b = x * w, z; memcpy(z.v, P::ZT, sizeof z.v);
        int vv = P::TWO_ADICITY;
        while (b != one()) {
            int k = 0; Fp t = b;
            while (t != one()) { t = t.sqr(); k++; }
            Fp wj = z;
            for (int j = 0; j < vv - k - 1; j++) wj = wj.sqr();
            z = wj.sqr(); b = b * z; x = x * wj; vv = k;
        }
        out = x; return true;
    }
    // canonical compare helper: is this (Montgomery) element "negative", i.e. y > -y ?
    bool lex_largest() const {
        uint64_t a[N], b[N]; to_raw(a); neg().to_raw(b);
        for (int i = N - 1; i >= 0; i--) { if (a[i] > b[i]) return true; if (a[i] < b[i]) return false; }
        return false;
    }
    static constexpr int BYTES = N * 8;
    // returns error code
    static int read(const uint8_t* p, bool flags_present, Fp& out, uint8_t& flags) {
        uint64_t raw[N]; memcpy(raw, p, BYTES);
        flags = 0;
        if (flags_present) {
            flags = (uint8_t)((raw[N - 1] >> 56) & 0xC0);
            raw[N - 1] &= ~((uint64_t)0xC0 << 56);
            if (flags == 0xC0) return E_UNEXPECTED_FLAGS;
        }
        if (geq_mod(raw)) return E_INVALID_DATA;
        out = from_raw(raw); return E_OK;
    }
    void write(uint8_t* p, uint8_t flags) const { uint64_t raw[N]; to_raw(raw); memcpy(p, raw, BYTES); p[BYTES - 1] |= flags; }
};

template <class P>
struct Fp2 {  // Fp[u]/(u^2+5)
    typedef Fp<P> B;
    B c0, c1;
    static Fp2 zero() { return Fp2{B::zero(), B::zero()}; }
    static Fp2 one() { return Fp2{B::one(), B::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const Fp2& o) const { return !(*this == o); }
    Fp2 operator+(const Fp2& o) const { return Fp2{c0 + o.c0, c1 + o.c1}; }
    Fp2 operator-(const Fp2& o) const { return Fp2{c0 - o.c0, c1 - o.c1}; }
    Fp2 neg() const { return Fp2{c0.neg(), c1.neg()}; }
    Fp2 dbl() const { return Fp2{c0.dbl(), c1.dbl()}; }
    static B mul5(const B& x) { B t = x.dbl().dbl(); return t + x; }
    Fp2 operator*(const Fp2& o) const {
        B v0 = c0 * o.c0, v1 = c1 * o.c1;
        return Fp2{v0 - mul5(v1), (c0 + c1) * (o.c0 + o.c1) - v0 - v1};
    }
    Fp2 sqr() const { return *this * *this; }
    Fp2 inv() const { B n = c0.sqr() + mul5(c1.sqr()); B ni = n.inv(); return Fp2{c0 * ni, (c1 * ni).neg()}; }
    bool sqrt(Fp2& out) const {
        if (c1.is_zero()) {
            B r;
            if (c0.sqrt(r)) { out = Fp2{r, B::zero()}; return true; }
            B q = (c0 * mul5(B::one()).inv()).neg();
            if (!q.sqrt(r)) return false;
            out = Fp2{B::zero(), r}; return true;
        }
        B norm = c0.sqr() + mul5(c1.sqr()), alpha;
        if (!norm.sqrt(alpha)) return false;
        B half; memcpy(half.v, P::HALF, sizeof half.v);
        B delta = (c0 + alpha) * half, r0;
        if (!delta.sqrt(r0)) { delta = (c0 - alpha) * half; if (!delta.sqrt(r0)) return false; }
        out = Fp2{r0, c1 * r0.dbl().inv()}; return true;
    }
    bool lex_largest() const { return c1.is_zero() ? c0.lex_largest() : c1.lex_largest(); }
    static constexpr int BYTES = 2 * B::BYTES;
    static int read(const uint8_t* p, bool flags_present, Fp2& out, uint8_t& flags) {
        uint8_t f0; flags = 0;
        int e = B::read(p, false, out.c0, f0);
        if (e) return e;
        return B::read(p + B::BYTES, flags_present, out.c1, flags);
    }
    void write(uint8_t* p, uint8_t flags) const { c0.write(p, 0); c1.write(p + B::BYTES, flags); }
};

// ------------------------------------------------------------------------------------------------
template <class F> struct Aff { F x, y; bool inf; };
template <class F> struct Jac {
    F X, Y, Z;
    static Jac identity() { return Jac{F::one(), F::one(), F::zero()}; }
    bool is_identity() const { return Z.is_zero(); }
};

template <class F> Jac<F> jdbl(const Jac<F>& p) {
    if (p.Z.is_zero()) return p;
    F A = p.X.sqr(), B = p.Y.sqr(), C = B.sqr();
    F D = ((p.X + B).sqr() - A - C).dbl();
    F E = A.dbl() + A, G = E.sqr();
    Jac<F> r;
    r.X = G - D.dbl();
    r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
    r.Z = (p.Y * p.Z).dbl();
    return r;
}
template <class F> Jac<F> jmadd(const Jac<F>& p, const Aff<F>& q) {
    if (q.inf) return p;
    if (p.Z.is_zero()) return Jac<F>{q.x, q.y, F::one()};
    F Z1Z1 = p.Z.sqr(), U2 = q.x * Z1Z1, S2 = q.y * p.Z * Z1Z1;
    F H = U2 - p.X, rr = S2 - p.Y;
    if (H.is_zero()) return rr.is_zero() ? jdbl(p) : Jac<F>::identity();
    rr = rr.dbl();
    F HH = H.sqr(), I = HH.dbl().dbl(), J = H * I, V = p.X * I;
    Jac<F> r;
    r.X = rr.sqr() - J - V.dbl();
    r.Y = rr * (V - r.X) - (p.Y * J).dbl();
    r.Z = (p.Z + H).sqr() - Z1Z1 - HH;
    return r;
}
template <class F> Jac<F> jadd(const Jac<F>& p, const Jac<F>& q) {
    if (p.Z.is_zero()) return q;
    if (q.Z.is_zero()) return p;
    F Z1Z1 = p.Z.sqr(), Z2Z2 = q.Z.sqr(), U1 = p.X * Z2Z2, U2 = q.X * Z1Z1;
    F S1 = p.Y * q.Z * Z2Z2, S2 = q.Y * p.Z * Z1Z1, H = U2 - U1, rr = S2 - S1;
    if (H.is_zero()) return rr.is_zero() ? jdbl(p) : Jac<F>::identity();
    rr = rr.dbl();
    F I = H.dbl().sqr(), J = H * I, V = U1 * I;
    Jac<F> r;
    r.X = rr.sqr() - J - V.dbl();
    r.Y = rr * (V - r.X) - (S1 * J).dbl();
    r.Z = ((p.Z + q.Z).sqr() - Z1Z1 - Z2Z2) * H;
    return r;
}
// ark-ec mul_bigint: MSB-first double-and-add over the bits of a little-endian u64 array
template <class F> Jac<F> jmul(const Aff<F>& b, const uint64_t* k, int nlimbs) {
    Jac<F> acc = Jac<F>::identity();
    if (b.inf) return acc;
    bool started = false;
    for (int i = nlimbs - 1; i >= 0; i--)
        for (int j = 63; j >= 0; j--) {
            bool bit = (k[i] >> j) & 1;
            if (started) acc = jdbl(acc);
            if (bit) { acc = jmadd(acc, b); started = true; }
        }
    return acc;
}

// k * P for a Jacobian base (MSB-first double-and-add, the generic `Group *= scalar` of ark-ec)
template <class F> Jac<F> jmul_jac(const Jac<F>& b, const uint64_t* k, int nlimbs) {
    Jac<F> acc = Jac<F>::identity();
    for (int i = nlimbs - 1; i >= 0; i--)
        for (int bit = 63; bit >= 0; bit--) {
            acc = jdbl(acc);
            if ((k[i] >> bit) & 1) acc = jadd(acc, b);
        }
    return acc;
}

// ------------------------------------------------------------------------------------------------
struct BlsG1 {
    typedef Fp<OBls377Fq> F; typedef Fp<OBls377Fr> Fr;
    static constexpr int USIZE = 96, CSIZE = 48, FRL = 4;
    static F b() { F r; memcpy(r.v, kO_bls_g1_b, sizeof r.v); return r; }
    static const uint64_t* order() { return kO_bls_r; }
};
struct BlsG2 {
    typedef Fp2<OBls377Fq> F; typedef Fp<OBls377Fr> Fr;
    static constexpr int USIZE = 192, CSIZE = 96, FRL = 4;
    static F b() { F r; memcpy(r.c0.v, kO_bls_g2_b0, sizeof r.c0.v); memcpy(r.c1.v, kO_bls_g2_b1, sizeof r.c1.v); return r; }
    static const uint64_t* order() { return kO_bls_r; }
};
struct BwG1 {
    typedef Fp<OBw6Fq> F; typedef Fp<OBls377Fq> Fr;
    static constexpr int USIZE = 192, CSIZE = 96, FRL = 6;
    static F b() { F r; memcpy(r.v, kO_bw6_g1_b, sizeof r.v); return r; }
    static const uint64_t* order() { return kO_bw6_r; }
};
struct BwG2 {
    typedef Fp<OBw6Fq> F; typedef Fp<OBls377Fq> Fr;
    static constexpr int USIZE = 192, CSIZE = 96, FRL = 6;
    static F b() { F r; memcpy(r.v, kO_bw6_g2_b, sizeof r.v); return r; }
    static const uint64_t* order() { return kO_bw6_r; }
};

template <class G> bool in_subgroup(const Aff<typename G::F>& p) { return jmul(p, G::order(), G::FRL).is_identity(); }
template <class G> bool on_curve(const Aff<typename G::F>& p) { return p.inf || p.y.sqr() == p.x.sqr() * p.x + G::b(); }

// Deserializer::read_element (setup-utils/src/io/read.rs:57-73)
template <class G> int decode(const uint8_t* p, bool compressed, int check, Aff<typename G::F>& out) {
    typedef typename G::F F;
    uint8_t fl; int e;
    out.inf = false;
    if (compressed) {
        if ((e = F::read(p, true, out.x, fl))) return e;
        if (fl & 0x40) out.inf = true;
        else {
            F y;
            if (!(out.x.sqr() * out.x + G::b()).sqrt(y)) return E_INVALID_DATA;
            out.y = (y.lex_largest() == ((fl & 0x80) != 0)) ? y : y.neg();
        }
    } else {
        uint8_t f0;
        if ((e = F::read(p, false, out.x, f0))) return e;
        if ((e = F::read(p + F::BYTES, true, out.y, fl))) return e;
        if (fl & 0x40) out.inf = true;
    }
    if (out.inf) { out.x = F::zero(); out.y = F::zero(); }
    else if (check == CHK_FULL || check == CHK_INGROUP) {
        if (!on_curve<G>(out) || !in_subgroup<G>(out)) return E_INVALID_DATA;
    }
    if ((check == CHK_FULL || check == CHK_NONZERO) && out.inf) return E_POINT_AT_INFINITY;
    return E_OK;
}
// Serializer::write_element (setup-utils/src/io/write.rs:30-35)
template <class G> void encode(uint8_t* p, bool compressed, const Aff<typename G::F>& a) {
    typedef typename G::F F;
    int sz = compressed ? G::CSIZE : G::USIZE;
    if (a.inf) { memset(p, 0, sz); p[sz - 1] = 0x40; return; }
    uint8_t fl = a.y.lex_largest() ? 0x80 : 0;
    if (compressed) a.x.write(p, fl);
    else { a.x.write(p, 0); a.y.write(p + F::BYTES, fl); }
}

// CurveGroup::normalize_batch: Montgomery's trick, identities skipped
template <class F> void normalize_batch(const std::vector<Jac<F>>& in, std::vector<Aff<F>>& out) {
    size_t n = in.size();
    out.resize(n);
    std::vector<F> pre(n);
    F acc = F::one();
    for (size_t i = 0; i < n; i++) if (!in[i].Z.is_zero()) { pre[i] = acc; acc = acc * in[i].Z; }
    F inv = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (in[i].Z.is_zero()) { out[i].inf = true; out[i].x = F::zero(); out[i].y = F::zero(); continue; }
        F zi = inv * pre[i]; inv = inv * in[i].Z;
        F zi2 = zi.sqr();
        out[i].x = in[i].X * zi2; out[i].y = in[i].Y * zi2 * zi; out[i].inf = false;
    }
}

// tau.pow([e])  (ark-ff Field::pow: MSB-first square-and-multiply over the u64)
template <class FrT> FrT fr_pow_u64(const FrT& tau, uint64_t e) {
    FrT r = FrT::one();
    bool started = false;
    for (int b = 63; b >= 0; b--) {
        if (started) r = r.sqr();
        if ((e >> b) & 1) { r = r * tau; started = true; }
    }
    return r;
}

// apply_powers over n elements (phase1/src/helpers/buffers.rs:77-97) with
// exps[i] = explicit or tau^(first+i); coeff optional.  Returns error code; *err_index = lowest bad index.
template <class G>
int apply_powers(const uint8_t* in, int in_c, int check, uint8_t* out, int out_c, size_t n, const uint8_t* powers,
                 const uint8_t* tau_le, uint64_t first, const uint8_t* coeff_le, uint64_t* err_index) {
    typedef typename G::F F; typedef typename G::Fr Fr;
    const int isz = in_c ? G::CSIZE : G::USIZE, osz = out_c ? G::CSIZE : G::USIZE, fb = G::FRL * 8;
    std::vector<Jac<F>> proj(n);
    Fr tau = Fr::zero(), coeff = Fr::one();
    uint64_t raw[G::FRL];
    if (tau_le) { memcpy(raw, tau_le, fb); tau = Fr::from_raw(raw); }
    if (coeff_le) { memcpy(raw, coeff_le, fb); coeff = Fr::from_raw(raw); }
    int err = 0; uint64_t bad = ~0ull;
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < n; i++) {
        Aff<F> p;
        int e = decode<G>(in + i * isz, in_c != 0, check, p);
        if (e) {
#pragma omp critical
            if (i < bad) { bad = i; err = e; }
            proj[i] = Jac<F>::identity();
            continue;
        }
        Fr s;
        if (powers) { uint64_t r2[G::FRL]; memcpy(r2, powers + i * fb, fb); s = Fr::from_raw(r2); }
        else s = fr_pow_u64(tau, first + i);
        if (coeff_le) s = s * coeff;
        uint64_t k[G::FRL]; s.to_raw(k);
        proj[i] = jmul(p, k, G::FRL);
    }
    if (err) { if (err_index) *err_index = bad; return err; }
    // normalize_batch in per-thread chunks (ark-ec parallel normalize_batch does the same)
    std::vector<Aff<F>> aff(n);
    const size_t chunk = 1024;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t c = 0; c < (n + chunk - 1) / chunk; c++) {
        size_t lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
        std::vector<Jac<F>> part(proj.begin() + lo, proj.begin() + hi);
        std::vector<Aff<F>> o;
        normalize_batch(part, o);
        for (size_t i = lo; i < hi; i++) encode<G>(out + i * osz, out_c != 0, o[i - lo]);
    }
    return E_OK;
}

// to_coeffs (setup-utils/src/groth16_utils.rs:44-53): domain.ifft over the group + normalize_batch, restated as an
// iterative decimation-in-frequency transform (Gentleman-Sande: butterflies (a, b) -> (a + b, w^j (a - b)), output in
// bit-reversed order) followed by the 1/n scaling — every twiddle and the scaling a full double-and-add scalar
// multiplication, as ark-poly's generic FFT over C::Group does (n/2 log n + n of them).
template <class G>
int group_ifft(const uint8_t* in, int in_c, int check, uint8_t* out, int out_c, size_t n, uint64_t* err_index) {
    typedef typename G::F F; typedef typename G::Fr Fr;
    const int isz = in_c ? G::CSIZE : G::USIZE, osz = out_c ? G::CSIZE : G::USIZE;
    int log_n = 0;
    while (((size_t)1 << log_n) < n) log_n++;
    if (((size_t)1 << log_n) != n || log_n > Fr::TWO_ADICITY_) return E_INVALID_LENGTH;
    std::vector<Jac<F>> v(n);
    for (size_t i = 0; i < n; i++) {
        Aff<F> p;
        int e = decode<G>(in + i * isz, in_c != 0, check, p);
        if (e) { if (err_index) *err_index = i; return e; }
        v[i] = p.inf ? Jac<F>::identity() : Jac<F>{p.x, p.y, F::one()};
    }
    // F::get_root_of_unity: TWO_ADIC_ROOT_OF_UNITY squared down to order n; the inverse transform uses w^-1
    Fr w = Fr::fft_root();
    for (int i = log_n; i < Fr::TWO_ADICITY_; i++) w = w.sqr();
    Fr winv = w.inv();
    for (size_t half = n >> 1; half >= 1; half >>= 1) {
        // twiddles of this stage: wm^j, wm = winv^(n / (2 half))
        Fr wm = winv;
        for (size_t s = n / (2 * half); s > 1; s >>= 1) wm = wm.sqr();
        std::vector<Fr> tw(half);
        tw[0] = Fr::one();
        for (size_t j = 1; j < half; j++) tw[j] = tw[j - 1] * wm;
#pragma omp parallel for schedule(dynamic, 16)
        for (size_t t = 0; t < n / 2; t++) {
            const size_t blk = t / half, j = t % half, lo = blk * 2 * half + j, hi = lo + half;
            Jac<F> a = v[lo], b = v[hi];
            v[lo] = jadd(a, b);
            Jac<F> nb = b; nb.Y = nb.Y.neg();
            Jac<F> d = jadd(a, nb);
            uint64_t k[G::FRL]; tw[j].to_raw(k);
            v[hi] = jmul_jac(d, k, G::FRL);
        }
    }
    Fr nn = Fr::zero();
    { uint64_t raw[G::FRL] = {0}; raw[0] = (uint64_t)n; nn = Fr::from_raw(raw); }
    uint64_t kinv[G::FRL]; nn.inv().to_raw(kinv);
    std::vector<Jac<F>> res(n);
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < n; i++) {
        size_t r = 0;
        for (int b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
        res[r] = jmul_jac(v[i], kinv, G::FRL);
    }
    std::vector<Aff<F>> aff;
    normalize_batch(res, aff);
    for (size_t i = 0; i < n; i++) encode<G>(out + i * osz, out_c != 0, aff[i]);
    return E_OK;
}

template <class G>
int transcode(const uint8_t* in, int in_c, int check, uint8_t* out, int out_c, size_t n, int rmul, uint64_t* err_index) {
    typedef typename G::F F;
    const int isz = in_c ? G::CSIZE : G::USIZE, osz = out_c ? G::CSIZE : G::USIZE;
    int err = 0; uint64_t bad = ~0ull;
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < n; i++) {
        Aff<F> p;
        int e = decode<G>(in + i * isz, in_c != 0, check, p);
        if (!e && rmul && !in_subgroup<G>(p)) e = E_INCORRECT_SUBGROUP;
        if (e) {
#pragma omp critical
            if (i < bad) { bad = i; err = e; }
            continue;
        }
        if (out) encode<G>(out + i * osz, out_c != 0, p);
    }
    if (err && err_index) *err_index = bad;
    return err;
}

// naive MSM (sum of k_i * P_i) used to pin merge_pairs at small sizes
template <class G>
int msm_naive(const uint8_t* pts, int compressed, size_t n, const uint8_t* scalars, uint8_t* out_uncompressed) {
    typedef typename G::F F;
    const int isz = compressed ? G::CSIZE : G::USIZE, fb = G::FRL * 8;
    Jac<F> acc = Jac<F>::identity();
    for (size_t i = 0; i < n; i++) {
        Aff<F> p;
        int e = decode<G>(pts + i * isz, compressed != 0, CHK_NO, p);
        if (e) return e;
        uint64_t k[G::FRL]; memcpy(k, scalars + i * fb, fb);
        acc = jadd(acc, jmul(p, k, G::FRL));
    }
    std::vector<Jac<F>> one(1, acc); std::vector<Aff<F>> a;
    normalize_batch(one, a);
    encode<G>(out_uncompressed, false, a[0]);
    return E_OK;
}

// ---- merge_pairs / power_pairs with the reference's MSM ---------------------------------------------------------
// VariableBaseMSM::msm_bigint as ark-ec 0.4.2 runs it for short-Weierstrass points (negation is free, so the signed
// "wNAF" bucket method): c = ln(n) + 2 (3 below 32 points), radix-2^c signed digits, one task per window, buckets
// 1 .. 2^(c-1), running-sum bucket reduction, windows combined high to low with c doublings each.
static inline size_t ln_without_floats(size_t a) {
    size_t lg = 0;
    while ((a >> (lg + 1)) != 0) lg++;
    return lg * 69 / 100;
}

template <class G>
Jac<typename G::F> msm_pippenger(const Aff<typename G::F>* pts, const uint64_t* scalars /*[n][FRL] canonical*/, size_t n, int num_bits) {
    typedef typename G::F F;
    const size_t c = n < 32 ? 3 : ln_without_floats(n) + 2;
    const size_t digits_count = (num_bits + c - 1) / c;
    const int64_t radix = (int64_t)1 << c, window_mask = radix - 1;
    // make_digits: signed radix-2^c digits of every scalar
    std::vector<int64_t> digits(n * digits_count);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        const uint64_t* k = scalars + i * G::FRL;
        int64_t carry = 0;
        for (size_t d = 0; d < digits_count; d++) {
            const size_t bit = d * c, u = bit / 64, sh = bit % 64;
            uint64_t bits = u < (size_t)G::FRL ? k[u] >> sh : 0;
            if (sh + c > 64 && u + 1 < (size_t)G::FRL) bits |= k[u + 1] << (64 - sh);
            int64_t coef = carry + (int64_t)(bits & (uint64_t)window_mask);
            carry = (coef + radix / 2) >> c;
            digits[i * digits_count + d] = coef - (carry << c);
        }
        digits[i * digits_count + digits_count - 1] += carry << c;
    }
    std::vector<Jac<F>> window_sums(digits_count);
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t w = 0; w < digits_count; w++) {
        // digits lie in [-2^(c-1), 2^(c-1)) — except the LAST one, which takes the final carry back and reaches 2^c when
        // the top window is full (253 = 11 * 23: every vector of 2^14 .. 2^15 - 1 pairs, scalars >= 2^252); its window
        // gets 2^c buckets so that buckets[d - 1] stays in range
        const size_t nbuckets = (size_t)1 << (w + 1 == digits_count ? c : c - 1);
        std::vector<Jac<F>> buckets(nbuckets, Jac<F>::identity());
        for (size_t i = 0; i < n; i++) {
            const int64_t d = digits[i * digits_count + w];
            if (d > 0) buckets[d - 1] = jmadd(buckets[d - 1], pts[i]);
            else if (d < 0) { Aff<F> m = pts[i]; m.y = m.y.neg(); buckets[-d - 1] = jmadd(buckets[-d - 1], m); }
        }
        Jac<F> run = Jac<F>::identity(), res = Jac<F>::identity();
        for (size_t b = buckets.size(); b-- > 0;) { run = jadd(run, buckets[b]); res = jadd(res, run); }
        window_sums[w] = res;
    }
    Jac<F> total = Jac<F>::identity();
    for (size_t w = digits_count; w-- > 1;) {
        total = jadd(total, window_sums[w]);
        for (size_t k = 0; k < c; k++) total = jdbl(total);
    }
    return jadd(total, window_sums[0]);
}

static inline uint64_t splitmix64(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// One vector of the verification loop, as the reference runs it (phase1/src/verification.rs:243-411 with the whole
// vector as one window): check_elements_are_nonzero_and_in_prime_order_subgroup (accumulator.rs:95-145: read_batch
// with OnlyNonZero, then p.mul_bigint(r).is_zero() per element), check_power_ratios (accumulator.rs:56-91: a SECOND
// read_batch of the same bytes, power_pairs = two msm_bigint over full-width random scalars, helpers.rs:371-390) and the
// write_batch of the uncompressed elements (verification.rs:271-274).  The 2 pairings per vector are not included.
template <class G>
int verify_vector(const uint8_t* in, int in_c, uint8_t* out, int out_c, size_t n, int subgroup, int ratio, uint64_t seed,
                  int decode_passes, uint8_t* s_out, uint8_t* sx_out, uint64_t* err_index) {
    typedef typename G::F F; typedef typename G::Fr Fr;
    const int isz = in_c ? G::CSIZE : G::USIZE, osz = out_c ? G::CSIZE : G::USIZE;
    std::vector<Aff<F>> pts(n);
    int err = 0; uint64_t bad = ~0ull;
    for (int pass = 0; pass < (decode_passes < 1 ? 1 : decode_passes); pass++) {
#pragma omp parallel for schedule(dynamic, 16)
        for (size_t i = 0; i < n; i++) {
            int e = decode<G>(in + i * isz, in_c != 0, CHK_NONZERO, pts[i]);
            if (!e && pass == 0 && subgroup && !in_subgroup<G>(pts[i])) e = E_INCORRECT_SUBGROUP;
            if (e) {
#pragma omp critical
                if (i < bad) { bad = i; err = e; }
            }
        }
        if (err) { if (err_index) *err_index = bad; return err; }
    }
    if (ratio && n >= 2) {
        const size_t m = n - 1;
        std::vector<uint64_t> rho(m * G::FRL);
        uint64_t st = seed;
        const int top_bits = Fr::BITS_ - 64 * (G::FRL - 1);
        for (size_t i = 0; i < m; i++) {  // Fr::rand: limbs masked to the modulus length, rejected when >= r
            uint64_t* k = &rho[i * G::FRL];
            do {
                for (int j = 0; j < G::FRL; j++) k[j] = splitmix64(st);
                if (top_bits < 64) k[G::FRL - 1] &= ((uint64_t)1 << top_bits) - 1;
            } while (Fr::geq_mod(k));
        }
        Jac<F> res[2];
        res[0] = msm_pippenger<G>(pts.data(), rho.data(), m, Fr::BITS_);
        res[1] = msm_pippenger<G>(pts.data() + 1, rho.data(), m, Fr::BITS_);
        std::vector<Jac<F>> two(res, res + 2); std::vector<Aff<F>> a;
        normalize_batch(two, a);
        encode<G>(s_out, false, a[0]);
        encode<G>(sx_out, false, a[1]);
    }
    if (out) {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) encode<G>(out + i * osz, out_c != 0, pts[i]);
    }
    return E_OK;
}

#define DISPATCH(curve, group, CALL)                 \
    do {                                             \
        if (curve == 0 && group == 0) return CALL(BlsG1); \
        if (curve == 0 && group == 1) return CALL(BlsG2); \
        if (curve == 1 && group == 0) return CALL(BwG1);  \
        if (curve == 1 && group == 1) return CALL(BwG2);  \
        return -1;                                   \
    } while (0)

}  // namespace

extern "C" {

int oracle_threads(void) { return omp_get_max_threads(); }
void oracle_set_threads(int n) { omp_set_num_threads(n); }

int oracle_apply_powers(int curve, int group, const uint8_t* in, int in_c, int check, uint8_t* out, int out_c, size_t n,
                        const uint8_t* powers, const uint8_t* tau, uint64_t first, const uint8_t* coeff, uint64_t* err_index) {
#define CALL(G) apply_powers<G>(in, in_c, check, out, out_c, n, powers, tau, first, coeff, err_index)
    DISPATCH(curve, group, CALL);
#undef CALL
}

int oracle_transcode(int curve, int group, const uint8_t* in, int in_c, int check, uint8_t* out, int out_c, size_t n,
                     int rmul_subgroup, uint64_t* err_index) {
#define CALL(G) transcode<G>(in, in_c, check, out, out_c, n, rmul_subgroup, err_index)
    DISPATCH(curve, group, CALL);
#undef CALL
}

int oracle_group_ifft(int curve, int group, const uint8_t* in, int in_c, int check, uint8_t* out, int out_c, size_t n,
                      uint64_t* err_index) {
    if (curve == 0) return group == 0 ? group_ifft<BlsG1>(in, in_c, check, out, out_c, n, err_index)
                                      : group_ifft<BlsG2>(in, in_c, check, out, out_c, n, err_index);
    return group == 0 ? group_ifft<BwG1>(in, in_c, check, out, out_c, n, err_index)
                      : group_ifft<BwG2>(in, in_c, check, out, out_c, n, err_index);
}

int oracle_msm(int curve, int group, const uint8_t* pts, int compressed, size_t n, const uint8_t* scalars, uint8_t* out) {
#define CALL(G) msm_naive<G>(pts, compressed, n, scalars, out)
    DISPATCH(curve, group, CALL);
#undef CALL
}

int oracle_powers(int curve, const uint8_t* tau, uint64_t start, uint64_t end, uint8_t* out) {
    if (curve == 0) {
        typedef Fp<OBls377Fr> Fr; uint64_t raw[4]; memcpy(raw, tau, 32); Fr t = Fr::from_raw(raw);
#pragma omp parallel for
        for (uint64_t i = start; i < end; i++) { uint64_t k[4]; fr_pow_u64(t, i).to_raw(k); memcpy(out + (i - start) * 32, k, 32); }
        return 0;
    }
    if (curve == 1) {
        typedef Fp<OBls377Fq> Fr; uint64_t raw[6]; memcpy(raw, tau, 48); Fr t = Fr::from_raw(raw);
#pragma omp parallel for
        for (uint64_t i = start; i < end; i++) { uint64_t k[6]; fr_pow_u64(t, i).to_raw(k); memcpy(out + (i - start) * 48, k, 48); }
        return 0;
    }
    return -1;
}

// Phase1::computation, Groth16, FULL or CHUNKED (phase1/src/computation.rs:40-193).  The caller passes the
// vector counts (from Phase1Parameters) so this file does not duplicate the size arithmetic under test.
int oracle_phase1_computation(int curve, const uint8_t* in, uint8_t* out, int cin, int cout, int check, uint64_t n_g1,
                              uint64_t n_other, uint64_t first_power, const uint8_t* tau, const uint8_t* alpha,
                              const uint8_t* beta) {
    const int g1u = curve == 0 ? 96 : 192, g1c = curve == 0 ? 48 : 96, g2u = 192, g2c = 96;
    const size_t s1i = cin ? g1c : g1u, s2i = cin ? g2c : g2u, s1o = cout ? g1c : g1u, s2o = cout ? g2c : g2u;
    size_t oi = 64, oo = 64;
    uint64_t bad;
    int e;
    if ((e = oracle_apply_powers(curve, 0, in + oi, cin, check, out + oo, cout, n_g1, nullptr, tau, first_power, nullptr, &bad))) return e;
    oi += n_g1 * s1i; oo += n_g1 * s1o;
    if ((e = oracle_apply_powers(curve, 1, in + oi, cin, check, out + oo, cout, n_other, nullptr, tau, first_power, nullptr, &bad))) return e;
    oi += n_other * s2i; oo += n_other * s2o;
    if ((e = oracle_apply_powers(curve, 0, in + oi, cin, check, out + oo, cout, n_other, nullptr, tau, first_power, alpha, &bad))) return e;
    oi += n_other * s1i; oo += n_other * s1o;
    if ((e = oracle_apply_powers(curve, 0, in + oi, cin, check, out + oo, cout, n_other, nullptr, tau, first_power, beta, &bad))) return e;
    oi += n_other * s1i; oo += n_other * s1o;
    // beta_g2 <- beta * beta_g2: tau^0 * beta
    uint8_t one[48] = {1};
    return oracle_apply_powers(curve, 1, in + oi, cin, check, out + oo, cout, 1, nullptr, one, 0, beta, &bad);
}

// merge of the two: explicit-scalar Pippenger MSM (pins msm_pippenger against the naive sum in tests)
int oracle_msm_pippenger(int curve, int group, const uint8_t* pts, int compressed, size_t n, const uint8_t* scalars, uint8_t* out) {
#define CALL(G) ([&]() -> int {                                                                      \
        typedef typename G::F F;                                                                     \
        const int isz = compressed ? G::CSIZE : G::USIZE;                                            \
        std::vector<Aff<F>> p(n);                                                                    \
        for (size_t i = 0; i < n; i++) { int e = decode<G>(pts + i * isz, compressed != 0, CHK_NO, p[i]); if (e) return e; } \
        std::vector<uint64_t> k(n * G::FRL);                                                         \
        memcpy(k.data(), scalars, n * G::FRL * 8);                                                   \
        std::vector<Jac<F>> one(1, msm_pippenger<G>(p.data(), k.data(), n, G::Fr::BITS_));          \
        std::vector<Aff<F>> a;                                                                       \
        normalize_batch(one, a);                                                                     \
        encode<G>(out, false, a[0]);                                                                 \
        return 0; })()
    DISPATCH(curve, group, CALL);
#undef CALL
}

int oracle_verify_vector(int curve, int group, const uint8_t* in, int in_c, uint8_t* out, int out_c, size_t n, int subgroup,
                         int ratio, uint64_t seed, int decode_passes, uint8_t* s_out, uint8_t* sx_out, uint64_t* err_index) {
#define CALL(G) verify_vector<G>(in, in_c, out, out_c, n, subgroup, ratio, seed, decode_passes, s_out, sx_out, err_index)
    DISPATCH(curve, group, CALL);
#undef CALL
}

// The per-vector loop of Phase1::verification over a whole Groth16 response (phase1/src/verification.rs:217-411):
// tau_g1, tau_g2, alpha_g1, beta_g1 through verify_vector, beta_g2 read with Full and re-emitted (:199-201).
// `pairs`: 4 x (s || sx) uncompressed like ss_phase1_verification_vectors.
int oracle_phase1_verification_vectors(int curve, const uint8_t* in, int cin, uint8_t* out, int cout, uint64_t n_g1,
                                       uint64_t n_other, uint64_t seed, int decode_passes, uint8_t* pairs, uint64_t* err_index) {
    const int g1u = curve == 0 ? 96 : 192, g1c = curve == 0 ? 48 : 96, g2u = 192, g2c = 96;
    const size_t s1i = cin ? g1c : g1u, s2i = cin ? g2c : g2u, s1o = cout ? g1c : g1u, s2o = cout ? g2c : g2u;
    const uint64_t cnt[4] = {n_g1, n_other, n_other, n_other};
    const int grp[4] = {0, 1, 0, 0};
    size_t oi = 64, oo = 64, po = 0;
    int e;
    for (int v = 0; v < 4; v++) {
        const size_t usz = grp[v] ? g2u : g1u;
        if (cnt[v]) {
            e = oracle_verify_vector(curve, grp[v], in + oi, cin, out ? out + oo : nullptr, cout, cnt[v], 1, cnt[v] >= 2, seed + v,
                                     decode_passes, pairs + po, pairs + po + usz, err_index);
            if (e) return e;
        }
        oi += cnt[v] * (grp[v] ? s2i : s1i);
        oo += cnt[v] * (grp[v] ? s2o : s1o);
        po += 2 * usz;
    }
    return oracle_transcode(curve, 1, in + oi, cin, CHK_FULL, out ? out + oo : nullptr, cout, 1, 0, err_index);
}

void oracle_force_portable_mul(int on) { g_force_portable = on != 0; }
int oracle_has_asm_mul(void) {
#if defined(ORACLE_MONT_ASM)
    return kHaveAdx ? 1 : 0;
#else
    return 0;
#endif
}

// ns per Montgomery multiplication of the BLS12-377 base field on one core (dependent chain), printed beside `cores`
double oracle_fq_mul_ns(int iters) {
    typedef Fp<OBls377Fq> F;
    F a = F::one() + F::one(), b;
    memcpy(b.v, kO_bls_g1_b, sizeof b.v);
    b = b + b + b;
    a = a * b + b;
    const double t0 = omp_get_wtime();
    for (int i = 0; i < iters; i++) a = a * b;
    const double t1 = omp_get_wtime();
    volatile uint64_t sink = a.v[0];
    (void)sink;
    return (t1 - t0) * 1e9 / iters;
}

}  // extern "C"
