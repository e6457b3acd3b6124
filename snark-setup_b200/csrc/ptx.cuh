// Carry-chain primitives for multi-limb integer arithmetic on sm_100a.
//
// On the device every primitive is ONE PTX instruction that reads/writes the CC.CF carry flag
// (add.cc / addc.cc / mad.lo.cc / madc.hi.cc ...).  ptxas fuses an adjacent
// mad.lo.cc + madc.hi.cc pair that targets an aligned register pair into a single
// IMAD.WIDE.U32(.X) — that is why fp.cuh lays its accumulators out as even/odd limb pairs.
//
// When this header is compiled by a host compiler (tests/emul only) the same primitives are
// emulated with 64-bit arithmetic and an explicit thread-local carry, so the limb algorithms in
// fp.cuh can be unit-tested on a box without a GPU.  The emulation is never part of the product
// library: api.cu refuses to build without __CUDACC__.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SS_HD __host__ __device__ __forceinline__
#define SS_D __device__ __forceinline__
#else
#define SS_HD inline
#define SS_D inline
#endif

namespace ss {

#if !defined(__CUDA_ARCH__)
// host emulation of the PTX carry flag
inline uint32_t& cc_flag() {
    static thread_local uint32_t cf = 0;
    return cf;
}
#endif

SS_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }

SS_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

#if defined(__CUDA_ARCH__)
#define SS_ASM2(name, ptx)                                                     \
    SS_D uint32_t name(uint32_t a, uint32_t b) {                               \
        uint32_t r;                                                            \
        asm volatile(ptx " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));           \
        return r;                                                              \
    }
#define SS_ASM3(name, ptx)                                                     \
    SS_D uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                   \
        uint32_t r;                                                            \
        asm volatile(ptx " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); \
        return r;                                                              \
    }
SS_ASM2(add_cc, "add.cc.u32")
SS_ASM2(addc_cc, "addc.cc.u32")
SS_ASM2(addc, "addc.u32")
SS_ASM2(sub_cc, "sub.cc.u32")
SS_ASM2(subc_cc, "subc.cc.u32")
SS_ASM2(subc, "subc.u32")
SS_ASM3(mad_lo_cc, "mad.lo.cc.u32")
SS_ASM3(madc_lo_cc, "madc.lo.cc.u32")
SS_ASM3(madc_lo, "madc.lo.u32")
SS_ASM3(mad_hi_cc, "mad.hi.cc.u32")
SS_ASM3(madc_hi_cc, "madc.hi.cc.u32")
SS_ASM3(madc_hi, "madc.hi.u32")
#undef SS_ASM2
#undef SS_ASM3
#else
inline uint32_t add_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a + b;
    cc_flag() = (uint32_t)(t >> 32);
    return (uint32_t)t;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a + b + cc_flag();
    cc_flag() = (uint32_t)(t >> 32);
    return (uint32_t)t;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cc_flag(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a - b;
    cc_flag() = (uint32_t)(t >> 63);  // borrow
    return (uint32_t)t;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a - b - cc_flag();
    cc_flag() = (uint32_t)(t >> 63);
    return (uint32_t)t;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cc_flag(); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return addc(a * b, c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
#endif

}  // namespace ss
