// Microbenchmarks behind the integer-pipe roofline: measures, on the box,
//   * raw IMAD / IMAD.WIDE issue rate (the "MAC32 peak" the roofline is quoted against)
//   * Montgomery multiplications/s of fp_mul for 8/12/24 limbs at several occupancies
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DSS_MUL_INLINE] -o imad_bench.bin imad_bench.cu
#include <cstdio>
#include <vector>
#include "../snark-setup_b200/csrc/ec.cuh"
using namespace ss;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_imad(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b = seed | 1, c = threadIdx.x;
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = a[i] * b + c;
    }
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void k_imad_wide(uint64_t* out, int iters, uint32_t seed) {
    uint64_t a[8];
    uint32_t b = seed | 1;
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t lo = (uint32_t)a[i];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(lo), "r"(b));
        }
    }
    uint64_t r = 0;
    for (int i = 0; i < 8; i++) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// carry-chained wide MACs: the exact instruction the multiplier is made of
__global__ void k_imad_wide_x(uint32_t* out, int iters, uint32_t seed) {
    uint32_t acc[16], a[8], b = seed | 1;
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < iters; it++) {
        acc[0] = mad_lo_cc(a[0], b, acc[0]);
        acc[1] = madc_hi_cc(a[0], b, acc[1]);
#pragma unroll
        for (int j = 1; j < 8; j++) {
            acc[2 * j] = madc_lo_cc(a[j], b, acc[2 * j]);
            acc[2 * j + 1] = madc_hi_cc(a[j], b, acc[2 * j + 1]);
        }
        b += acc[15];
    }
    uint32_t r = 0;
    for (int i = 0; i < 16; i++) r ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void k_imad_hi(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b = seed | 1, c = threadIdx.x;
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 0x9e3779b1u + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
    }
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// BLS12-377 Fq with a compile-time INV = 0xffffffff: ptxas then emits the m*p products as separate
// IMAD.X + IMAD.HI.U32.X pairs instead of IMAD.WIDE.U32.X (411 vs 305 IMAD-pipe instructions)
struct Bls377FqSplit : Bls377Fq {
    SS_HD static uint32_t inv() { return 0xffffffffu; }
};

__global__ void k_dfma(double* out, int iters, double seed) {
    double a[8], b = seed, c = 1.0 / 3.0;
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
    }
    double r = 0;
    for (int i = 0; i < 8; i++) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// IMAD and DFMA issued from the same warp: do the two pipes overlap?
__global__ void k_imad_dfma(double* out, int iters, double seed, uint32_t iseed) {
    double a[4], b = seed, c = 1.0 / 3.0;
    uint32_t x[8], y = iseed | 1, z = threadIdx.x;
    for (int i = 0; i < 4; i++) a[i] = threadIdx.x + i;
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = fma(a[i], b, c);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = x[i] * y + z;
    }
    double r = 0;
    for (int i = 0; i < 4; i++) r += a[i];
    uint32_t q = 0;
    for (int i = 0; i < 8; i++) q ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + q;
}

template <class P>
__global__ void k_fpmul(uint32_t* out, int iters, uint32_t seed) {
    Fp<P> x, y;
    for (int i = 0; i < P::N; i++) { x.l[i] = seed + threadIdx.x * 31 + i; y.l[i] = seed * 3 + blockIdx.x + i; }
    x.l[P::N - 1] = 0; y.l[P::N - 1] = 0;
    for (int it = 0; it < iters; it++) { Fp<P> r = fp_mul(x, y); y = x; x = r; }
    uint32_t r = 0;
    for (int i = 0; i < P::N; i++) r ^= x.l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class P>
__global__ void k_fpsqr(uint32_t* out, int iters, uint32_t seed) {
    Fp<P> x;
    for (int i = 0; i < P::N; i++) x.l[i] = seed + threadIdx.x * 31 + i;
    x.l[P::N - 1] = 0;
    for (int it = 0; it < iters; it++) x = fp_sqr(x);
    uint32_t r = 0;
    for (int i = 0; i < P::N; i++) r ^= x.l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// two independent multiplications per iteration (ILP 2)
template <class P>
__global__ void k_fpmul2(uint32_t* out, int iters, uint32_t seed) {
    Fp<P> x, y, u, v;
    for (int i = 0; i < P::N; i++) { x.l[i] = seed + threadIdx.x * 31 + i; y.l[i] = seed * 3 + blockIdx.x + i; u.l[i] = x.l[i] ^ 5; v.l[i] = y.l[i] ^ 9; }
    x.l[P::N - 1] = 0; y.l[P::N - 1] = 0; u.l[P::N - 1] = 0; v.l[P::N - 1] = 0;
    for (int it = 0; it < iters; it++) { Fp<P> r = fp_mul_inl(x, y); Fp<P> s = fp_mul_inl(u, v); y = x; x = r; v = u; u = s; }
    uint32_t r = 0;
    for (int i = 0; i < P::N; i++) r ^= x.l[i] ^ u.l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class K>
static double time_kernel(K launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();  // warm
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    uint32_t* out; CK(cudaMalloc(&out, (size_t)sms * 16 * 1024 * 8));
    const int iters = 4096;
    for (int tpb : {128, 256, 512, 1024}) {
        int blocks = sms * (2048 / tpb);
        double t = time_kernel([&] { k_imad<<<blocks, tpb>>>(out, iters, 12345); }, 5);
        printf("{\"bench\": \"imad\", \"threads_per_sm\": 2048, \"tpb\": %d, \"gops\": %.1f}\n", tpb, (double)blocks * tpb * iters * 8 / t * 1e-9);
        t = time_kernel([&] { k_imad_hi<<<blocks, tpb>>>(out, iters, 12345); }, 5);
        printf("{\"bench\": \"imad_hi\", \"threads_per_sm\": 2048, \"tpb\": %d, \"gops\": %.1f}\n", tpb, (double)blocks * tpb * iters * 8 / t * 1e-9);
        t = time_kernel([&] { k_imad_wide<<<blocks, tpb>>>((uint64_t*)out, iters, 12345); }, 5);
        printf("{\"bench\": \"imad_wide\", \"threads_per_sm\": 2048, \"tpb\": %d, \"gops\": %.1f}\n", tpb, (double)blocks * tpb * iters * 8 / t * 1e-9);
        t = time_kernel([&] { k_imad_wide_x<<<blocks, tpb>>>(out, iters, 12345); }, 5);
        printf("{\"bench\": \"imad_wide_x_chain\", \"threads_per_sm\": 2048, \"tpb\": %d, \"gops\": %.1f}\n", tpb, (double)blocks * tpb * iters * 8 / t * 1e-9);
    }
    {
        int tpb = 256, blocks = sms * 8;
        double t = time_kernel([&] { k_dfma<<<blocks, tpb>>>((double*)out, iters, 1.0000001); }, 5);
        printf("{\"bench\": \"dfma\", \"gops\": %.1f}\n", (double)blocks * tpb * iters * 8 / t * 1e-9);
        t = time_kernel([&] { k_imad_dfma<<<blocks, tpb>>>((double*)out, iters, 1.0000001, 777); }, 5);
        printf("{\"bench\": \"imad8+dfma4 per iter\", \"imad_gops\": %.1f, \"dfma_gops\": %.1f}\n",
               (double)blocks * tpb * iters * 8 / t * 1e-9, (double)blocks * tpb * iters * 4 / t * 1e-9);
    }
    CK(cudaGetLastError());
#if defined(SS_MUL_INLINE)
    const char* mode = "inline";
#else
    const char* mode = "call";
#endif
    for (int warps_per_sm : {4, 8, 16, 32, 48, 64}) {
        int tpb = 128, blocks = sms * warps_per_sm * 32 / tpb;
        const int it = 2000;
        double t;
        t = time_kernel([&] { k_fpmul<Bls377Fr><<<blocks, tpb>>>(out, it, 7); }, 3);
        printf("{\"bench\": \"fp_mul\", \"mode\": \"%s\", \"limbs\": 8, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", mode, warps_per_sm, (double)blocks * tpb * it / t * 1e-9);
        t = time_kernel([&] { k_fpmul<Bls377Fq><<<blocks, tpb>>>(out, it, 7); }, 3);
        printf("{\"bench\": \"fp_mul\", \"mode\": \"%s\", \"limbs\": 12, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", mode, warps_per_sm, (double)blocks * tpb * it / t * 1e-9);
        t = time_kernel([&] { k_fpsqr<Bls377Fq><<<blocks, tpb>>>(out, it, 7); }, 3);
        printf("{\"bench\": \"fp_sqr\", \"mode\": \"%s\", \"limbs\": 12, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", mode, warps_per_sm, (double)blocks * tpb * it / t * 1e-9);
        t = time_kernel([&] { k_fpmul<Bls377FqSplit><<<blocks, tpb>>>(out, it, 7); }, 3);
        printf("{\"bench\": \"fp_mul_split_mp\", \"limbs\": 12, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", warps_per_sm, (double)blocks * tpb * it / t * 1e-9);
        t = time_kernel([&] { k_fpmul2<Bls377Fq><<<blocks, tpb>>>(out, it, 7); }, 3);
        printf("{\"bench\": \"fp_mul_ilp2_inline\", \"limbs\": 12, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", warps_per_sm, (double)blocks * tpb * it * 2 / t * 1e-9);
        if (warps_per_sm <= 32) {
            t = time_kernel([&] { k_fpmul<Bw6Fq><<<blocks, tpb>>>(out, it / 4, 7); }, 3);
            printf("{\"bench\": \"fp_mul\", \"mode\": \"%s\", \"limbs\": 24, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", mode, warps_per_sm, (double)blocks * tpb * (it / 4) / t * 1e-9);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
