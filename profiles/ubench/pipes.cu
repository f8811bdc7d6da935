// Issue-rate micro-benchmark for the instructions the DP kernels are built from (sm_100a).
// Every test runs 8 independent dependency chains per thread, 8 warps per SM sub-partition, and reports
// warp-instructions per cycle per SM sub-partition (1.0 = the scheduler's issue limit).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t d; asm volatile("prmt.b32 %0,%1,%2,%3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
    if (OP == 0) return __vadd2(a, b);
    if (OP == 1) return __vimax3_s16x2(a, b, c);
    if (OP == 2) return __viaddmax_s16x2(a, b, c);
    if (OP == 3) return prmt(a, b, 0xFDB9);
    if (OP == 4) return (a & b) ^ c;                       // LOP3
    if (OP == 5) return h2u(__hadd2(u2h(a), u2h(b)));      // HADD2
    if (OP == 6) return h2u(__hmax2(u2h(a), u2h(b)));      // HMNMX2
    if (OP == 7) return h2u(__hfma2_relu(u2h(a), u2h(b), u2h(c)));
    if (OP == 8) return a * b + c;                         // IMAD
    if (OP == 9) return a + b + c;                         // IADD3
    if (OP == 10) return __vmaxs2(a, b);                   // VIMNMX 16x2 two-input
    if (OP == 11) return __vsub2(a, b);
    if (OP == 12) return h2u(__hfma2(u2h(a), u2h(b), u2h(c)));
    if (OP == 13) return __vminu2(a, b);
    return a;
}

// MIX: alternate OPA and OPB on independent chains (4 chains each)
template <int OPA, int OPB>
__global__ void k(uint32_t *out, uint32_t s0, uint32_t s1, long long *cyc)
{
    uint32_t v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = s0 * (threadIdx.x + i + 1);
    uint32_t b = s1 | 1, c = s0 ^ 0x3c003c00u;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) v[i] = (i & 1) ? op<OPB>(v[i], b, c) : op<OPA>(v[i], b, c);
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OPA, int OPB> void run(const char *name, uint32_t *out, long long *cyc, int sm)
{
    const int threads = 1024;        // 32 warps per SM = 8 per sub-partition
    k<OPA, OPB><<<sm, threads>>>(out, 3, 5, cyc);
    k<OPA, OPB><<<sm, threads>>>(out, 3, 5, cyc);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sm * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sm; ++i) avg += (double)h[i]; avg /= sm;
    const double winst_per_smsp = (double)ITERS * CHAINS * 8;     // 8 warps per sub-partition
    printf("%-34s %.3f warp-inst/clk/SMSP\n", name, winst_per_smsp / avg);
}

int main()
{
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, (size_t)sm * 1024 * 4); cudaMalloc(&cyc, 256 * 8);
    printf("SMs %d\n", sm);
    run<0, 0>("VIADD.16x2", out, cyc, sm);
    run<11, 11>("VSUB 16x2", out, cyc, sm);
    run<1, 1>("VIMNMX3.S16x2", out, cyc, sm);
    run<10, 10>("VIMNMX.S16x2 (2-input)", out, cyc, sm);
    run<13, 13>("VIMNMX.U16x2 (min)", out, cyc, sm);
    run<2, 2>("VIADDMNMX.S16x2", out, cyc, sm);
    run<3, 3>("PRMT", out, cyc, sm);
    run<4, 4>("LOP3", out, cyc, sm);
    run<9, 9>("IADD3", out, cyc, sm);
    run<8, 8>("IMAD", out, cyc, sm);
    run<5, 5>("HADD2", out, cyc, sm);
    run<12, 12>("HFMA2", out, cyc, sm);
    run<6, 6>("HMNMX2", out, cyc, sm);
    run<7, 7>("HFMA2.RELU", out, cyc, sm);
    run<0, 5>("VIADD.16x2 + HADD2", out, cyc, sm);
    run<1, 5>("VIMNMX3.S16x2 + HADD2", out, cyc, sm);
    run<2, 12>("VIADDMNMX.S16x2 + HFMA2", out, cyc, sm);
    run<3, 5>("PRMT + HADD2", out, cyc, sm);
    run<4, 8>("LOP3 + IMAD", out, cyc, sm);
    run<0, 8>("VIADD.16x2 + IMAD", out, cyc, sm);
    run<0, 1>("VIADD.16x2 + VIMNMX3", out, cyc, sm);
    run<6, 5>("HMNMX2 + HADD2", out, cyc, sm);
    run<6, 0>("HMNMX2 + VIADD.16x2", out, cyc, sm);
    return 0;
}
