// Second issue-rate micro-benchmark (round 2): which pipe do the 32-bit adds of the offset-form recurrence (k_fill.cuh) use?
// 8 chains per thread (loop unrolled 8x: 64 operations per 3 loop-control instructions), each op takes its second variable operand from the neighbouring chain so that nothing folds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu && ./pipes2      (check the loop bodies with cuobjdump -sass)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t d; asm volatile("prmt.b32 %0,%1,%2,%3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("lop3.b32 %0,%1,%2,%3,0xf8;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t iadd3n(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("{.reg .b32 t; sub.s32 t,%1,%2; add.s32 %0,t,%3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t isub(uint32_t a, uint32_t b) { uint32_t d; asm volatile("sub.s32 %0,%1,%2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t imadsub(uint32_t a, uint32_t b, uint32_t one) { uint32_t d; asm volatile("mad.lo.s32 %0,%1,%2,%3;" : "=r"(d) : "r"(b), "r"(one), "r"(a)); return d; }

template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t n, uint32_t b, uint32_t c)
{
    if (OP == 0) return __vadd2(a, n);                      // VIADD.16x2
    if (OP == 1) return __vimax3_s16x2(a, n, c);
    if (OP == 2) return __viaddmax_u16x2(a, n, c);
    if (OP == 3) return prmt(a, n, 0xFDB9);
    if (OP == 4) return lop3(a, n, c);                      // LOP3 (a | (n & c))
    if (OP == 5) return iadd3n(a, n, c);                    // IADD3 a - n + c
    if (OP == 6) return isub(a, n);                         // 2-input subtract: IADD3 or IMAD.IADD, the compiler's choice
    if (OP == 7) return imadsub(a, n, b);                   // IMAD with a run-time multiplier
    return a;
}

template <int OPA, int OPB, int NA>           // NA of the 8 chains run OPA, the others OPB
__global__ void k(uint32_t *out, uint32_t s0, uint32_t s1, long long *cyc)
{
    uint32_t v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = s0 * (threadIdx.x + i + 1);
    uint32_t b = s1, c = s0 ^ 0x3c003c00u;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) v[i] = i < NA ? op<OPA>(v[i], v[(i + 3) & 7], b, c) : op<OPB>(v[i], v[(i + 3) & 7], b, c);
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OPA, int OPB, int NA> void run(const char *name, uint32_t *out, long long *cyc, int sm)
{
    const int threads = 1024;        // 32 warps per SM = 8 per sub-partition
    k<OPA, OPB, NA><<<sm, threads>>>(out, 3, 1, cyc);
    k<OPA, OPB, NA><<<sm, threads>>>(out, 3, 1, cyc);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sm * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sm; ++i) avg += (double)h[i]; avg /= sm;
    printf("%-44s %.3f warp-inst/clk/SMSP\n", name, (double)ITERS * CHAINS * 8 / avg);
}

int main()
{
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, (size_t)sm * 1024 * 4); cudaMalloc(&cyc, 256 * 8);
    run<0, 0, 8>("VIADD.16x2", out, cyc, sm);
    run<3, 3, 8>("PRMT", out, cyc, sm);
    run<1, 1, 8>("VIMNMX3.S16x2", out, cyc, sm);
    run<2, 2, 8>("VIADDMNMX.U16x2", out, cyc, sm);
    run<4, 4, 8>("LOP3 (pinned)", out, cyc, sm);
    run<5, 5, 8>("IADD3 a-b+c (3 variable inputs)", out, cyc, sm);
    run<6, 6, 8>("2-input subtract", out, cyc, sm);
    run<7, 7, 8>("IMAD (run-time multiplier)", out, cyc, sm);
    run<5, 0, 4>("IADD3 + VIADD.16x2  4:4", out, cyc, sm);
    run<5, 3, 4>("IADD3 + PRMT  4:4", out, cyc, sm);
    run<5, 1, 4>("IADD3 + VIMNMX3  4:4", out, cyc, sm);
    run<7, 3, 4>("IMAD + PRMT  4:4", out, cyc, sm);
    run<7, 5, 4>("IMAD + IADD3  4:4", out, cyc, sm);
    run<4, 0, 4>("LOP3 + VIADD.16x2  4:4", out, cyc, sm);
    run<4, 3, 4>("LOP3 + PRMT  4:4", out, cyc, sm);
    run<4, 0, 2>("LOP3 + VIADD.16x2  2:6", out, cyc, sm);
    run<0, 3, 3>("VIADD.16x2 + PRMT  3:5", out, cyc, sm);
    run<0, 3, 2>("VIADD.16x2 + PRMT  2:6", out, cyc, sm);
    run<2, 0, 4>("VIADDMNMX.U16x2 + VIADD.16x2  4:4", out, cyc, sm);
    return 0;
}
