// k_misc.cuh — small plumbing kernels: contig unpack + reverse complement, exclusive scans,
// per-read minimizer self-occurrence counts (input of minimap2's mm_seed_mz_flt, seed.c).
#pragma once
#include <cuda_runtime.h>
#include "mm_types.cuh"

namespace telr {

// nt4 byte copies of every contig: forward strand at ctg_off[l], reverse complement right after it.
// (reference: get_rev_comp_sequence, TELR_utility.py:67-73, called at TELR_te.py:624-627)
__global__ void k_unpack_contigs(const uint32_t *seq2, const uint32_t *nmask, int n_loci, const int64_t *contig_off,
                                 const int32_t *contig_len, const int64_t *ctg_boff, uint8_t *ctg_bytes)
{
    for (int l = blockIdx.x; l < n_loci; l += gridDim.x) {
        const int L = contig_len[l];
        const int64_t so = contig_off[l], bo = ctg_boff[l];
        for (int i = threadIdx.x; i < L; i += blockDim.x) {
            int64_t p = so + i;
            int c = (seq2[p >> 4] >> (2 * (p & 15))) & 3;
            if ((nmask[p >> 5] >> (p & 31)) & 1) c = 4;
            ctg_bytes[bo + i] = (uint8_t)c;
            ctg_bytes[bo + L + (L - 1 - i)] = (uint8_t)(c < 4 ? 3 - c : 4);
        }
    }
}

// out[i] = sum_{j<i} in[j], out[n] = total.  Single CTA of 1024 threads.
template <class TI> __global__ void __launch_bounds__(1024) k_excl_scan(const TI *in, int64_t *out, int n, int64_t *max_out)
{
    __shared__ int64_t ws[33];
    __shared__ int64_t carry_s, mx_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0, mx_s = 0;
    __syncthreads();
    int64_t mymax = 0;
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int64_t v = i < n ? (int64_t)in[i] : 0, x = v;
        mymax = v > mymax ? v : mymax;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) ws[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int64_t s = ws[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int64_t y = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= d) s += y;
            }
            ws[lane] = s;
        }
        __syncthreads();
        int64_t pre = carry_s + (wid ? ws[wid - 1] : 0) + x - v;
        if (i < n) out[i] = pre;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = pre + v;
        __syncthreads();
    }
    // block max
#pragma unroll
    for (int d = 16; d; d >>= 1) { int64_t y = __shfl_xor_sync(0xffffffffu, mymax, d); mymax = y > mymax ? y : mymax; }
    if (lane == 0) ws[wid] = mymax;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t m = 0;
        for (int i = 0; i < 32; ++i) m = ws[i] > m ? ws[i] : m;
        out[n] = carry_s;
        if (max_out) *max_out = m;
    }
}

// For every read minimizer, how many minimizers of the same read carry the same (hash,span) word.
// One CTA per read; open-addressing table in a per-CTA slice of global scratch (L2 resident).
__global__ void __launch_bounds__(256) k_self_count(int n_reads, const int64_t *mz_off, const uint64_t *mz_x, uint16_t *selfcnt,
                                                    uint64_t *tab_keys, uint32_t *tab_cnt, int64_t tab_stride)
{
    uint64_t *keys = tab_keys + (int64_t)blockIdx.x * tab_stride;
    uint32_t *cnts = tab_cnt + (int64_t)blockIdx.x * tab_stride;
    for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const int64_t b = mz_off[r];
        const int n = (int)(mz_off[r + 1] - b);
        if (n <= 10) {          // no filter can apply below min_mid_occ (>= 10): counts are irrelevant
            for (int i = threadIdx.x; i < n; i += blockDim.x) selfcnt[b + i] = 1;
            continue;
        }
        int sz = 64;
        while (sz < 2 * n) sz <<= 1;
        const uint32_t m = (uint32_t)sz - 1;
        for (int i = threadIdx.x; i < sz; i += blockDim.x) keys[i] = ~0ULL, cnts[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t x = mz_x[b + i];
            uint32_t s = (uint32_t)(mix64(x) >> 20) & m;
            for (;;) {
                unsigned long long old = atomicCAS((unsigned long long *)&keys[s], ~0ULL, (unsigned long long)x);
                if (old == ~0ULL || old == x) { atomicAdd(&cnts[s], 1u); break; }
                s = (s + 1) & m;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t x = mz_x[b + i];
            uint32_t s = (uint32_t)(mix64(x) >> 20) & m;
            while (keys[s] != x) s = (s + 1) & m;
            uint32_t c = cnts[s];
            selfcnt[b + i] = (uint16_t)(c > 65535u ? 65535u : c);
        }
        __syncthreads();
    }
}

}  // namespace telr
