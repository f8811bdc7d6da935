// k_depth.cuh — kernels (e)+(f): per-base depth of every contig strand from alignment M-blocks, window
// medians and the AF reduction, one CTA per locus, depth held in shared memory.
//
// Replaces `samtools depth -aa -r chr:S-E` + statistics.median (get_median_cov, TELR_te.py:870-884), the
// window selection of get_te_cov (TELR_te.py:841-867) and get_flank_cov (TELR_te.py:518-550), the ratio rule
// get_te_flank_ratio (TELR_te.py:564-575) and the AF block (TELR_te.py:810-835).  Region semantics: the
// reference pastes 0-based BED coordinates into a 1-based inclusive samtools region, so "c:S-E" covers
// D[max(S,1)-1 .. min(E,L)-1].  Medians are carried as 2*median integers; rounding/clamping of AF and the
// text formatting stay in Python (telr_b200/stage4.py).
//
// Two accumulation modes (A/B under ncu): 0 = shared-memory atomicAdd per covered base (north_star (e)),
// 1 = +1/-1 marks at block boundaries followed by a block-wide prefix sum.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "mm_types.cuh"

namespace telr {

constexpr int DP_THREADS = 256;

struct DepthArgs {
    int32_t n_loci, mode;
    const int32_t *contig_len, *te_start, *te_end, *locus_read_begin;
    const int64_t *prob_blk_off; const int32_t *prob_blk_cnt; const int2 *blocks;
    int32_t flank_len, flank_off, te_len, te_off;
    int32_t *depth; const int64_t *depth_off;     // optional output: fw then rc per locus
    int32_t *cov2x; double *af;
    int32_t max_len;
    // contigs whose depth row does not fit the shared-memory row (smem_ints entries) use a per-CTA row in global memory
    int32_t smem_ints; int32_t *grow; int64_t grow_stride;
    const uint8_t *locus_bad;     // optional: 1 = locus outside the implemented scope (reported as -4 / NaN, never fails the batch)
};

__device__ int dp_block_sum(int v, int *ws)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if (lane == 0) ws[wid] = v;
    __syncthreads();
    int s = 0;
    for (int i = 0; i < DP_THREADS / 32; ++i) s += ws[i];
    return s;
}

// 2 * median of sd[beg, end) ; -3 when empty.  All threads of the CTA must call.
__device__ int median2x(const int32_t *sd, int L, int S, int E, int *ws)
{
    long long beg = (long long)S - 1, end = E;
    if (beg < 0) beg = 0;
    if (end > L) end = L;
    if (beg >= end) return -3;
    const int n = (int)(end - beg);
    const int32_t *a = sd + beg;
    const int k0 = (n - 1) >> 1, k1 = n >> 1;
    __shared__ int r0, r1;
    if (n <= 1024) {
        for (int i = threadIdx.x; i < n; i += DP_THREADS) {
            int v = a[i], rank = 0;
            for (int j = 0; j < n; ++j) { int wv = a[j]; rank += (wv < v) || (wv == v && j < i); }
            if (rank == k0) r0 = v;
            if (rank == k1) r1 = v;
        }
        __syncthreads();
    } else {
        for (int which = 0; which < 2; ++which) {
            const int k = which ? k1 : k0;
            int lo = 0, hi = 1 << 24;           // smallest v with #(a <= v) >= k+1
            while (lo < hi) {
                int mid = (lo + hi) >> 1, c = 0;
                for (int i = threadIdx.x; i < n; i += DP_THREADS) c += a[i] <= mid;
                c = dp_block_sum(c, ws);
                if (c >= k + 1) hi = mid; else lo = mid + 1;
            }
            if (threadIdx.x == 0) { if (which) r1 = lo; else r0 = lo; }
        }
        __syncthreads();
    }
    int r = r0 + r1;
    __syncthreads();
    return r;
}

__device__ __forceinline__ bool af_ratio(int te2, int fl2, double *r)
{
    if (te2 <= 0 || fl2 <= 0) return false;
    *r = ((double)te2 / 2.0) / ((double)fl2 / 2.0);
    return !(*r > 1.5);
}

__global__ void __launch_bounds__(DP_THREADS) k_depth_af(const __grid_constant__ DepthArgs A)
{
    extern __shared__ __align__(16) int32_t sd_smem[];      // [min(max_len + 2, smem_ints)]
    __shared__ int ws[40];
    __shared__ int cov[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int l = blockIdx.x; l < A.n_loci; l += gridDim.x) {
        const int L = A.contig_len[l];
        const int ts = A.te_start[l], te = A.te_end[l];
        const int rb = A.locus_read_begin[l], nr = A.locus_read_begin[l + 1] - rb;
        if (L <= 0 || (A.locus_bad && A.locus_bad[l])) {
            if (tid < 8) A.cov2x[(int64_t)l * 8 + tid] = L <= 0 ? -2 : -4;
            if (tid == 0) A.af[l] = nan("");
            continue;
        }
        int32_t *sd = (L + 2 <= A.smem_ints) ? sd_smem : A.grow + (int64_t)blockIdx.x * A.grow_stride;
        for (int strand = 0; strand < 2; ++strand) {
            for (int i = tid; i <= L; i += DP_THREADS) sd[i] = 0;
            __syncthreads();
            const int p0 = 2 * rb + strand * nr;
            if (A.mode == 0) {
                for (int pr = wid; pr < nr; pr += DP_THREADS / 32) {
                    const int2 *blk = A.blocks + A.prob_blk_off[p0 + pr];
                    const int nb = A.prob_blk_cnt[p0 + pr];
                    for (int b = 0; b < nb; ++b) {
                        const int2 bk = blk[b];
                        for (int i = lane; i < bk.y; i += 32) {
                            int pos = bk.x + i;
                            if (pos >= 0 && pos < L) atomicAdd(&sd[pos], 1);
                        }
                    }
                }
                __syncthreads();
            } else {
                for (int pr = wid; pr < nr; pr += DP_THREADS / 32) {
                    const int2 *blk = A.blocks + A.prob_blk_off[p0 + pr];
                    const int nb = A.prob_blk_cnt[p0 + pr];
                    for (int b = lane; b < nb; b += 32) {
                        const int2 bk = blk[b];
                        int s = bk.x < 0 ? 0 : bk.x, e = bk.x + bk.y > L ? L : bk.x + bk.y;
                        if (e > s) { atomicAdd(&sd[s], 1); atomicAdd(&sd[e], -1); }
                    }
                }
                __syncthreads();
                // block-wide inclusive scan over sd[0..L)
                int carry = 0;
                for (int base = 0; base < L; base += DP_THREADS) {
                    int i = base + tid, v = i < L ? sd[i] : 0, x = v;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { int yv = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += yv; }
                    if (lane == 31) ws[wid] = x;
                    __syncthreads();
                    int pre = 0;
                    for (int k = 0; k < wid; ++k) pre += ws[k];
                    int tot = 0;
                    for (int k = 0; k < DP_THREADS / 32; ++k) tot += ws[k];
                    if (i < L) sd[i] = carry + pre + x;
                    carry += tot;
                    __syncthreads();
                }
            }
            if (A.depth) {
                int32_t *out = A.depth + A.depth_off[l] + (int64_t)strand * L;
                for (int i = tid; i < L; i += DP_THREADS) out[i] = sd[i];
            }
            // window medians (get_te_cov / get_flank_cov)
            int c4[4] = {-1, -1, -1, -1};
            if (ts >= 0) {
                const int s = strand ? L - te : ts, e = strand ? L - ts : te;
                const int fl = A.flank_len, fo = A.flank_off, ti = A.te_len, to = A.te_off;
                bool whole = true;
                if (ti && s + to + ti < e) {
                    whole = false;
                    c4[0] = median2x(sd, L, s + to, s + to + ti, ws);
                    c4[1] = median2x(sd, L, e - ti - to, e - to, ws);
                }
                if (whole) { c4[0] = median2x(sd, L, s, e, ws); c4[1] = c4[0]; }
                if (s - fl - fo >= 0) c4[2] = median2x(sd, L, s - fl - fo, s - fo, ws);
                if (e + fl + fo <= L) c4[3] = median2x(sd, L, e + fo, e + fl + fo, ws);
            } else c4[0] = c4[1] = c4[2] = c4[3] = -2;
            if (tid == 0) for (int k = 0; k < 4; ++k) cov[strand * 4 + k] = c4[k];
            __syncthreads();
        }
        if (tid == 0) {
            double t5 = 0, t3 = 0, f;
            bool h5 = af_ratio(cov[0], cov[2], &t5), h3 = af_ratio(cov[4], cov[6], &t3);
            if (h5 && h3) f = fabs(t5 - t3) <= 0.3 ? (t5 + t3) / 2 : nan("");
            else if (h5) f = t5;
            else if (h3) f = t3;
            else f = nan("");
            if (ts < 0) f = nan("");
            A.af[l] = f;
            for (int k = 0; k < 8; ++k) A.cov2x[(int64_t)l * 8 + k] = cov[k];
        }
        __syncthreads();
    }
}

}  // namespace telr
