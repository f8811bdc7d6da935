// k_chain.cuh — kernels (b)+(c): per-locus contig minimizer index in shared memory, seed collection,
// anchor sort, chaining DP with warp-shuffle max-scan, RMQ re-chaining, region generation.
//
// Replaces, per (read x contig strand), minimap2's mm_idx_gen/mm_idx_get (index.c), mm_seed_mz_flt /
// mm_collect_matches (seed.c), collect_seed_hits + radix_sort_128x (map.c), mg_lchain_dp /
// mg_lchain_rmq (lchain.c), mm_gen_regs / mm_set_parent / mm_select_sub (hit.c) — the work done inside
// the `minimap2 -a -x <preset>` process the reference spawns at TELR_te.py:503-506.
//
// A persistent CTA takes one (locus, strand) at a time: it hashes that contig strand's minimizers into a
// shared-memory table once, then its 8 warps stream the locus's reads against it.
#pragma once
#include <cuda_runtime.h>
#include "mm_align.cuh"

namespace telr {

#ifndef TELR_CH_THREADS
#define TELR_CH_THREADS 1024
#endif
constexpr int CH_THREADS = TELR_CH_THREADS;
constexpr int CH_WARPS = CH_THREADS / 32;
// Contig minimizer index: open-addressing table (hash -> occurrence list).  Contigs with up to IDX_MAXMZ minimizers
// (~22 kb) use the shared-memory instance; longer ones (up to IDX_BIG_MAXMZ, ~360 kb) a per-CTA slice of global
// memory (L2-resident) with the same code.
template <int SLOTS, int MAXMZ, class U> struct IdxT {
    static constexpr int kSlots = SLOTS, kMax = MAXMZ;
    using slot_t = U;
    uint64_t keys[SLOTS];
    uint32_t cnt[SLOTS];
    uint32_t fill[SLOTS];
    U start[SLOTS];
    uint32_t occ_y[MAXMZ];
    U slot_of[MAXMZ];
    uint8_t occ_span[MAXMZ];
};
constexpr int IDX_MAXMZ = 4096, IDX_BIG_MAXMZ = 65536;
using IdxSmall = IdxT<8192, IDX_MAXMZ, uint16_t>;
using IdxBig = IdxT<131072, IDX_BIG_MAXMZ, uint32_t>;
struct IdxSmem {
    IdxSmall t;
    int hist[260];
    int ws[40];
    int mid_occ, n_keys, item;
};

struct ChainArgs {
    Opt o;
    int32_t n_loci, n_reads, mode;                 // mode 0: count anchors, 1: fill + chain
    const int32_t *locus_read_begin;
    const int64_t *mz_off; const uint64_t *mz_x; const uint32_t *mz_y; const uint16_t *selfcnt;
    const int32_t *read_len; const uint32_t *read_hash;
    int32_t *prob_na;                              // [n_prob]
    int32_t *prob_read, *prob_ls;                  // [n_prob]
    const int64_t *prob_aoff;                      // [n_prob+1]
    const int64_t *prob_roff;                      // [n_prob+1]   region capacity offsets
    Anchor *anchors; Reg *regs;
    int32_t *prob_nregs, *prob_nca;
    int32_t *prob_replen;                          // [n_prob] query bases under filtered high-occurrence seeds (MAPQ)
    uint8_t *warp_scratch; size_t warp_scratch_stride; int32_t max_na;   // per-warp: kept-minimizer list of the fill pass
    int32_t *work_counter; int32_t *err;
    int32_t *dp_counter, *rmq_counter;      // dynamic problem queues of k_chain_dp / k_chain_rmq (problem sizes vary 7x)
    unsigned long long *stat_anchors;
    // per-problem scratch for the chaining kernels (ChainScratch + HitScratch carved at prob_soff[p])
    uint8_t *prob_scratch; const int64_t *prob_soff;
    int32_t *prob_nu, *prob_m;      // chains found; anchors to re-chain (0 = no re-chaining)
    int32_t n_prob;
    IdxBig *idx_big;                // [gridDim.x] global-memory index slices for long contigs
    uint8_t *locus_bad;             // [n_loci] set when a contig exceeds IDX_BIG_MAXMZ minimizers: the locus is reported as unsupported (-4), the batch goes on
};

__device__ __forceinline__ void prob_carve(const ChainArgs &A, int p, int n_a, ChainScratch &cs, HitScratch &hs, int *cap_regs)
{
    uint8_t *b = A.prob_scratch + A.prob_soff[p];
    chain_scratch_carve(cs, b, (size_t)n_a + 1);
    *cap_regs = 2 * (n_a / 3) + 4;
    hit_scratch_carve(hs, b + chain_scratch_bytes((size_t)n_a + 1), (size_t)*cap_regs + 4);
}

__device__ __forceinline__ uint32_t idx_hash(uint64_t key) { return (uint32_t)(mix64(key) >> 24); }

// occurrences of a minimizer hash in the contig index
template <class IDX> __device__ __forceinline__ int idx_lookup(const IDX &I, uint64_t key, int *start)
{
    uint32_t s = idx_hash(key) & (IDX::kSlots - 1);
    for (;;) {
        uint64_t kk = I.keys[s];
        if (kk == key) { *start = (int)I.start[s]; return (int)I.cnt[s]; }
        if (kk == ~0ULL) return 0;
        s = (s + 1) & (IDX::kSlots - 1);
    }
}

template <class IDX> __device__ void idx_build(IDX &I, IdxSmem &C, const Opt &o, int n_c, const uint64_t *cx, const uint32_t *cy)
{
    const int tid = threadIdx.x;
    constexpr int SLOTS = IDX::kSlots;
    for (int i = tid; i < SLOTS; i += CH_THREADS) I.keys[i] = ~0ULL, I.cnt[i] = 0, I.fill[i] = 0;
    for (int i = tid; i < 260; i += CH_THREADS) C.hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_c; i += CH_THREADS) {
        const uint64_t key = cx[i] >> 8;
        uint32_t s = idx_hash(key) & (SLOTS - 1);
        for (;;) {
            unsigned long long old = atomicCAS((unsigned long long *)&I.keys[s], ~0ULL, (unsigned long long)key);
            if (old == ~0ULL || old == key) break;
            s = (s + 1) & (SLOTS - 1);
        }
        atomicAdd(&I.cnt[s], 1u);
        I.slot_of[i] = (typename IDX::slot_t)s;
    }
    __syncthreads();
    {   // exclusive scan of cnt over slots -> start
        int sum = 0;
        for (int c = 0; c < SLOTS / CH_THREADS; ++c) sum += (int)I.cnt[tid * (SLOTS / CH_THREADS) + c];
        int tot, pre = block_excl_scan_t<CH_THREADS>(sum, &tot, C.ws);
        for (int c = 0; c < SLOTS / CH_THREADS; ++c) {
            int s = tid * (SLOTS / CH_THREADS) + c;
            I.start[s] = (typename IDX::slot_t)pre;
            pre += (int)I.cnt[s];
        }
    }
    __syncthreads();
    for (int i = tid; i < n_c; i += CH_THREADS) {
        int s = (int)I.slot_of[i];
        int at = (int)I.start[s] + (int)atomicAdd(&I.fill[s], 1u);
        I.occ_y[at] = cy[i];
        I.occ_span[at] = (uint8_t)(cx[i] & 0xff);
    }
    __syncthreads();
    // lists in (span, position) order == the order minimap2's per-bucket sort leaves (buckets <= 64 entries)
    for (int s = tid; s < SLOTS; s += CH_THREADS) {
        int c = (int)I.cnt[s];
        if (c > 0) atomicAdd(&C.hist[c < 255 ? c : 255], 1);
        if (c > 1) {
            int b = (int)I.start[s];
            for (int i = 1; i < c; ++i) {
                uint32_t y = I.occ_y[b + i]; uint8_t sp = I.occ_span[b + i];
                uint64_t kk = (uint64_t)sp << 32 | y;
                int j = i;
                while (j > 0 && ((uint64_t)I.occ_span[b + j - 1] << 32 | I.occ_y[b + j - 1]) > kk) {
                    I.occ_y[b + j] = I.occ_y[b + j - 1], I.occ_span[b + j] = I.occ_span[b + j - 1];
                    --j;
                }
                I.occ_y[b + j] = y, I.occ_span[b + j] = sp;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {     // mm_idx_cal_max_occ + mm_mapopt_update
        int nk = 0;
        for (int c = 1; c <= 255; ++c) nk += C.hist[c];
        int mid = 0x7fffffff;
        if (nk > 0 && o.mid_occ_frac > 0.f) {
            uint32_t kk = (uint32_t)((1. - (double)o.mid_occ_frac) * nk);
            int acc = 0, c;
            for (c = 1; c <= 255; ++c) { acc += C.hist[c]; if ((uint32_t)acc > kk) break; }
            mid = c + 1;
        }
        if (mid < o.min_mid_occ) mid = o.min_mid_occ;
        if (o.max_mid_occ > o.min_mid_occ && mid > o.max_mid_occ) mid = o.max_mid_occ;
        C.mid_occ = mid; C.n_keys = nk;
    }
    __syncthreads();
}

// ---- chaining DP: candidates of anchor i are scored 32 at a time; the sequential early-exit rule
// (skip counter driven by t[] marks) is reproduced with three warp scans ----
__device__ void chain_dp_warp(const Opt &o, int n, const Anchor *a, ChainScratch &s)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    int32_t *f = s.f, *p = s.p, *v = s.v, *t = s.t;
    const int max_dist_x = tmax(o.max_gap, o.bw), max_dist_y = max_dist_x;
    for (int i = lane; i < n; i += 32) t[i] = 0;
    __syncwarp();
    int st = 0, max_ii = -1;
    for (int i = 0; i < n; ++i) {
        const Anchor ai = a[i];
        while (st < i && (ai.x >> 32 != a[st].x >> 32 || ai.x > a[st].x + (uint64_t)max_dist_x)) ++st;
        if (i - st > o.max_chain_iter) st = i - o.max_chain_iter;
        int32_t max_f = (int32_t)(ai.y >> 32 & 0xff);
        int max_j = -1, n_skip = 0, end_j = st - 1;
        for (int jb = i - 1; jb >= st; jb -= 32) {
            const int j = jb - lane;
            int32_t sc = INT32_MIN; int pj = -1;
            if (j >= st) {
                sc = link_score(ai, a[j], max_dist_x, max_dist_y, o.bw, o.chn_pen_gap, o.chn_pen_skip);
                if (sc != INT32_MIN) { sc += f[j]; pj = p[j]; }
            }
            const bool valid = sc != INT32_MIN;
            if (valid && pj >= 0) t[pj] = i;
            __syncwarp();
            const bool tj = valid && t[j] == i;
            // exclusive prefix max of candidate scores
            int32_t inc = valid ? sc : INT32_MIN;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int32_t y = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc = inc > y ? inc : y; }
            int32_t exc = __shfl_up_sync(FULL, inc, 1);
            if (lane == 0) exc = INT32_MIN;
            const int32_t pm = exc > max_f ? exc : max_f;
            const bool is_new = valid && sc > pm;
            const bool is_skip = valid && !is_new && tj;
            // skip counter: N_k = max(N_{k-1} + d_k, 0)  ==  P_k - min(-N_0, min_{m<=k} P_m)
            // prefix sum of (+1 skip, -1 new) from two ballots; the prefix minimum is only needed lane by lane when the
            // counter can pass the limit inside this block of 32 candidates, otherwise one warp reduction gives the carry
            const unsigned skm = __ballot_sync(FULL, is_skip), nwm = __ballot_sync(FULL, is_new), le = (2u << lane) - 1u;
            const int P = __popc(skm & le) - __popc(nwm & le);
            int N; unsigned brk = 0;
            if (n_skip + __popc(skm) <= o.max_chain_skip) {
                const int Mall = __reduce_min_sync(FULL, P);
                N = P - (Mall < -n_skip ? Mall : -n_skip);      // exact on lane 31, which is the only lane that uses it
            } else {
                int M = P;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(FULL, M, d); if (lane >= d) M = M < y ? M : y; }
                N = P - (M < -n_skip ? M : -n_skip);
                brk = __ballot_sync(FULL, is_skip && N > o.max_chain_skip);
            }
            const int bl = brk ? __ffs(brk) - 1 : 32;
            const unsigned newm = __ballot_sync(FULL, is_new) & (bl >= 32 ? FULL : ((1u << bl) - 1));
            if (newm) {
                int ln = 31 - __clz(newm);
                max_f = __shfl_sync(FULL, sc, ln);
                max_j = jb - ln;
            }
            if (brk) { end_j = jb - bl; break; }
            n_skip = __shfl_sync(FULL, N, 31);
        }
        if (max_ii < 0 || ai.x - a[max_ii].x > (uint64_t)(int64_t)max_dist_x) {
            int32_t bf = INT32_MIN; int bj = -1;       // argmax f over [st, i-1], ties to the largest j
            for (int j = i - 1 - lane; j >= st; j -= 32) { int32_t fj = f[j]; if (fj > bf) bf = fj, bj = j; }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                int32_t of = __shfl_xor_sync(FULL, bf, d); int oj = __shfl_xor_sync(FULL, bj, d);
                if (oj >= 0 && (bj < 0 || of > bf || (of == bf && oj > bj))) bf = of, bj = oj;
            }
            max_ii = bj;
        }
        if (max_ii >= 0 && max_ii < end_j) {
            int32_t tmp = link_score(ai, a[max_ii], max_dist_x, max_dist_y, o.bw, o.chn_pen_gap, o.chn_pen_skip);
            if (tmp != INT32_MIN && max_f < tmp + f[max_ii]) max_f = tmp + f[max_ii], max_j = max_ii;
        }
        if (lane == 0) {
            f[i] = max_f, p[i] = max_j;
            v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
        }
        if (max_ii < 0 || (ai.x - a[max_ii].x <= (uint64_t)(int64_t)max_dist_x && f[max_ii] < max_f)) max_ii = i;
        __syncwarp();
    }
}

// ---- re-chaining pass (mg_lchain_rmq).  The outer range-minimum query is a warp-wide scan over the active
// anchors; the inner tree (anchors within rmq_inner_dist on the target) is kept as an array W sorted by
// (query position, index), maintained incrementally with warp-parallel shifts, and scanned downwards 32
// candidates at a time with the same prefix-scan emulation of the skip rule as the first pass. ----
__device__ __forceinline__ bool key_lt(const Anchor *a, int j1, int32_t y2, int j2)   // (y_j1, j1) < (y2, j2)
{
    int32_t y1 = (int32_t)a[j1].y;
    return y1 < y2 || (y1 == y2 && j1 < j2);
}
// number of elements of W with key < (y, j)
__device__ __forceinline__ int w_lower(const int32_t *W, int nw, const Anchor *a, int32_t y, int j)
{
    int lo = 0, hi = nw;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (key_lt(a, W[mid], y, j)) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ void w_insert(int32_t *W, int &nw, const Anchor *a, int j)
{
    const int lane = threadIdx.x & 31;
    const int pos = w_lower(W, nw, a, (int32_t)a[j].y, j);
    for (int top = nw - 1; top >= pos; top -= 32) {      // shift [pos, nw) right by one, highest chunk first
        int k = top - lane, v = 0;
        if (k >= pos) v = W[k];
        __syncwarp();
        if (k >= pos) W[k + 1] = v;
        __syncwarp();
    }
    if (lane == 0) W[pos] = j;
    ++nw;
    __syncwarp();
}
__device__ void w_erase(int32_t *W, int &nw, const Anchor *a, int j)
{
    const int lane = threadIdx.x & 31;
    const int pos = w_lower(W, nw, a, (int32_t)a[j].y, j);
    if (pos >= nw || W[pos] != j) return;
    for (int b = pos + 1; b < nw; b += 32) {             // shift (pos, nw) left by one, lowest chunk first
        int k = b + lane, v = 0;
        if (k < nw) v = W[k];
        __syncwarp();
        if (k < nw) W[k - 1] = v;
        __syncwarp();
    }
    --nw;
    __syncwarp();
}

__device__ void chain_rmq_warp(const Opt &o, int n, const Anchor *a, ChainScratch &s)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    int32_t *f = s.f, *p = s.p, *v = s.v, *t = s.t, *W = s.ord;
    for (int i = lane; i < n; i += 32) t[i] = -1, v[i] = 0;
    __syncwarp();
    RmqWin w;
    rmq_win_init(w, o);
    int nw = 0;
    // Outer range query.  Upstream asks a balanced tree for the best priority among the anchors in the window whose y lies in
    // (yi - max_dist, yi); the first version here scanned the whole window for every anchor (O(n^2 / 32) per problem, 31 % of a
    // map-hifi step).  Now the anchors are blocked by index, 32 per block, and every block keeps a summary of its ACTIVE anchors
    // (inserted, not yet erased): the best anchor by the query's own total order (priority, then larger y, then larger index) and
    // bounds on y.  A query takes the summary of a block whose y range lies inside the query range, skips a block whose range lies
    // outside, and looks at the 32 anchors of the few blocks that straddle; the answer is the maximum of the same set under the
    // same order, hence identical.  Bounds are only ever loosened by erasures (a loose bound costs a look, never a result).
    const int nblk = (n + 31) >> 5;
    int32_t *bsum = reinterpret_cast<int32_t *>(s.z);          // [nblk][4]: best anchor (-1 none), min y, max y of the active anchors
    for (int b = lane; b < nblk; b += 32) { bsum[4 * b] = -1; bsum[4 * b + 1] = INT32_MAX; bsum[4 * b + 2] = INT32_MIN; }
    __syncwarp();
    int act_lo = 0, act_hi = 0;                                 // active anchors: [act_lo, act_hi)
    for (int i = 0; i < n; ++i) {
        // ---- window maintenance (same order of operations as the two trees upstream: insert, then erase) ----
        {
            const int old_i0 = w.i0, old_sti = w.st_inner;
            rmq_win_advance(w, o, i, a);
            if (w.max_dist_inner > 0) {
                for (int j = old_i0; j < w.i0; ++j) w_insert(W, nw, a, j);
                for (int j = old_sti; j < w.st_inner; ++j) w_erase(W, nw, a, j);
            }
            const int new_lo = w.st < w.i0 ? w.st : w.i0, new_hi = w.i0;
            if (lane == 0)
                for (int j = act_hi; j < new_hi; ++j) {        // insertions: fold the anchor into its block's summary
                    int32_t *bs = bsum + 4 * (j >> 5);
                    const int32_t yj = (int32_t)a[j].y;
                    const int cur = bs[0];
                    if (cur < 0 || rmq_better(rmq_pri(a[j], f[j], o.chn_pen_gap), yj, j, rmq_pri(a[cur], f[cur], o.chn_pen_gap), (int32_t)a[cur].y, cur)) bs[0] = j;
                    if (yj < bs[1]) bs[1] = yj;
                    if (yj > bs[2]) bs[2] = yj;
                }
            __syncwarp();
            if (new_lo > act_lo) {                              // erasures: rebuild the summaries of the blocks that lost anchors
                for (int b = act_lo >> 5; b <= (new_lo - 1) >> 5 && b < nblk; ++b) {
                    const int j = (b << 5) + lane;
                    const bool act = j >= new_lo && j < new_hi;
                    int bj = act ? j : -1; double bp = 0.0; int32_t by = act ? (int32_t)a[j].y : 0;
                    if (act) bp = rmq_pri(a[j], f[j], o.chn_pen_gap);
                    int32_t mn = act ? by : INT32_MAX, mx = act ? by : INT32_MIN;
#pragma unroll
                    for (int d = 16; d; d >>= 1) {
                        int ob = __shfl_xor_sync(FULL, bj, d); double op = __shfl_xor_sync(FULL, bp, d); int32_t oy = __shfl_xor_sync(FULL, by, d);
                        if (ob >= 0 && (bj < 0 || rmq_better(op, oy, ob, bp, by, bj))) bj = ob, bp = op, by = oy;
                    }
                    mn = __reduce_min_sync(FULL, mn); mx = __reduce_max_sync(FULL, mx);
                    if (lane == 0) { bsum[4 * b] = bj; bsum[4 * b + 1] = mn; bsum[4 * b + 2] = mx; }
                }
                __syncwarp();
            }
            act_lo = new_lo; act_hi = new_hi;
        }
        const Anchor ai = a[i];
        const int32_t yi = (int32_t)ai.y;
        // ---- outer RMQ over the block summaries ----
        int best = -1; double bp = 0.0; int32_t by = 0;
        if (act_hi > act_lo) {
            const int32_t ylo = yi - w.max_dist;
            const int b0 = act_lo >> 5, b1 = (act_hi - 1) >> 5;
            for (int bb = b0; bb <= b1; bb += 32) {
                const int b = bb + lane;
                bool look = false;
                if (b <= b1) {
                    const int cj = bsum[4 * b]; const int32_t mn = bsum[4 * b + 1], mx = bsum[4 * b + 2];
                    if (cj >= 0 && !(mn > yi || mx <= ylo)) {
                        if (mx < yi && mn > ylo) {              // the whole block qualifies: its summary answers for it
                            const double pri = rmq_pri(a[cj], f[cj], o.chn_pen_gap);
                            const int32_t cy = (int32_t)a[cj].y;
                            if (best < 0 || rmq_better(pri, cy, cj, bp, by, best)) best = cj, bp = pri, by = cy;
                        } else look = true;
                    }
                }
                unsigned lm = __ballot_sync(FULL, look);
                while (lm) {                                    // blocks that straddle the query range: their 32 anchors
                    const int lb = bb + __ffs(lm) - 1;
                    lm &= lm - 1;
                    const int j = (lb << 5) + lane;
                    if (j >= act_lo && j < act_hi) {
                        const Anchor aj = a[j];
                        if (rmq_in_range(aj, j, yi, w.max_dist)) {
                            const double pri = rmq_pri(aj, f[j], o.chn_pen_gap);
                            if (best < 0 || rmq_better(pri, (int32_t)aj.y, j, bp, by, best)) best = j, bp = pri, by = (int32_t)aj.y;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            int ob = __shfl_xor_sync(FULL, best, d); double op = __shfl_xor_sync(FULL, bp, d); int32_t oy = __shfl_xor_sync(FULL, by, d);
            if (ob >= 0 && (best < 0 || rmq_better(op, oy, ob, bp, by, best))) best = ob, bp = op, by = oy;
        }
        int max_j = -1;
        int32_t max_f = (int32_t)(ai.y >> 32 & 0xff);
        if (best >= 0) {
            int exact, width;
            int32_t sc = f[best] + link_score_simple(ai, a[best], o.chn_pen_gap, o.chn_pen_skip, &exact, &width);
            if (width <= o.bw_long && sc > max_f) max_f = sc, max_j = best;
            if (!exact && w.max_dist_inner > 0 && yi > 0) {
                // candidates: keys <= (yi - 1, n), i.e. y <= yi - 1; walk down from the largest
                int top = w_lower(W, nw, a, yi, -1) - 1;         // (y, j) < (yi, -1)  <=>  y <= yi - 1
                int n_skip = 0;
                const int32_t ylo = yi - w.max_dist_inner;
                for (; top >= 0; top -= 32) {
                    const int k = top - lane;
                    int j = -1; int32_t scj = INT32_MIN; int pj = -1; bool inr = false;
                    if (k >= 0) {
                        j = W[k];
                        inr = (int32_t)a[j].y >= ylo;
                        if (inr) {
                            int w2;
                            int32_t sj = f[j] + link_score_simple(ai, a[j], o.chn_pen_gap, o.chn_pen_skip, 0, &w2);
                            if (w2 <= o.bw_long) scj = sj, pj = p[j];
                        }
                    }
                    const unsigned inm = __ballot_sync(FULL, inr);        // in-range lanes form a prefix (W is sorted)
                    const bool valid = scj != INT32_MIN;
                    if (valid && pj >= 0) t[pj] = i;
                    __syncwarp();
                    const bool tj = valid && t[j] == i;
                    int32_t inc = valid ? scj : INT32_MIN;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { int32_t yv = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc = inc > yv ? inc : yv; }
                    int32_t exc = __shfl_up_sync(FULL, inc, 1);
                    if (lane == 0) exc = INT32_MIN;
                    const int32_t pm = exc > max_f ? exc : max_f;
                    const bool is_new = valid && scj > pm;
                    const bool is_skip = valid && !is_new && tj;
                    const unsigned skm = __ballot_sync(FULL, is_skip), nwm = __ballot_sync(FULL, is_new), le = (2u << lane) - 1u;
                    const int P = __popc(skm & le) - __popc(nwm & le);
                    int N; unsigned brk = 0;
                    if (n_skip + __popc(skm) <= o.max_chain_skip) {
                        const int Mall = __reduce_min_sync(FULL, P);
                        N = P - (Mall < -n_skip ? Mall : -n_skip);
                    } else {
                        int M = P;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) { int yv = __shfl_up_sync(FULL, M, d); if (lane >= d) M = M < yv ? M : yv; }
                        N = P - (M < -n_skip ? M : -n_skip);
                        brk = __ballot_sync(FULL, is_skip && N > o.max_chain_skip);
                    }
                    const int bl = brk ? __ffs(brk) - 1 : 32;
                    const unsigned newm = __ballot_sync(FULL, is_new) & (bl >= 32 ? FULL : ((1u << bl) - 1));
                    if (newm) {
                        int ln = 31 - __clz(newm);
                        max_f = __shfl_sync(FULL, scj, ln);
                        max_j = __shfl_sync(FULL, j, ln);
                    }
                    if (brk) break;
                    if (inm != FULL) break;                                // ran past the y window (or the array)
                    n_skip = __shfl_sync(FULL, N, 31);
                }
            }
        }
        if (lane == 0) {
            f[i] = max_f, p[i] = max_j;
            v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
        }
        __syncwarp();
    }
}

// one (locus, strand) item: build the index, then stream the locus's reads against it (warp per read)
template <class IDX>
__device__ void chain_item(const ChainArgs &A, IDX &I, IdxSmem &C, int item, int rb, int nr, int strand, int64_t cb, int n_c)
{
    const Opt &o = A.o;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    (void)tid;
    idx_build(I, C, o, n_c, A.mz_x + cb, A.mz_y + cb);
    const int mid_occ = C.mid_occ;
    for (int r = wid; r < nr; r += CH_WARPS) {
        const int read = rb + r, pidx = 2 * rb + strand * nr + r;
        const int64_t qb = A.mz_off[read];
        const int n = (int)(A.mz_off[read + 1] - qb);
        const uint64_t *qx = A.mz_x + qb; const uint32_t *qy = A.mz_y + qb; const uint16_t *qc = A.selfcnt + qb;
        const int qlen = A.read_len[read];
        const bool do_flt = o.q_occ_frac > 0.0f && mid_occ > 0 && n > mid_occ;
        const float thr = (float)n * o.q_occ_frac;
        // per-warp scratch: minimizers that survive the query-occurrence filter, their occurrences in the index, a work list
        uint8_t *wsb = A.warp_scratch + (size_t)(blockIdx.x * CH_WARPS + wid) * A.warp_scratch_stride;
        const size_t ws_cap = A.warp_scratch_stride / 12;
        int32_t *kept = (int32_t *)wsb, *tarr = kept + ws_cap, *cidx = tarr + ws_cap;
        int nk = 0;
        for (int ib = 0; ib < n; ib += 32) {
            int i = ib + lane;
            bool keep = false;
            if (i < n) { int c = qc[i]; keep = !(do_flt && c > mid_occ && (float)c > thr); }
            unsigned m = __ballot_sync(FULL, keep);
            if (keep) kept[nk + __popc(m & ((1u << lane) - 1))] = i;
            nk += __popc(m);
        }
        __syncwarp();
        bool hi = false;
        for (int j = lane; j < nk; j += 32) {
            int st, t = idx_lookup(I, qx[kept[j]] >> 8, &st);
            tarr[j] = t;
            hi |= t > mid_occ;
        }
        hi = __any_sync(FULL, hi);
        int rep_len = 0;
        if (hi) {       // rare: streaks of high-occurrence seeds are thinned out sequentially (mm_seed_select)
            __syncwarp();
            if (lane == 0) rep_len = seeds_filter(mid_occ, o.max_max_occ, o.occ_dist, nk, kept, qx, qy, tarr, cidx, qlen);
            rep_len = __shfl_sync(FULL, rep_len, 0);
            __syncwarp();
        }
        if (A.mode == 0) {
            int na = 0;
            for (int j = lane; j < nk; j += 32) na += tarr[j] > 0 ? tarr[j] : 0;
#pragma unroll
            for (int d = 16; d; d >>= 1) na += __shfl_xor_sync(FULL, na, d);
            if (lane == 0) { A.prob_na[pidx] = na; A.prob_read[pidx] = read; A.prob_ls[pidx] = item; }
            __syncwarp();
            continue;
        }
        // ---------------- fill mode ----------------
        if (lane == 0) A.prob_replen[pidx] = rep_len;
        const int64_t ao = A.prob_aoff[pidx];
        const int n_a = (int)(A.prob_aoff[pidx + 1] - ao);
        Anchor *a = A.anchors + ao;
        if (n_a == 0) { __syncwarp(); continue; }
        int run = 0;
        for (int jb = 0; jb < nk; jb += 32) {
            int j = jb + lane, t = 0, st = 0; uint64_t x = 0; uint32_t y = 0; bool tandem = false;
            if (j < nk) {
                int i = kept[j];
                x = qx[i]; y = qy[i];
                t = tarr[j] > 0 ? tarr[j] : 0;
                if (t) {
                    idx_lookup(I, x >> 8, &st);
                    if (j > 0 && (qx[kept[j - 1]] >> 8) == (x >> 8)) tandem = true;
                    if (j < nk - 1 && (qx[kept[j + 1]] >> 8) == (x >> 8)) tandem = true;
                }
            }
            int pre = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int yv = __shfl_up_sync(FULL, pre, d); if (lane >= d) pre += yv; }
            int tot = __shfl_sync(FULL, pre, 31);
            pre -= t;
            const uint32_t q_span = (uint32_t)(x & 0xff), q_pos = y;
            for (int k = 0; k < t; ++k) {
                uint32_t ry = I.occ_y[st + k];
                int32_t rpos = (int32_t)(ry >> 1);
                Anchor an;
                if ((ry & 1) == (q_pos & 1)) {
                    an.x = (uint64_t)(uint32_t)rpos;
                    an.y = (uint64_t)q_span << 32 | (q_pos >> 1);
                } else {
                    an.x = 1ULL << 63 | (uint64_t)(uint32_t)rpos;
                    an.y = (uint64_t)q_span << 32 | (uint32_t)(qlen - ((int32_t)(q_pos >> 1) + 1 - (int32_t)q_span) - 1);
                }
                if (tandem) an.y |= SEED_TANDEM;
                a[run + pre + k] = an;
            }
            run += tot;
        }
        if (lane == 0) atomicAdd(A.stat_anchors, (unsigned long long)n_a);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(CH_THREADS) k_chain(const __grid_constant__ ChainArgs A)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    IdxSmem &C = *reinterpret_cast<IdxSmem *>(smem_raw);
    const Opt &o = A.o;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const int n_items = 2 * A.n_loci;
    for (;;) {
        if (tid == 0) C.item = atomicAdd(A.work_counter, 1);
        __syncthreads();
        const int item = C.item;
        __syncthreads();
        if (item >= n_items) break;
        const int l = item >> 1, strand = item & 1;
        const int rb = A.locus_read_begin[l], nr = A.locus_read_begin[l + 1] - rb;
        const int cseq = A.n_reads + strand * A.n_loci + l;
        const int64_t cb = A.mz_off[cseq];
        const int n_c = (int)(A.mz_off[cseq + 1] - cb);
        if (n_c > IDX_BIG_MAXMZ || (n_c > IDX_MAXMZ && !A.idx_big)) {
            if (tid == 0) { if (A.locus_bad) A.locus_bad[l] = 1; else atomicOr(A.err, 16); }
            for (int r = tid; r < nr; r += CH_THREADS) {
                int pidx = 2 * rb + strand * nr + r;
                if (A.mode == 0) A.prob_na[pidx] = 0, A.prob_read[pidx] = rb + r, A.prob_ls[pidx] = item;
                else A.prob_nregs[pidx] = 0, A.prob_nca[pidx] = 0;
            }
            continue;
        }
        if (n_c <= IDX_MAXMZ) chain_item(A, C.t, C, item, rb, nr, strand, cb, n_c);
        else chain_item(A, A.idx_big[blockIdx.x], C, item, rb, nr, strand, cb, n_c);
        __syncthreads();
    }
}

// ---- sequential steps, one THREAD per problem (latency hidden by thread-level parallelism) ----
// Sequential per-problem kernels: only every TP_STRIDE-th lane owns a problem.  A warp of 32 unrelated sequential
// jobs executes the union of their control paths (6 of 32 threads active on average, ncu); spreading the jobs over
// 32 / TP_STRIDE times more warps trades idle lanes, which these latency-bound kernels do not miss, for divergence.
constexpr int TP_CHAIN = 32;      // k_chain_sort / k_chain_bt / k_chain_regs: one problem per warp (seed+chain stage 48.0 -> 33.0 ms at 296 loci)
constexpr int TP_FINISH = 8;      // k_al_finish: four problems per warp (11.4 -> 7.5 ms)
template <int STRIDE> __host__ __device__ __forceinline__ int tp_blocks(int n_prob, int threads) { return (int)(((long long)n_prob * STRIDE + threads - 1) / threads); }
template <int STRIDE> __device__ __forceinline__ int tp_problem(int n_prob)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t % STRIDE) return -1;
    const int p = t / STRIDE;
    return p < n_prob ? p : -1;
}

// Warp-cooperative form of rs_sort_emul for anchors keyed by x (same permutation, element for element): up to 64 elements a stable
// rank sort (what the insertion sort of klib leaves); above that the MSD radix passes with the level skip, the histogram and
// the bucket offsets computed by the 32 lanes, the unstable cycle-leader pass on lane 0 exactly as upstream walks it (its
// deposit order is what fixes the order of equal keys), and the insertion sorts of the small buckets one bucket per lane.
struct SortSmem { int32_t bb[256], be[256]; };
__device__ void rs_sort_warp(Anchor *a, int n, int32_t *scratch, SortSmem &H)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    if (n <= 64) {
        if (n < 2) return;
        uint4 *tmp = reinterpret_cast<uint4 *>(&H);
        for (int e = lane; e < n; e += 32) tmp[e] = *reinterpret_cast<const uint4 *>(&a[e]);
        __syncwarp();
        for (int e = lane; e < n; e += 32) {
            const uint4 me = tmp[e];
            const uint64_t ke = (uint64_t)me.y << 32 | me.x;
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const uint64_t kj = (uint64_t)tmp[j].y << 32 | tmp[j].x;
                rank += kj < ke || (kj == ke && j < e);
            }
            *reinterpret_cast<uint4 *>(&a[rank]) = me;
        }
        __syncwarp();
        return;
    }
    int32_t *stk = scratch + 512;
    if (lane == 0) { stk[0] = 0; stk[1] = n; stk[2] = 56; }
    __syncwarp();
    int sp = 1;
    while (sp > 0) {
        --sp;
        const int beg = stk[3 * sp], end = stk[3 * sp + 1];
        int s = stk[3 * sp + 2];
        __syncwarp();
        for (;;) {                      // levels on which every key has the same digit are no-ops upstream
            const uint64_t k0 = a[beg].x >> s & 255;
            bool diff = false;
            for (int i = beg + 1 + lane; i < end; i += 32) diff |= (a[i].x >> s & 255) != k0;
            if (__any_sync(FULL, diff) || s == 0) break;
            s -= 8;
        }
        for (int k = lane; k < 256; k += 32) H.be[k] = 0;
        __syncwarp();
        for (int i = beg + lane; i < end; i += 32) atomicAdd(&H.be[(int)(a[i].x >> s & 255)], 1);
        __syncwarp();
        {
            int c[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = H.be[8 * lane + j]; sum += c[j]; }
            int pre = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(FULL, pre, d); if (lane >= d) pre += y; }
            int acc = beg + pre - sum;
#pragma unroll
            for (int j = 0; j < 8; ++j) { H.bb[8 * lane + j] = acc; acc += c[j]; H.be[8 * lane + j] = acc; }
        }
        __syncwarp();
        if (lane == 0) {                // the cycle-leader permutation, in upstream's order
            for (int k = 0; k < 256;) {
                if (H.bb[k] != H.be[k]) {
                    int l = (int)(a[H.bb[k]].x >> s & 255);
                    if (l != k) {
                        Anchor tmp = a[H.bb[k]], swp;
                        do {
                            swp = tmp; tmp = a[H.bb[l]]; a[H.bb[l]++] = swp;
                            l = (int)(tmp.x >> s & 255);
                        } while (l != k);
                        a[H.bb[k]++] = tmp;
                    } else ++H.bb[k];
                } else ++k;
            }
        }
        __syncwarp();
        if (s) {
            for (int k = lane; k < 256; k += 32) {          // small buckets: one per lane at a time
                const int b0 = k ? H.be[k - 1] : beg, sz = H.be[k] - b0;
                if (sz > 1 && sz <= 64) ins_sort(a + b0, sz, KeyX());
            }
            if (lane == 0)
                for (int k = 0; k < 256; ++k) {
                    const int b0 = k ? H.be[k - 1] : beg, sz = H.be[k] - b0;
                    if (sz > 64) { stk[3 * sp] = b0; stk[3 * sp + 1] = H.be[k]; stk[3 * sp + 2] = s - 8; ++sp; }
                }
            sp = __shfl_sync(FULL, sp, 0);
        }
        __syncwarp();
    }
}

// Warp forms of chain_backtrack / chain_compact (mm_chain.cuh): the filter, the sort, the clearing and the copies are spread
// over the lanes; the walk that extracts the chains stays sequential on lane 0 (every step depends on the marks of the last).
__device__ void chain_backtrack_warp(int n, ChainScratch &s, int min_cnt, int min_sc, int max_drop, int *n_u_, int *n_v_, SortSmem &H)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int32_t *f = s.f, *p = s.p;
    int32_t *v = s.v, *t = s.t;
    Anchor *z = s.z;
    int n_z = 0;
    *n_u_ = *n_v_ = 0;
    for (int ib = 0; ib < n; ib += 32) {
        const int i = ib + lane;
        const bool keep = i < n && f[i] >= min_sc;
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) { Anchor e; e.x = (uint64_t)f[i]; e.y = (uint64_t)i; z[n_z + __popc(m & ((1u << lane) - 1))] = e; }
        n_z += __popc(m);
    }
    __syncwarp();
    if (n_z == 0) return;
    rs_sort_warp(z, n_z, s.sortws, H);
    for (int i = lane; i < n; i += 32) t[i] = 0;
    __syncwarp();
    int n_u = 0, n_v = 0;
    if (lane == 0) {
        for (int k = n_z - 1; k >= 0; --k) {
            if (t[z[k].y] == 0) {
                int n_v0 = n_v;
                int32_t end_i = bk_end(max_drop, z, f, p, t, k), i;
                for (i = (int32_t)z[k].y; i != end_i; i = p[i]) v[n_v++] = i, t[i] = 1;
                int32_t sc = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
                if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) s.u[n_u++] = (uint64_t)sc << 32 | (uint32_t)(n_v - n_v0);
                else n_v = n_v0;
            }
        }
    }
    __syncwarp();
    *n_u_ = __shfl_sync(FULL, n_u, 0); *n_v_ = __shfl_sync(FULL, n_v, 0);
}

__device__ void chain_compact_warp(int n_u, ChainScratch &s, Anchor *a, SortSmem &H)
{
    const int lane = threadIdx.x & 31;
    Anchor *b = s.b, *w = s.z;
    uint64_t *u = s.u, *u2 = s.u2;
    int k = 0;
    for (int i = 0; i < n_u; ++i) {             // chains listed end -> start become start -> end
        const int ni = (int32_t)u[i];
        for (int j = lane; j < ni; j += 32) b[k + j] = a[s.v[k + (ni - j - 1)]];
        k += ni;
    }
    __syncwarp();
    if (lane == 0) {
        int kk = 0;
        for (int i = 0; i < n_u; ++i) { w[i].x = b[kk].x; w[i].y = (uint64_t)kk << 32 | (uint32_t)i; kk += (int32_t)u[i]; }
    }
    __syncwarp();
    rs_sort_warp(w, n_u, s.sortws, H);
    k = 0;
    for (int i = 0; i < n_u; ++i) {             // chains by target position
        const int j = (int32_t)w[i].y, ni = (int32_t)u[j];
        const Anchor *src = &b[w[i].y >> 32];
        for (int c = lane; c < ni; c += 32) a[k + c] = src[c];
        if (lane == 0) u2[i] = u[j];
        k += ni;
    }
    __syncwarp();
    for (int i = lane; i < n_u; i += 32) u[i] = u2[i];
    __syncwarp();
}

__global__ void __launch_bounds__(128) k_chain_sort(const __grid_constant__ ChainArgs A)
{
    __shared__ SortSmem HS[4];
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= A.n_prob) return;
    const int n_a = (int)(A.prob_aoff[p + 1] - A.prob_aoff[p]);
    if (lane == 0) { A.prob_nu[p] = 0; A.prob_m[p] = 0; A.prob_nregs[p] = 0; A.prob_nca[p] = 0; }
    if (n_a == 0) return;
    ChainScratch cs; HitScratch hs; int cap;
    prob_carve(A, p, n_a, cs, hs, &cap);
    rs_sort_warp(A.anchors + A.prob_aoff[p], n_a, cs.sortws, HS[threadIdx.x >> 5]);
}

// chaining DP, one WARP per problem
__global__ void __launch_bounds__(256) k_chain_dp(const __grid_constant__ ChainArgs A)
{
    const int lane = threadIdx.x & 31;
    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(A.dp_counter, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= A.n_prob) break;
        const int n_a = (int)(A.prob_aoff[p + 1] - A.prob_aoff[p]);
        if (n_a == 0) continue;
        ChainScratch cs; HitScratch hs; int cap;
        prob_carve(A, p, n_a, cs, hs, &cap);
        chain_dp_warp(A.o, n_a, A.anchors + A.prob_aoff[p], cs);
    }
}

__global__ void __launch_bounds__(128) k_chain_bt(const __grid_constant__ ChainArgs A)
{
    __shared__ SortSmem HS[4];
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= A.n_prob) return;
    const int n_a = (int)(A.prob_aoff[p + 1] - A.prob_aoff[p]);
    if (n_a == 0) return;
    const Opt &o = A.o;
    SortSmem &H = HS[threadIdx.x >> 5];
    ChainScratch cs; HitScratch hs; int cap;
    prob_carve(A, p, n_a, cs, hs, &cap);
    Anchor *a = A.anchors + A.prob_aoff[p];
    const int qlen = A.read_len[A.prob_read[p]];
    int n_u = 0, n_v = 0, m = 0;
    chain_backtrack_warp(n_a, cs, o.min_cnt, o.min_chain_score, o.bw, &n_u, &n_v, H);
    if (n_u > 0) {
        chain_compact_warp(n_u, cs, a, H);
        if (o.bw_long > o.bw && n_u > 1) {
            int32_t st = (int32_t)a[0].y, en = (int32_t)a[(int32_t)cs.u[0] - 1].y;
            if (qlen - (en - st) > o.rmq_rescue_size || en - st > qlen * o.rmq_rescue_ratio) {
                for (int i = 0; i < n_u; ++i) m += (int32_t)cs.u[i];
                rs_sort_warp(a, m, cs.sortws, H);
            }
        }
    }
    if (lane == 0) { A.prob_nu[p] = n_u; A.prob_m[p] = m; }
}

// re-chaining DP, one WARP per flagged problem
__global__ void __launch_bounds__(256) k_chain_rmq(const __grid_constant__ ChainArgs A)
{
    const int lane = threadIdx.x & 31;
    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(A.rmq_counter, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= A.n_prob) break;
        const int m = A.prob_m[p];
        if (m == 0) continue;
        const int n_a = (int)(A.prob_aoff[p + 1] - A.prob_aoff[p]);
        ChainScratch cs; HitScratch hs; int cap;
        prob_carve(A, p, n_a, cs, hs, &cap);
        chain_rmq_warp(A.o, m, A.anchors + A.prob_aoff[p], cs);
    }
}

__global__ void __launch_bounds__(128) k_chain_regs(const __grid_constant__ ChainArgs A)
{
    __shared__ SortSmem HS[4];
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= A.n_prob) return;
    const int n_a = (int)(A.prob_aoff[p + 1] - A.prob_aoff[p]);
    if (n_a == 0) return;
    const Opt &o = A.o;
    SortSmem &H = HS[threadIdx.x >> 5];
    ChainScratch cs; HitScratch hs; int cap_regs;
    prob_carve(A, p, n_a, cs, hs, &cap_regs);
    Anchor *a = A.anchors + A.prob_aoff[p];
    Reg *regs = A.regs + A.prob_roff[p];
    const int read = A.prob_read[p], qlen = A.read_len[read];
    int n_u = A.prob_nu[p], n_v = 0;
    const int m = A.prob_m[p];
    if (m > 0) {
        chain_backtrack_warp(m, cs, o.min_cnt, o.min_chain_score, o.bw_long, &n_u, &n_v, H);
        if (n_u > 0) chain_compact_warp(n_u, cs, a, H);
    }
    if (lane != 0) return;          // region bookkeeping: sequential
    int n_regs = 0, nca = 0;
    if (n_u > 0) {
        if (n_u > cap_regs) { atomicOr(A.err, TELR_ERR_REGCAP); n_u = cap_regs; }
        for (int i = 0; i < n_u; ++i) nca += (int32_t)cs.u[i];
        uint32_t hash = A.read_hash[read];
        hash ^= wang_hash32((uint32_t)qlen) + o.seed_term;
        hash = wang_hash32(hash);
        regs_from_chains(hash, qlen, n_u, cs.u, a, regs, hs);
        n_regs = n_u;
        regs_set_parent(o, n_regs, regs, hs);
        regs_select_sub(o, 1, &n_regs, regs, hs, cap_regs);
    }
    A.prob_nregs[p] = n_regs;
    A.prob_nca[p] = nca;
}

}  // namespace telr
