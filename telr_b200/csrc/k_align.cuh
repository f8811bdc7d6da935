// k_align.cuh — kernel (d): base-level alignment.  One warp owns one (read x contig strand) problem at a
// time (dynamic queue), runs the alignment coroutine of mm_align.cuh on lane 0 and executes every DP
// request warp-wide.
//
// Replaces ksw_extd2_sse / ksw_ll_i16 (minimap2 ksw2_extd2_sse.c, ksw2_ll_sse.c) and the driver around
// them (align.c) inside the `minimap2 -a` process the reference spawns at TELR_te.py:503-506.
// Integer DP on the ALU pipes; no tensor cores (this is not a dense contraction).
//
// warp_extd2 (general path): anti-diagonal sweep, 32 cells per step, difference recurrence
// (u,v,x,y,x2,y2 as int8 in an L1-resident per-warp scratch), direction bytes to the warp's traceback
// buffer, exact max tracking with the reference's tie order, z-drop, approximate-max mode.
#pragma once
#include <cuda_runtime.h>
#include "mm_align.cuh"
#include "k_fill.cuh"

namespace telr {

// The cycle census (TELR_CENSUS=1) costs 1.5 KB of code inside the hottest loop of an instruction-cache-bound kernel, so it is
// compiled in only with -DTELR_CENSUS_BUILD=1 (profiles/census.sh builds that variant).
#ifndef TELR_CENSUS_BUILD
#define TELR_CENSUS_BUILD 0
#endif
#define AL_CENSUS(A) (TELR_CENSUS_BUILD && (A).census)
constexpr int ALN_REC_INTS = 24;        // int32 words per alignment record written by k_al_finish
constexpr int AL_THREADS = 128;
constexpr int AL_WARPS = AL_THREADS / 32;
#ifndef TELR_AL_BLOCKS
#define TELR_AL_BLOCKS 4
#endif
constexpr int AL_BLOCKS_PER_SM = TELR_AL_BLOCKS;     // 16 resident warps per SM: the kernel is issue-bound (profiles/), more warps only add spills
constexpr int DPU = 4;                 // independent 32-cell chunks per lane per DP iteration
constexpr int DP_SCOLS = 1024;         // columns of DP state kept in shared memory per warp (power of two)
struct VecSmem;

struct DpScratch {
    int8_t *u, *v, *x, *y, *x2, *y2;   // [maxT] each
    int32_t *H;                        // [maxT]
    int32_t *ll;                       // [6 * maxT] local-probe rows
    uint8_t *dir; int64_t dir_cap;
    uint32_t *ezcig; int32_t ezcap;
    uint32_t *bnd;                     // fast fill path: pass-boundary values, 3 words per query row pair
    // shared-memory circular window of DP_SCOLS columns (used when the live band fits)
    int8_t *s_state; int32_t *s_H; const uint2 *stab;
    VecSmem *vsm;                      // same window plus staged sequence codes (vectorised path)
};

// ---- round-based alignment (BSP): plan (thread/problem) -> dp (warp/task) -> traceback (thread/task) ----
struct AlWork {                 // per work item (a problem with at least one chain)
    int32_t pidx, done, pending, has_task;
    int64_t dir_off;            // this round's slice of the pool (direction bytes or local-probe rows / spilled DP state)
    int64_t cig_off, ez_off;    // CIGAR arena and per-task CIGAR buffer (uint32 units) in AlignArgs::cigs
    int32_t cig_cap, ez_cap;
};

struct AlignArgs {
    Opt o;
    const Opt *d_opt;           // the same options in device global memory (AlnCtx::o outlives a kernel launch)
    int32_t n_prob, read_base, n_work;
    const int32_t *read_len;
    const int32_t *contig_len; const int64_t *ctg_boff; const uint8_t *ctg_bytes;
    const uint8_t *read_bytes; const int64_t *rbyte_off;       // nt4 bytes of every read: forward then reverse complement
    const int32_t *prob_read, *prob_ls, *prob_nca;
    int32_t *prob_nregs;
    const int64_t *prob_aoff, *prob_roff;
    Anchor *anchors; Reg *regs;
    uint8_t *prob_scratch; const int64_t *prob_soff;           // chaining scratch, reused (HitScratch, long-gap list)
    const int32_t *prob_replen;
    const int32_t *work_list;
    AlWork *work; AlnCtx *actx; DpTask *tasks; DpRes *res;
    uint32_t *cigs;
    uint8_t *warp_scratch; size_t warp_scratch_stride; int32_t max_tlen, max_qlen, use_fast, use_vec, census; int64_t dir_cap;    // per resident warp: spilled DP state + traceback bytes
    uint8_t *big; int64_t big_cap; int32_t n_big; int32_t *big_lock;                          // shared large traceback buffers
    unsigned long long *rc;      // [2] work queue head
    int32_t *err;
    unsigned long long *stat_cells, *stat_tasks;
    int2 *blocks; unsigned long long *n_blocks; int64_t blocks_cap;
    int64_t *prob_blk_off; int32_t *prob_blk_cnt;
    int32_t *aln_out; unsigned long long *n_aln; int64_t aln_cap;
    uint32_t *cig_out; unsigned long long *n_cig; int64_t cig_out_cap;
    // task queues of k_al_queue: one ring per role (0 gap fills, 1 everything else, 2 gap fills of the 12-column instance), entries = work item + 1
    int32_t *q_ring[3]; int32_t q_cap; int32_t ext_per8, wide_per8;
    int32_t *q_state;            // AQ_* counters
};
constexpr int AQ_ROLES = 3;
constexpr int AQ_MAX_SM = 1024;        // per-SM role words follow the counters in q_state
enum { AQ_HEAD0 = 0, AQ_TAIL0 = 3, AQ_AVAIL0 = 6, AQ_DONE = 9, AQ_START = 10, AQ_STEAL = 11, AQ_SLOTS = 16 };

// Bounded extension (exact work saving).  An extension only reports its maximum (max, max_t, max_q): the caller never reads
// zdropped of an extension, and with end_bonus <= 0 the query-end rule `mqe + end_bonus > max` cannot hold (mqe <= max).  A cell
// (i, j) of anti-diagonal r scores at most a * min(i + 1, j + 1) - gap(|i - j|) -- that many matches at best, and the offset
// between the coordinates has to be paid for by gaps, one gap being the cheapest way.  Once one sequence is exhausted
// |i - j| >= r - 2 (len - 1) grows with r; when the bound is not above the maximum found so far, no later anti-diagonal can
// change the result.  Typical case: a read overhanging the contig end by thousands of bases, where ksw2 keeps ~750
// anti-diagonals of a few dozen cells alive until the band runs out (11 % of the alignment kernel's time on config 2).
// The oracle applies the same rule (orc_ksw_extd2), so cell counts stay equal; tests show that no output depends on it.
__device__ __forceinline__ bool ext_bound_stop(int a, int q, int e, int q2, int e2, int qlen, int tlen, int r, int32_t ez_max)
{
    const int dt = r - 2 * (tlen - 1), dq = r - 2 * (qlen - 1), dmin = dt > dq ? dt : dq;
    if (dmin <= 0) return false;
    int mcap = tlen < qlen ? tlen : qlen;
    if ((r >> 1) + 1 < mcap) mcap = (r >> 1) + 1;
    const long long g1 = (long long)q + (long long)e * dmin, g2 = (long long)q2 + (long long)e2 * dmin;
    return (long long)a * mcap - (g1 < g2 ? g1 : g2) <= (long long)ez_max;
}

__device__ __forceinline__ int dp_base(const uint8_t *p, int step, int comp, int i)
{
    int b = p[(ptrdiff_t)i * step];
    return comp ? (b >= 4 ? 4 : 3 - b) : b;
}

}  // namespace telr
#include "k_extv.cuh"
namespace telr {
static_assert(sizeof(TbSmem) <= sizeof(((VecSmem *)0)->H), "fill traceback window aliases the idle H window");

struct EzPush {
    uint32_t *c; int n, cap;
    __device__ __forceinline__ void push(uint32_t op, int len)
    {
        if (n == 0 || op != (c[n - 1] & 0xf)) { if (n < cap) c[n] = (uint32_t)len << 4 | op; ++n; }
        else c[n - 1] += (uint32_t)len << 4;
    }
};

// two-piece affine extension / global DP, warp-wide.  R lives in shared memory.
template <bool SMEM>
__device__ void warp_extd2_impl(const Opt &o, const DpTask &T, DpRes &R, DpScratch &S, unsigned long long *cells_acc, int32_t *err)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int qlen = T.qlen, tlen = T.tlen, flag = T.flag;
    if (lane == 0) { res_reset(R); R.cigar = S.ezcig; }
    __syncwarp();
    if (qlen <= 0 || tlen <= 0) return;
    int q = o.q, e = o.e, q2 = o.q2, e2 = o.e2;
    if (q2 + e2 < q + e) { int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }
    const int qe = q + e, qe2 = q2 + e2;
    int w = T.w < 0 ? (tlen > qlen ? tlen : qlen) : T.w;
    {
        int min_sc = -o.b < -o.sc_ambi ? -o.b : -o.sc_ambi;
        if (-min_sc > 2 * (q + e)) return;
    }
    int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
    if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
    const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
    int ncol = qlen < tlen ? qlen : tlen;
    if (ncol > w + 1) ncol = w + 1;
    uint8_t *p = S.dir;
    if ((int64_t)(qlen + tlen - 1) * ncol > S.dir_cap) {
        if (lane == 0) { atomicOr(err, TELR_ERR_DIRCAP); R.zdropped = 1; }
        __syncwarp();
        return;
    }
    const bool approx = flag & KSW_APPROX_MAX, right = flag & KSW_RIGHT;
    // column t lives at slot (t & CM): the whole target for the global variant, a circular window in shared memory otherwise
    constexpr int CM = SMEM ? DP_SCOLS - 1 : 0x7fffffff;
    int8_t *u = SMEM ? S.s_state : S.u, *v = SMEM ? S.s_state + DP_SCOLS : S.v, *x = SMEM ? S.s_state + 2 * DP_SCOLS : S.x;
    int8_t *y = SMEM ? S.s_state + 3 * DP_SCOLS : S.y, *x2 = SMEM ? S.s_state + 4 * DP_SCOLS : S.x2, *y2 = SMEM ? S.s_state + 5 * DP_SCOLS : S.y2;
    int32_t *H = SMEM ? S.s_H : S.H;
    if (!SMEM) {
        for (int t = lane; t < tlen; t += 32) {
            u[t] = v[t] = x[t] = y[t] = (int8_t)(-q - e);
            x2[t] = y2[t] = (int8_t)(-q2 - e2);
        }
    }
    __syncwarp();
    // uniform running state (identical in every lane)
    int pst = -1, pen = -1;
    int32_t ez_max = 0, ez_max_t = -1, ez_max_q = -1, ez_mqe = KSW_NEG_INF, ez_mqe_t = -1, ez_mte = KSW_NEG_INF, ez_mte_q = -1;
    int32_t ez_score = KSW_NEG_INF, zdropped = 0;
    int32_t H0 = 0, last_H0_t = 0;
    unsigned long long cells = 0;
    const int nr = qlen + tlen - 1;
    const bool bounded = (flag & KSW_EXTZ_ONLY) && !approx && T.end_bonus <= 0;
    for (int r = 0; r < nr; ++r) {
        if (bounded && ext_bound_stop(o.a, q, e, q2, e2, qlen, tlen, r, ez_max)) { zdropped = 1; break; }
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
        if (en > (r + w) >> 1) en = (r + w) >> 1;
        if (st > en) { zdropped = 1; break; }
        cells += (unsigned long long)(en - st + 1);
        const int bnd = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
        int sx1, sx21, sv1;        // left neighbour of the first cell
        if (st > 0) {
            if (st - 1 >= pst && st - 1 <= pen) sx1 = x[(st - 1) & CM], sx21 = x2[(st - 1) & CM], sv1 = v[(st - 1) & CM];
            else sx1 = -q - e, sx21 = -q2 - e2, sv1 = -q - e;
        } else sx1 = -q - e, sx21 = -q2 - e2, sv1 = bnd;
        if (lane == 0) {
            if (SMEM && en > pen) {     // a column enters the band: its slot still holds column en - DP_SCOLS
                const int se = en & CM;
                u[se] = v[se] = x[se] = y[se] = (int8_t)(-q - e);
                x2[se] = y2[se] = (int8_t)(-q2 - e2);
            }
            if (en == r) { y[r & CM] = (int8_t)(-q - e); y2[r & CM] = (int8_t)(-q2 - e2); u[r & CM] = (int8_t)bnd; }
        }
        __syncwarp();
        uint8_t *pr = p + (int64_t)r * ncol;
        // DPU chunks of 32 cells per lane per iteration (independent dependency chains); groups run from high t to
        // low t and every group loads all its inputs before storing, so in-place updates never overtake a reader
        const int nchunks = ((en - st) >> 5) + 1;
        for (int cb = ((nchunks - 1) / DPU) * DPU; cb >= 0; cb -= DPU) {
            int ut[DPU], yt[DPU], y2t[DPU], v1[DPU], x1[DPU], x21[DPU], zz[DPU], tt[DPU];
#pragma unroll
            for (int k = 0; k < DPU; ++k) {
                const int t = st + ((cb + k) << 5) + lane;
                tt[k] = (cb + k < nchunks && t <= en) ? t : -1;
                ut[k] = yt[k] = y2t[k] = zz[k] = 0; v1[k] = sv1, x1[k] = sx1, x21[k] = sx21;
                if (tt[k] >= 0) {
                    const int ts = t & CM, tp = (t - 1) & CM;
                    ut[k] = u[ts], yt[k] = y[ts], y2t[k] = y2[ts];
                    if (t > st) v1[k] = v[tp], x1[k] = x[tp], x21[k] = x2[tp];
                    int qc = dp_base(T.q, T.qstep, T.qcomp, r - t), tc = dp_base(T.t, T.tstep, 0, t);
                    zz[k] = (qc > 3 || tc > 3) ? -o.sc_ambi : qc == tc ? o.a : -o.b;
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < DPU; ++k) {
                if (tt[k] < 0) continue;
                const int t = tt[k], ts = t & CM;
                int z = zz[k];
                int a = x1[k] + v1[k], b = yt[k] + ut[k], a2 = x21[k] + v1[k], b2 = y2t[k] + ut[k], d;
                if (!right) {
                    d = a > z ? 1 : 0;  z = z > a ? z : a;
                    d = b > z ? 2 : d;  z = z > b ? z : b;
                    d = a2 > z ? 3 : d; z = z > a2 ? z : a2;
                    d = b2 > z ? 4 : d; z = z > b2 ? z : b2;
                } else {
                    d = z > a ? 0 : 1;  z = z > a ? z : a;
                    d = z > b ? d : 2;  z = z > b ? z : b;
                    d = z > a2 ? d : 3; z = z > a2 ? z : a2;
                    d = z > b2 ? d : 4; z = z > b2 ? z : b2;
                }
                if (z > o.a) z = o.a;
                u[ts] = (int8_t)(z - v1[k]);
                v[ts] = (int8_t)(z - ut[k]);
                int tmp = z - q;  a -= tmp, b -= tmp;
                tmp = z - q2;     a2 -= tmp, b2 -= tmp;
                if (!right) {
                    x[ts] = (int8_t)((a > 0 ? a : 0) - qe);     d |= a > 0 ? 0x08 : 0;
                    y[ts] = (int8_t)((b > 0 ? b : 0) - qe);     d |= b > 0 ? 0x10 : 0;
                    x2[ts] = (int8_t)((a2 > 0 ? a2 : 0) - qe2); d |= a2 > 0 ? 0x20 : 0;
                    y2[ts] = (int8_t)((b2 > 0 ? b2 : 0) - qe2); d |= b2 > 0 ? 0x40 : 0;
                } else {
                    x[ts] = (int8_t)((a >= 0 ? a : 0) - qe);     d |= a >= 0 ? 0x08 : 0;
                    y[ts] = (int8_t)((b >= 0 ? b : 0) - qe);     d |= b >= 0 ? 0x10 : 0;
                    x2[ts] = (int8_t)((a2 >= 0 ? a2 : 0) - qe2); d |= a2 >= 0 ? 0x20 : 0;
                    y2[ts] = (int8_t)((b2 >= 0 ? b2 : 0) - qe2); d |= b2 >= 0 ? 0x40 : 0;
                }
                pr[t - st] = (uint8_t)d;
            }
            __syncwarp();
        }
        if (!approx) {
            int32_t max_H, max_t, Hen, Hst;
            if (r > 0) {
                Hen = en > 0 ? H[(en - 1) & CM] + u[en & CM] : H[en & CM] + v[en & CM];     // uniform load, before H[en-1] is advanced
                __syncwarp();
                const int en1 = st + (en - st) / 4 * 4;
                int32_t bh = KSW_NEG_INF * 2; int brank = 0x7fffffff, bt = -1, hst = 0;
                for (int t = st + lane; t < en; t += 32) {
                    int32_t h = H[t & CM] + v[t & CM];
                    H[t & CM] = h;
                    if (t == st) hst = h;
                    int rank = t < en1 ? 1 + (((t - st) & 3) << 20) + ((t - st) >> 2) : 1 + (4 << 20) + (t - en1);
                    if (h > bh || (h == bh && rank < brank)) bh = h, brank = rank, bt = t;
                }
                if (lane == 0) { H[en & CM] = Hen; if (Hen > bh || (Hen == bh)) bh = Hen, brank = 0, bt = en; }
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    int32_t oh = __shfl_xor_sync(FULL, bh, d); int orank = __shfl_xor_sync(FULL, brank, d), ot = __shfl_xor_sync(FULL, bt, d);
                    if (oh > bh || (oh == bh && orank < brank)) bh = oh, brank = orank, bt = ot;
                }
                max_H = bh, max_t = bt;
                Hst = st == en ? Hen : __shfl_sync(FULL, hst, 0);
                __syncwarp();
            } else {
                Hen = Hst = (int32_t)v[0] - qe;
                if (lane == 0) H[0] = Hen;
                max_H = Hen, max_t = 0;
                __syncwarp();
            }
            if (en == tlen - 1 && Hen > ez_mte) ez_mte = Hen, ez_mte_q = r - en;
            if (r - st == qlen - 1 && Hst > ez_mqe) ez_mqe = Hst, ez_mqe_t = st;
            // ksw_apply_zdrop
            bool stop = false;
            if (max_H > ez_max) ez_max = max_H, ez_max_t = max_t, ez_max_q = r - max_t;
            else if (max_t >= ez_max_t && r - max_t >= ez_max_q) {
                int tl = max_t - ez_max_t, ql = (r - max_t) - ez_max_q, l = tl > ql ? tl - ql : ql - tl;
                if (T.zdrop >= 0 && ez_max - max_H > T.zdrop + l * e2) zdropped = 1, stop = true;
            }
            if (stop) break;
            if (r == nr - 1 && en == tlen - 1) ez_score = Hen;
        } else {
            if (r > 0) {
                if (last_H0_t >= st && last_H0_t <= en && last_H0_t + 1 >= st && last_H0_t + 1 <= en) {
                    int d0 = v[last_H0_t & CM], d1 = u[(last_H0_t + 1) & CM];
                    if (d0 > d1) H0 += d0; else H0 += d1, ++last_H0_t;
                } else if (last_H0_t >= st && last_H0_t <= en) H0 += v[last_H0_t & CM];
                else ++last_H0_t, H0 += u[last_H0_t & CM];
            } else H0 = (int32_t)v[0] - qe, last_H0_t = 0;
            if (r == nr - 1 && en == tlen - 1) ez_score = H0;
        }
        pst = st, pen = en;
    }
    if (lane == 0) {
        atomicAdd(cells_acc, cells);
        R.max = ez_max; R.max_t = ez_max_t; R.max_q = ez_max_q; R.mqe = ez_mqe; R.mqe_t = ez_mqe_t;
        R.mte = ez_mte; R.mte_q = ez_mte_q; R.score = ez_score; R.zdropped = zdropped;
    }
    __syncwarp();
}

// memory a forward pass needs for its direction bytes (row r of the anti-diagonal sweep at r * ncol)
__host__ __device__ __forceinline__ int64_t dp_dir_bytes(int qlen, int tlen, int w_in)
{
    if (qlen <= 0 || tlen <= 0) return 0;
    int w = w_in < 0 ? (tlen > qlen ? tlen : qlen) : w_in;
    int ncol = qlen < tlen ? qlen : tlen;
    if (ncol > w + 1) ncol = w + 1;
    return (int64_t)(qlen + tlen - 1) * ncol;
}

// ksw_backtrack over the direction bytes written by warp_extd2; sequential, one thread
__device__ void extd2_traceback(const DpTask &T, DpRes &R, const uint8_t *p, uint32_t *ezcig, int ezcap, int32_t *err)
{
    const int qlen = T.qlen, tlen = T.tlen, flag = T.flag;
    R.n_cigar = 0; R.cigar = ezcig; R.reach_end = 0;
    if (qlen <= 0 || tlen <= 0) return;
    const int w = T.w < 0 ? (tlen > qlen ? tlen : qlen) : T.w;
    int ncol = qlen < tlen ? qlen : tlen;
    if (ncol > w + 1) ncol = w + 1;
    int i0 = -1, j0 = -1;
    if (!R.zdropped && !(flag & KSW_EXTZ_ONLY)) i0 = tlen - 1, j0 = qlen - 1;
    else if (!R.zdropped && (flag & KSW_EXTZ_ONLY) && R.mqe + T.end_bonus > R.max) R.reach_end = 1, i0 = R.mqe_t, j0 = qlen - 1;
    else if (R.max_t >= 0 && R.max_q >= 0) i0 = R.max_t, j0 = R.max_q;
    if (i0 < 0 || j0 < 0) return;
    EzPush ep; ep.c = ezcig; ep.n = 0; ep.cap = ezcap;
    int i = i0, j = j0, state = 0;
    while (i >= 0 && j >= 0) {
        int r = i + j, force_state = -1;
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
        if (en > (r + w) >> 1) en = (r + w) >> 1;
        if (i < st) force_state = 2;
        if (i > en) force_state = 1;
        uint32_t tmp = force_state < 0 ? p[(int64_t)r * ncol + (i - st)] : 0;
        if (state == 0) state = tmp & 7;
        else if (!(tmp >> (state + 2) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (force_state >= 0) state = force_state;
        if (state == 0) ep.push(0, 1), --i, --j;
        else if (state == 1 || state == 3) ep.push(2, 1), --i;
        else ep.push(1, 1), --j;
    }
    if (i >= 0) ep.push(2, i + 1);
    if (j >= 0) ep.push(1, j + 1);
    if (ep.n > ep.cap) { atomicOr(err, TELR_ERR_CIGCAP); ep.n = 0; }
    if (!(flag & KSW_REV_CIGAR))
        for (int k = 0; k < ep.n >> 1; ++k) { uint32_t t = ep.c[k]; ep.c[k] = ep.c[ep.n - 1 - k]; ep.c[ep.n - 1 - k] = t; }
    R.n_cigar = ep.n;
}

// returns 1 when the vectorised path produced the direction bytes (sign-bit format, vec_stride rows), 0 for the scalar path
__device__ __forceinline__ int warp_extd2(const Opt &o, const DpTask &T, DpRes &R, DpScratch &S, unsigned long long *cells_acc, int32_t *err)
{
    if (S.vsm && vec_ok(o, T.qlen, T.tlen, T.w) && vec_dir_bytes(T.qlen, T.tlen, T.w) <= S.dir_cap) {
        if (warp_extd2_vec(o, T, R, *S.vsm, S.stab, S.dir, cells_acc)) return 1;
    }
    warp_extd2_impl<false>(o, T, R, S, cells_acc, err);      // ambiguous bases or a band wider than the window: state arrays in global memory
    return 0;
}

// local affine Smith-Waterman score probe with end coordinates (first maximum in target-major order)
__device__ void warp_ll(const Opt &o, const DpTask &T, DpRes &R, DpScratch &S, unsigned long long *cells_acc)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int qlen = T.qlen, tlen = T.tlen, gapoe = o.q + o.e, gape = o.e;
    int32_t *Ha = S.ll, *Hb = Ha + tlen, *Hc = Hb + tlen, *Ea = Hc + tlen, *Eb = Ea + tlen, *F = Eb + tlen;
    for (int i = lane; i < tlen; i += 32) Ha[i] = Hb[i] = Hc[i] = Ea[i] = Eb[i] = F[i] = 0;
    __syncwarp();
    int32_t bh = 0; int bi = 0x7fffffff, bj = 0x7fffffff;
    for (int d = 0; d < qlen + tlen - 1; ++d) {
        const int ist = d - qlen + 1 > 0 ? d - qlen + 1 : 0, ien = d < tlen - 1 ? d : tlen - 1;
        int32_t *Hcur = Ha, *Hp2 = Hb; const int32_t *Eprev = Ea; int32_t *Ecur = Eb;
        // diagonal d lives in buffer d % 3, so d-2 is in buffer (d+1) % 3
        switch (d % 3) { case 0: Hcur = Ha; Hp2 = Hb; break; case 1: Hcur = Hb; Hp2 = Hc; break; default: Hcur = Hc; Hp2 = Ha; break; }
        if (d & 1) Eprev = Eb, Ecur = Ea;
        for (int i = ist + lane; i <= ien; i += 32) {
            const int j = d - i;
            int qc = dp_base(T.q, T.qstep, T.qcomp, j), tc = dp_base(T.t, T.tstep, 0, i);
            int32_t s = (qc > 3 || tc > 3) ? -o.sc_ambi : qc == tc ? o.a : -o.b;
            int32_t hd = (i > 0 && j > 0) ? Hp2[i - 1] : 0;
            int32_t ee = i > 0 ? Eprev[i - 1] : 0, ff = j > 0 ? F[i] : 0;
            int32_t h = hd + s;
            if (h < ee) h = ee;
            if (h < ff) h = ff;
            if (h < 0) h = 0;
            Hcur[i] = h;
            ee -= gape; if (ee < h - gapoe) ee = h - gapoe; if (ee < 0) ee = 0;
            ff -= gape; if (ff < h - gapoe) ff = h - gapoe; if (ff < 0) ff = 0;
            Ecur[i] = ee; F[i] = ff;
            if (h > bh || (h == bh && h > 0 && (i < bi || (i == bi && j < bj)))) bh = h, bi = i, bj = j;
        }
        __syncwarp();
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        int32_t oh = __shfl_xor_sync(FULL, bh, d); int oi = __shfl_xor_sync(FULL, bi, d), oj = __shfl_xor_sync(FULL, bj, d);
        if (oh > bh || (oh == bh && oh > 0 && (oi < bi || (oi == bi && oj < bj)))) bh = oh, bi = oi, bj = oj;
    }
    if (lane == 0) {
        atomicAdd(cells_acc, (unsigned long long)qlen * (unsigned long long)tlen);
        R.ll_score = bh; R.ll_qe = bh > 0 ? bj : -1; R.ll_te = bh > 0 ? bi : -1;
    }
    __syncwarp();
}

__global__ void k_unpack_reads(int n_reads, const uint32_t *seq2, const uint32_t *nmask, const int64_t *read_off, const int32_t *read_len,
                               const int64_t *rbyte_off, uint8_t *out)
{
    for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const int n = read_len[r];
        const int64_t off = read_off[r];
        uint8_t *fw = out + rbyte_off[r], *rc = fw + n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            int64_t pp = off + i;
            int c = (seq2[pp >> 4] >> (2 * (pp & 15))) & 3;
            if ((nmask[pp >> 5] >> (pp & 31)) & 1) c = 4;
            fw[i] = (uint8_t)c;
            rc[n - 1 - i] = (uint8_t)(c < 4 ? 3 - c : 4);
        }
    }
}

__global__ void __launch_bounds__(128) k_al_init(const __grid_constant__ AlignArgs A)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= A.n_work) return;
    AlWork &W = A.work[w];
    const int pidx = A.work_list[w];
    const int read = A.prob_read[pidx], ls = A.prob_ls[pidx], l = ls >> 1, strand = ls & 1;
    const int qlen = A.read_len[read], L = A.contig_len[l];
    const int n_a = (int)(A.prob_aoff[pidx + 1] - A.prob_aoff[pidx]);
    W.pidx = pidx; W.done = 0; W.pending = 0; W.has_task = 0; W.dir_off = 0;
    AlnCtx &c = A.actx[w];
    c.o = A.d_opt;
    c.tseq = A.ctg_bytes + A.ctg_boff[l] + (strand ? L : 0); c.tlen = L;
    c.qseq[0] = A.read_bytes + A.rbyte_off[read]; c.qseq[1] = c.qseq[0] + qlen; c.qlen = qlen;
    c.a = A.anchors + A.prob_aoff[pidx]; c.n_a = A.prob_nca[pidx];
    c.regs = A.regs + A.prob_roff[pidx]; c.n_regs = A.prob_nregs[pidx];
    c.cap_regs = (int)(A.prob_roff[pidx + 1] - A.prob_roff[pidx]);
    c.cig = A.cigs + W.cig_off; c.cig_top = 0; c.cig_cap = (uint32_t)W.cig_cap;
    {   // the chaining scratch of this problem is free now: reuse its region-bookkeeping part and the candidate list
        uint8_t *b = A.prob_scratch + A.prob_soff[pidx];
        ChainScratch cs;
        chain_scratch_carve(cs, b, (size_t)n_a + 1);
        hit_scratch_carve(c.hs, b + chain_scratch_bytes((size_t)n_a + 1), (size_t)(2 * (n_a / 3) + 8));
        c.K = cs.ord; c.capK = n_a;
    }
    c.err = 0; c.n_tasks = 0; c.defer_finish = 1; c.rep_len = A.prob_replen ? A.prob_replen[pidx] : 0;
    c.phase = PH_START;
    res_reset(A.res[w]);
    A.res[w].cigar = A.cigs + W.ez_off;
}


// coroutine + DP + traceback for one problem per warp (persistent warps, dynamic queue, no global barriers).
// The coroutine state lives in shared memory while the warp owns the problem and is written back for k_al_finish.
struct AlWarpSmem { AlnCtx c; DpTask task; DpRes res; int more, role; };

__global__ void __launch_bounds__(AL_THREADS, AL_BLOCKS_PER_SM) k_al_fused(const __grid_constant__ AlignArgs A)
{
    __shared__ AlWarpSmem WS[AL_WARPS];
    __shared__ uint2 stab[256];
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    VecSmem *DS = reinterpret_cast<VecSmem *>(dyn_smem);
    const Opt &o = A.o;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    AlWarpSmem &W = WS[wid];
    uint8_t *base = A.warp_scratch + (size_t)(blockIdx.x * AL_WARPS + wid) * A.warp_scratch_stride;
    const size_t maxT = ((size_t)A.max_tlen + 64) & ~(size_t)15;
    DpScratch S;
    S.u = (int8_t *)base; base += maxT; S.v = (int8_t *)base; base += maxT; S.x = (int8_t *)base; base += maxT;
    S.y = (int8_t *)base; base += maxT; S.x2 = (int8_t *)base; base += maxT; S.y2 = (int8_t *)base; base += maxT;
    S.H = (int32_t *)base; base += maxT * 4;
    S.ll = (int32_t *)base; base += maxT * 4 * 6;
    S.bnd = (uint32_t *)base; base += (((size_t)A.max_qlen + 64) & ~(size_t)15) * 6;
    base = (uint8_t *)(((uintptr_t)base + 255) & ~(uintptr_t)255);
    uint8_t *own_dir = base;
    S.s_state = nullptr; S.s_H = nullptr; S.vsm = A.use_vec ? &DS[wid] : nullptr; S.stab = stab;
    vec_fill_stab(stab, o);
    for (;;) {
        int wi = 0;
        if (lane == 0) wi = (int)atomicAdd(&A.rc[2], 1ULL);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= A.n_work) break;
        {   // coroutine state: global -> shared
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&A.actx[wi]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&W.c);
            for (int i = lane; i < (int)(sizeof(AlnCtx) / 4); i += 32) dst[i] = src[i];
        }
        if (lane == 0) { res_reset(W.res); W.res.cigar = A.cigs + A.work[wi].ez_off; }
        S.ezcig = A.cigs + A.work[wi].ez_off; S.ezcap = A.work[wi].ez_cap;
        __syncwarp();
        for (;;) {
            long long t0 = 0;
            if (lane == 0) {
                if (AL_CENSUS(A)) t0 = clock64();
                W.more = aln_next(W.c, W.res, W.task) ? 1 : 0;
                if (AL_CENSUS(A)) atomicAdd(&A.rc[4], (unsigned long long)(clock64() - t0));
            }
            __syncwarp();
            if (!W.more) break;
            if (AL_CENSUS(A)) t0 = clock64();
            if (W.task.kind == 0) {
                const bool fast = A.use_fast && fill_fast_ok(W.task);
                const int64_t need = fast ? (int64_t)W.task.qlen * fill_stride(W.task.tlen) : vec_dir_bytes(W.task.qlen, W.task.tlen, W.task.w);
                int big_slot = -1;
                S.dir = own_dir; S.dir_cap = A.dir_cap;
                if (need > A.dir_cap && need <= A.big_cap && A.n_big > 0) {
                    if (lane == 0) {        // take one of the shared large traceback buffers (holders never wait on anything)
                        unsigned ns = 64;
                        for (int sl = (blockIdx.x * AL_WARPS + wid) % A.n_big;; sl = (sl + 1) % A.n_big) {
                            if (atomicCAS(&A.big_lock[sl], 0, 1) == 0) { big_slot = sl; break; }
                            __nanosleep(ns);
                            if (ns < 4096) ns <<= 1;
                        }
                        __threadfence();
                    }
                    big_slot = __shfl_sync(FULL, big_slot, 0);
                    S.dir = A.big + (int64_t)big_slot * A.big_cap; S.dir_cap = A.big_cap;
                }
                bool done_fast = false;
                if (fast && need <= S.dir_cap) done_fast = warp_fill_fast(o, W.task, W.res, S.dir, S.bnd, A.stat_cells);
                int vec = 0;
                if (!done_fast) vec = warp_extd2(o, W.task, W.res, S, A.stat_cells, A.err);
                __syncwarp();
                const int path = done_fast ? 0 : vec ? 1 : 2;
                if (AL_CENSUS(A) && lane == 0) {
                    long long t1 = clock64();
                    atomicAdd(&A.rc[5 + path], (unsigned long long)(t1 - t0));
                    atomicAdd(&A.rc[12 + path], 1ULL);
                    atomicAdd(&A.rc[16 + path], (unsigned long long)W.task.qlen * (unsigned long long)W.task.tlen);
                    if (path == 0) {        // gap fills by target length: one 256-column pass, or two
                        const int tl = W.task.tlen;
                        const int b = tl <= 256 ? 0 : tl <= 288 ? 1 : tl <= 320 ? 2 : tl <= 384 ? 3 : tl <= 512 ? 4 : 5;
                        atomicAdd(&A.rc[64 + b], 1ULL);
                        atomicAdd(&A.rc[70 + b], (unsigned long long)(t1 - t0));
                        atomicAdd(&A.rc[76 + b], (unsigned long long)W.task.qlen * (unsigned long long)W.task.tlen);
                    }
                    if (path == 1) {        // shape census of the general DP: buckets of min(qlen, tlen)
                        const int mn = W.task.qlen < W.task.tlen ? W.task.qlen : W.task.tlen;
                        const int b = mn < 64 ? 0 : mn < 128 ? 1 : mn < 256 ? 2 : mn < 512 ? 3 : mn < 1024 ? 4 : 5;
                        atomicAdd(&A.rc[32 + b], (unsigned long long)(t1 - t0));
                        atomicAdd(&A.rc[38 + b], 1ULL);
                        if (W.res.zdropped) atomicAdd(&A.rc[44 + b], 1ULL);
                        atomicAdd(&A.rc[50 + b], (unsigned long long)(W.res.max_t + W.res.max_q + 2));
                        atomicAdd(&A.rc[56 + b], (unsigned long long)(W.task.qlen + W.task.tlen));
                    }
                    t0 = t1;
                }
                if (done_fast) fill_traceback(W.task, W.res, S.dir, S.ezcig, S.ezcap, A.err, *reinterpret_cast<TbSmem *>(DS[wid].H));     // the DP state window is idle during traceback
                if (lane == 0) {
                    if (!done_fast) { if (vec) extd2_traceback_vec(W.task, W.res, S.dir, S.ezcig, S.ezcap, A.err); else extd2_traceback(W.task, W.res, S.dir, S.ezcig, S.ezcap, A.err); }
                    if (big_slot >= 0) { __threadfence(); atomicExch(&A.big_lock[big_slot], 0); }
                    if (AL_CENSUS(A)) atomicAdd(&A.rc[8 + path], (unsigned long long)(clock64() - t0));
                }
            } else {
                warp_ll(o, W.task, W.res, S, A.stat_cells);
                if (AL_CENSUS(A) && lane == 0) { atomicAdd(&A.rc[11], (unsigned long long)(clock64() - t0)); atomicAdd(&A.rc[15], 1ULL); }
            }
            __syncwarp();
        }
        {   // coroutine state: shared -> global (k_al_finish continues from PH_FINISH)
            uint32_t *dst = reinterpret_cast<uint32_t *>(&A.actx[wi]);
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&W.c);
            for (int i = lane; i < (int)(sizeof(AlnCtx) / 4); i += 32) dst[i] = src[i];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Task-type-specialised form of the same kernel.  k_al_fused lets every warp run whatever its read needs next, so the 16 warps
// of an SM sit in different loops and the hot code (fill forward 14 KB + fill traceback 7 KB + windowed DP 13 KB) sits at the
// 32 KB instruction-cache capacity; every specialised loop body added to it made the whole kernel slower (profiles/README.md).
// Here an SM has a ROLE -- gap fills, or everything else (extensions, exact re-runs, local probes) -- and the problems move
// between the roles through two device-wide rings: a warp executes the pending DP request of a problem, advances the
// problem's coroutine (its state lives in global memory between visits), keeps the problem while the next request belongs to
// its own role, and otherwise publishes it on the other role's ring.  A problem is in at most one ring at a time, so a ring of
// n_work + 1 slots can never overflow.  Workers whose ring is empty start fresh problems (longest read first), then help the
// other role; the kernel ends when every problem has finished.  No global barrier anywhere.
__device__ __forceinline__ int aq_role_of(const AlignArgs &A, const DpTask &T)
{
    if (!(T.kind == 0 && A.use_fast && fill_fast_ok(T))) return 1;
#if TELR_FILL_FC12
    if (A.wide_per8 > 0 && fill_width(T.tlen) == 12) return 2;         // 257..384 target columns: one pass of the 12-column instance, on its own SMs
#endif
    return 0;
}

// lane 0 only.  The caller has written the problem's state to global memory (all lanes) and synchronised the warp.
__device__ __forceinline__ void aq_push(const AlignArgs &A, int role, int wi)
{
    __threadfence();                                                     // release: state before the ring entry
    const unsigned t = (unsigned)atomicAdd(&A.q_state[AQ_TAIL0 + role], 1);
    atomicExch(&A.q_ring[role][t % (unsigned)A.q_cap], wi + 1);
    atomicAdd(&A.q_state[AQ_AVAIL0 + role], 1);
}
// lane 0 only; -1 when the ring is empty
__device__ __forceinline__ int aq_pop(const AlignArgs &A, int role)
{
    if (*(volatile int32_t *)&A.q_state[AQ_AVAIL0 + role] <= 0) return -1;
    if (atomicSub(&A.q_state[AQ_AVAIL0 + role], 1) <= 0) { atomicAdd(&A.q_state[AQ_AVAIL0 + role], 1); return -1; }
    const unsigned h = (unsigned)atomicAdd(&A.q_state[AQ_HEAD0 + role], 1);
    int32_t *slot = &A.q_ring[role][h % (unsigned)A.q_cap];
    int v;
    while ((v = atomicExch(slot, 0)) == 0) __nanosleep(32);               // its pusher holds the ticket and is about to write
    return v - 1;
}

__device__ __forceinline__ void aq_exec(const AlignArgs &A, const Opt &o, AlWarpSmem &W, DpScratch &S, uint8_t *own_dir, VecSmem &vs)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    if (W.task.kind == 0) {
        const bool fast = A.use_fast && fill_fast_ok(W.task);
        const int64_t need = fast ? (int64_t)W.task.qlen * fill_stride(W.task.tlen) : vec_dir_bytes(W.task.qlen, W.task.tlen, W.task.w);
        int big_slot = -1;
        S.dir = own_dir; S.dir_cap = A.dir_cap;
        if (need > A.dir_cap && need <= A.big_cap && A.n_big > 0) {
            if (lane == 0) {        // take one of the shared large traceback buffers (holders never wait on anything)
                unsigned ns = 64;
                for (int sl = (blockIdx.x * AL_WARPS + (threadIdx.x >> 5)) % A.n_big;; sl = (sl + 1) % A.n_big) {
                    if (atomicCAS(&A.big_lock[sl], 0, 1) == 0) { big_slot = sl; break; }
                    __nanosleep(ns);
                    if (ns < 4096) ns <<= 1;
                }
                __threadfence();
            }
            big_slot = __shfl_sync(FULL, big_slot, 0);
            S.dir = A.big + (int64_t)big_slot * A.big_cap; S.dir_cap = A.big_cap;
        }
        bool done_fast = false;
        if (fast && need <= S.dir_cap) done_fast = warp_fill_fast(o, W.task, W.res, S.dir, S.bnd, A.stat_cells, A.wide_per8 > 0);
        int vec = 0;
        if (!done_fast) vec = warp_extd2(o, W.task, W.res, S, A.stat_cells, A.err);
        __syncwarp();
        if (done_fast) fill_traceback(W.task, W.res, S.dir, S.ezcig, S.ezcap, A.err, *reinterpret_cast<TbSmem *>(vs.H));
        if (lane == 0) {
            if (!done_fast) { if (vec) extd2_traceback_vec(W.task, W.res, S.dir, S.ezcig, S.ezcap, A.err); else extd2_traceback(W.task, W.res, S.dir, S.ezcig, S.ezcap, A.err); }
            if (big_slot >= 0) { __threadfence(); atomicExch(&A.big_lock[big_slot], 0); }
        }
    } else {
        warp_ll(o, W.task, W.res, S, A.stat_cells);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(AL_THREADS, AL_BLOCKS_PER_SM) k_al_queue(const __grid_constant__ AlignArgs A)
{
    __shared__ AlWarpSmem WS[AL_WARPS];
    __shared__ uint2 stab[256];
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    VecSmem *DS = reinterpret_cast<VecSmem *>(dyn_smem);
    const Opt &o = A.o;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    AlWarpSmem &W = WS[wid];
    uint8_t *base = A.warp_scratch + (size_t)(blockIdx.x * AL_WARPS + wid) * A.warp_scratch_stride;
    const size_t maxT = ((size_t)A.max_tlen + 64) & ~(size_t)15;
    DpScratch S;
    S.u = (int8_t *)base; base += maxT; S.v = (int8_t *)base; base += maxT; S.x = (int8_t *)base; base += maxT;
    S.y = (int8_t *)base; base += maxT; S.x2 = (int8_t *)base; base += maxT; S.y2 = (int8_t *)base; base += maxT;
    S.H = (int32_t *)base; base += maxT * 4;
    S.ll = (int32_t *)base; base += maxT * 4 * 6;
    S.bnd = (uint32_t *)base; base += (((size_t)A.max_qlen + 64) & ~(size_t)15) * 6;
    base = (uint8_t *)(((uintptr_t)base + 255) & ~(uintptr_t)255);
    uint8_t *own_dir = base;
    S.s_state = nullptr; S.s_H = nullptr; S.vsm = A.use_vec ? &DS[wid] : nullptr; S.stab = stab;
    vec_fill_stab(stab, o);
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    // The role belongs to the SM (all its warps run one loop); it starts from a fixed interleave and MOVES when the SM's ring
    // runs dry: a warp that finds neither work for its role nor a fresh problem re-assigns its SM to the role with the longest
    // backlog, so any mix of fills and extensions (presets, coverage, read lengths) balances itself without two loops sharing
    // an instruction cache for longer than the hand-over.
    int32_t *sm_role = A.q_state + AQ_SLOTS + (smid & (AQ_MAX_SM - 1));
    if (threadIdx.x == 0)
        atomicCAS(sm_role, 0, 1 + ((int)(smid & 7u) < A.ext_per8 ? 1 : (int)(smid & 7u) < A.ext_per8 + A.wide_per8 ? 2 : 0));
    __syncthreads();
    unsigned idle_ns = 64;
    for (;;) {
        // ---- take a problem: the SM's ring, then a fresh problem, then move the SM to the longest ring ----
        int wi = -1, fresh = 0;
        if (lane == 0) {
            int role = *(volatile int32_t *)sm_role - 1;
            wi = aq_pop(A, role);
            if (wi < 0 && *(volatile int32_t *)&A.q_state[AQ_START] < A.n_work) {
                const int t = atomicAdd(&A.q_state[AQ_START], 1);
                if (t < A.n_work) { wi = t; fresh = 1; }
            }
            if (wi < 0) {
                int best = -1, best_n = 0;
                for (int r = 0; r < AQ_ROLES; ++r) {
                    const int n = *(volatile int32_t *)&A.q_state[AQ_AVAIL0 + r];
                    if (r != role && n > best_n) best = r, best_n = n;
                }
                if (best >= 0) {
                    wi = aq_pop(A, best);
                    if (wi >= 0) { atomicExch(sm_role, best + 1); atomicAdd(&A.q_state[AQ_STEAL], 1); role = best; }
                }
            }
            if (wi < 0 && *(volatile int32_t *)&A.q_state[AQ_DONE] >= A.n_work) wi = -2;
            if (wi >= 0 && !fresh) __threadfence();                     // acquire: the previous owner's writes
            W.role = role;
        }
        wi = __shfl_sync(FULL, wi, 0); fresh = __shfl_sync(FULL, fresh, 0);
        if (wi == -2) break;
        if (wi < 0) { __nanosleep(idle_ns); if (idle_ns < 2048) idle_ns <<= 1; continue; }
        idle_ns = 64;
        {   // coroutine state (and, for a queued problem, its pending request): global -> shared
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&A.actx[wi]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&W.c);
            for (int i = lane; i < (int)(sizeof(AlnCtx) / 4); i += 32) dst[i] = __ldcg(src + i);
            if (!fresh) {
                const uint32_t *ts = reinterpret_cast<const uint32_t *>(&A.tasks[wi]);
                uint32_t *td = reinterpret_cast<uint32_t *>(&W.task);
                for (int i = lane; i < (int)(sizeof(DpTask) / 4); i += 32) td[i] = __ldcg(ts + i);
            }
        }
        if (lane == 0) { res_reset(W.res); W.res.cigar = A.cigs + A.work[wi].ez_off; }
        S.ezcig = A.cigs + A.work[wi].ez_off; S.ezcap = A.work[wi].ez_cap;
        __syncwarp();
        int have_task = !fresh;
        for (;;) {
            if (have_task) aq_exec(A, o, W, S, own_dir, DS[wid]);
            if (lane == 0) {
                W.more = aln_next(W.c, W.res, W.task) ? 1 : 0;
                if (W.more) W.more = 1 + aq_role_of(A, W.task);
            }
            __syncwarp();
            const int more = W.more;
            if (more && more - 1 == W.role) { have_task = 1; continue; }
            {   // the problem leaves this warp: coroutine state (and the request it waits for) shared -> global
                uint32_t *dst = reinterpret_cast<uint32_t *>(&A.actx[wi]);
                const uint32_t *src = reinterpret_cast<const uint32_t *>(&W.c);
                for (int i = lane; i < (int)(sizeof(AlnCtx) / 4); i += 32) dst[i] = src[i];
                if (more) {
                    uint32_t *td = reinterpret_cast<uint32_t *>(&A.tasks[wi]);
                    const uint32_t *ts = reinterpret_cast<const uint32_t *>(&W.task);
                    for (int i = lane; i < (int)(sizeof(DpTask) / 4); i += 32) td[i] = ts[i];
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (more) aq_push(A, more - 1, wi);
                else { __threadfence(); atomicAdd(&A.q_state[AQ_DONE], 1); }
            }
            __syncwarp();
            break;
        }
    }
}

// Deferred mm_update_extra statistics (matches, block length, ambiguous bases, best local score dp_max), one warp per
// problem.  Lane 0 runs the sequential CIGAR clean-up (fix_cigar); the per-base pass is then spread over the lanes one
// CIGAR op each.  The score recurrence s <- max(0, s + x) with its running maximum is a monoid over ops:
//   f(s) = max(s + A, B),   best prefix value given s = max(s + MA, MB)
// combined left-to-right by an ordered warp reduction, so the result equals the sequential scan exactly.
struct RfMono { int32_t A, B, MA, MB; };
__device__ __forceinline__ RfMono rf_combine(const RfMono &l, const RfMono &r)
{
    RfMono o;
    o.A = l.A + r.A;
    o.B = max(l.B + r.A, r.B);
    o.MA = max(l.MA, l.A + r.MA);
    o.MB = max(max(l.MB, l.B + r.MA), r.MB);
    return o;
}

__global__ void __launch_bounds__(256) k_al_regfin(const __grid_constant__ AlignArgs A)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int NEG = -(1 << 28);
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < A.n_work; w += nw) {
        AlnCtx &c = A.actx[w];
        if (c.phase != PH_FINISH) continue;
        const Opt &o = *c.o;
        for (int i = 0; i < c.n_regs; ++i) {
            Reg &r = c.regs[i];
            if (!r.need_fin) continue;
            if (!r.has_p) { if (lane == 0) r.need_fin = 0; __syncwarp(); continue; }
            const uint8_t *qseq = &c.qseq[r.rev][r.fin_q], *tseq = &c.tseq[r.fin_t];
            int qshift = 0, tshift = 0;
            if (lane == 0) fix_cigar(c, r, qseq, tseq, &qshift, &tshift);
            __syncwarp();
            qshift = __shfl_sync(FULL, qshift, 0); tshift = __shfl_sync(FULL, tshift, 0);
            qseq += qshift; tseq += tshift;
            const uint32_t *cg = c.cig + r.cig;
            const int n_cigar = r.n_cigar;
            int qoff = 0, toff = 0, s = 0, mx = 0, blen = 0, mlen = 0, n_ambi = 0;
            for (int kb = 0; kb < n_cigar; kb += 32) {
                const int k = kb + lane;
                int op = 3, len = 0;
                if (k < n_cigar) { op = (int)(cg[k] & 0xf); len = (int)(cg[k] >> 4); }
                // query / target offsets of this lane's op: exclusive prefix sums of the lengths it consumes
                int ql = (op == 0 || op == 1) ? len : 0, tl = (op == 0 || op == 2) ? len : 0;
                int qp = ql, tp = tl;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int a = __shfl_up_sync(FULL, qp, d), b = __shfl_up_sync(FULL, tp, d);
                    if (lane >= d) qp += a, tp += b;
                }
                const int qtot = __shfl_sync(FULL, qp, 31), ttot = __shfl_sync(FULL, tp, 31);
                const int q0 = qoff + qp - ql, t0 = toff + tp - tl;
                RfMono m; m.A = 0; m.B = NEG; m.MA = NEG; m.MB = NEG;      // identity
                int amb = 0, diff = 0;
                if (op == 0) {
                    int ss = 0, bb = NEG, ma = NEG;
                    for (int l = 0; l < len; ++l) {
                        const int cq = qseq[q0 + l], ct = tseq[t0 + l];
                        if (ct > 3 || cq > 3) ++amb; else if (ct != cq) ++diff;
                        const int x = sc_pair(o, ct, cq);
                        // append the single-base element (A = x, B = 0, MA = x, MB = 0)
                        ma = max(ma, ss + x);
                        bb = max(bb + x, 0);
                        ss += x;
                        m.MB = max(m.MB, bb);
                    }
                    if (len > 0) { m.A = ss; m.B = bb; m.MA = ma; }
                } else if (op == 1 || op == 2) {
                    const uint8_t *sq = op == 1 ? qseq + q0 : tseq + t0;
                    for (int l = 0; l < len; ++l) amb += sq[l] > 3;
                    m.A = -(o.q + o.e * len); m.B = 0;                      // s <- max(0, s - cost); the maximum is not updated
                }
                // ordered reduction: lane 0 ends with op kb .. kb+31 combined left to right
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    RfMono rr;
                    rr.A = __shfl_down_sync(FULL, m.A, d); rr.B = __shfl_down_sync(FULL, m.B, d);
                    rr.MA = __shfl_down_sync(FULL, m.MA, d); rr.MB = __shfl_down_sync(FULL, m.MB, d);
                    if ((lane & (2 * d - 1)) == 0) m = rf_combine(m, rr);
                }
                const int tA = __shfl_sync(FULL, m.A, 0), tB = __shfl_sync(FULL, m.B, 0), tMA = __shfl_sync(FULL, m.MA, 0), tMB = __shfl_sync(FULL, m.MB, 0);
                mx = max(mx, max(s + tMA, tMB));
                s = max(s + tA, tB);
                // counts
                int cb = op == 3 ? 0 : len - amb, cm = op == 0 ? len - (amb + diff) : 0, ca = amb;
#pragma unroll
                for (int d = 16; d; d >>= 1) { cb += __shfl_xor_sync(FULL, cb, d); cm += __shfl_xor_sync(FULL, cm, d); ca += __shfl_xor_sync(FULL, ca, d); }
                blen += cb; mlen += cm; n_ambi += ca;
                qoff += qtot; toff += ttot;
            }
            if (lane == 0) {
                r.blen = blen; r.mlen = mlen; r.n_ambi += n_ambi;
                r.dp_max = mx;          // (int32_t)(mx + .499) of an integer-valued maximum
                r.need_fin = 0;
            }
            __syncwarp();
        }
    }
}

// deferred statistics + final region pass + outputs, one thread per problem
__global__ void __launch_bounds__(128) k_al_finish(const __grid_constant__ AlignArgs A)
{
    const int w = tp_problem<TP_FINISH>(A.n_work);
    if (w < 0) return;
    AlnCtx &c = A.actx[w];
    const int pidx = A.work[w].pidx;
    const int read = A.prob_read[pidx], strand = A.prob_ls[pidx] & 1;
    if (c.phase == PH_FINISH) aln_finish(c);
    if (c.err) atomicOr(A.err, c.err);
    atomicAdd(A.stat_tasks, (unsigned long long)c.n_tasks);
    A.prob_nregs[pidx] = c.n_regs;
    // M-blocks of every record samtools depth would count (everything but SECONDARY)
    int nb = 0, ncg = 0;
    for (int i = 0; i < c.n_regs; ++i) {
        const Reg &r = c.regs[i];
        ncg += r.n_cigar;
        if (r.parent != r.id) continue;
        const uint32_t *cg = c.cig + r.cig;
        for (int k = 0; k < r.n_cigar; ++k) nb += (cg[k] & 0xf) == 0;
    }
    long long boff = (long long)atomicAdd(A.n_blocks, (unsigned long long)nb);
    if (boff + nb > A.blocks_cap) { atomicOr(A.err, 32); nb = 0; boff = 0; }
    A.prob_blk_off[pidx] = boff; A.prob_blk_cnt[pidx] = nb;
    if (nb) {
        int wq = 0;
        for (int i = 0; i < c.n_regs; ++i) {
            const Reg &r = c.regs[i];
            if (r.parent != r.id) continue;
            const uint32_t *cg = c.cig + r.cig;
            int pos = r.rs;
            for (int k = 0; k < r.n_cigar; ++k) {
                int op = cg[k] & 0xf, len = (int)(cg[k] >> 4);
                if (op == 0) { A.blocks[boff + wq++] = make_int2(pos, len); pos += len; }
                else if (op == 2) pos += len;
            }
        }
    }
    if (A.aln_out && c.n_regs > 0) {
        long long ao = (long long)atomicAdd(A.n_aln, (unsigned long long)c.n_regs);
        long long co = (long long)atomicAdd(A.n_cig, (unsigned long long)ncg);
        if (ao + c.n_regs > A.aln_cap || co + ncg > A.cig_out_cap) atomicOr(A.err, 64);
        else {
            for (int i = 0; i < c.n_regs; ++i) {
                const Reg &r = c.regs[i];
                int32_t *oo = A.aln_out + (ao + i) * ALN_REC_INTS;
                oo[0] = read + A.read_base; oo[1] = strand; oo[2] = r.rs; oo[3] = r.re; oo[4] = r.qs; oo[5] = r.qe; oo[6] = r.rev;
                oo[7] = (r.rev ? 0x10 : 0) | (r.parent != r.id ? 0x100 : !r.sam_pri ? 0x800 : 0);
                oo[8] = r.dp_max; oo[9] = r.mlen; oo[10] = r.blen; oo[11] = r.n_cigar;
                oo[12] = (int32_t)(co & 0xffffffffLL); oo[13] = (int32_t)(co >> 32);
                oo[14] = pidx + 2 * A.read_base; oo[15] = i;
                oo[16] = r.mapq; oo[17] = r.dp_score; oo[18] = r.cnt; oo[19] = r.score; oo[20] = r.subsc; oo[21] = r.n_ambi; oo[22] = r.inv; oo[23] = r.n_sub;
                const uint32_t *cg = c.cig + r.cig;
                for (int k = 0; k < r.n_cigar; ++k) A.cig_out[co + k] = cg[k];
                co += r.n_cigar;
            }
        }
    }
}

}  // namespace telr
