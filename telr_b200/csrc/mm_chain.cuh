// mm_chain.cuh — sequential parts of anchor chaining (minimap2 lchain.c) for one read x contig strand.
// The candidate scan of the chaining DP is warp-parallel in k_chain.cuh; what is here runs on one lane
// (or on the host in tests/emu): scoring, backtracking, compaction, RMQ re-chaining.
#pragma once
#include "mm_sort.cuh"

namespace telr {

// fast log2 used by the chaining gap penalty; fp32 without fused multiply-add
TELR_HD float fast_log2(float x)
{
    union { float f; uint32_t i; } z;
    z.f = x;
    float log_2 = (float)(int)(((z.i >> 23) & 255) - 128);
    z.i &= ~(255u << 23);
    z.i += 127u << 23;
    float t = TELR_FADD(TELR_FMUL(-0.34484843f, z.f), 2.02466578f);
    t = TELR_FADD(TELR_FMUL(t, z.f), -0.67487759f);
    return TELR_FADD(log_2, t);
}

TELR_HD int gap_penalty(int dd, int dg, float pen_gap, float pen_skip)
{
    float lin = TELR_FADD(TELR_FMUL(pen_gap, (float)dd), TELR_FMUL(pen_skip, (float)dg));
    float lg = dd >= 1 ? fast_log2((float)(dd + 1)) : 0.0f;
    return (int)TELR_FADD(lin, TELR_FMUL(.5f, lg));
}

// score of extending the chain ending at anchor j with anchor i (first-pass DP); INT32_MIN = not allowed
TELR_HD int32_t link_score(const Anchor &ai, const Anchor &aj, int max_dist_x, int max_dist_y, int bw, float pen_gap,
                           float pen_skip)
{
    int32_t dq = (int32_t)ai.y - (int32_t)aj.y;
    if (dq <= 0 || dq > max_dist_x) return INT32_MIN;
    int32_t dr = (int32_t)(ai.x - aj.x);
    if (dr == 0 || dq > max_dist_y) return INT32_MIN;
    int32_t dd = dr > dq ? dr - dq : dq - dr;
    if (dd > bw) return INT32_MIN;
    int32_t dg = dr < dq ? dr : dq;
    int32_t span = (int32_t)(aj.y >> 32 & 0xff);
    int32_t sc = span < dg ? span : dg;
    if (dd || dg > span) sc -= gap_penalty(dd, dg, pen_gap, pen_skip);
    return sc;
}

// re-chaining variant (no distance tests; reports diagonal width and exactness)
TELR_HD int32_t link_score_simple(const Anchor &ai, const Anchor &aj, float pen_gap, float pen_skip, int *exact, int *width)
{
    int32_t dq = (int32_t)ai.y - (int32_t)aj.y;
    int32_t dr = (int32_t)(ai.x - aj.x);
    int32_t dd = dr > dq ? dr - dq : dq - dr;
    int32_t dg = dr < dq ? dr : dq;
    int32_t span = (int32_t)(aj.y >> 32 & 0xff);
    int32_t sc = span < dg ? span : dg;
    *width = dd;
    if (exact) *exact = (dd == 0 && dg <= span);
    if (dd || dq > span) sc -= gap_penalty(dd, dg, pen_gap, pen_skip);
    return sc;
}

// per-warp scratch for chaining n anchors
struct ChainScratch {
    Anchor *b;        // [n]  compaction copy
    Anchor *z;        // [n]  (score,index) pairs for backtracking / region sort keys
    int32_t *f, *p, *v, *t;   // [n] each
    uint64_t *u, *u2; // [n/3+2]
    int32_t *ord;     // [n] RMQ inner candidates
    int32_t *sortws;  // rs_scratch_words(n)
};
TELR_HD size_t chain_scratch_bytes(size_t n)
{
    size_t s = 0;
    s += (n + 1) * sizeof(Anchor) * 2;
    s += (n + 1) * 4 * 5;
    s += (n / 3 + 4) * 8 * 2;
    s += (size_t)rs_scratch_words((int)n) * 4;
    return (s + 255) / 256 * 256;
}
TELR_HD void chain_scratch_carve(ChainScratch &s, uint8_t *base, size_t n)
{
    s.b = (Anchor *)base; base += (n + 1) * sizeof(Anchor);
    s.z = (Anchor *)base; base += (n + 1) * sizeof(Anchor);
    s.u = (uint64_t *)base; base += (n / 3 + 4) * 8;
    s.u2 = (uint64_t *)base; base += (n / 3 + 4) * 8;
    s.f = (int32_t *)base; base += (n + 1) * 4;
    s.p = (int32_t *)base; base += (n + 1) * 4;
    s.v = (int32_t *)base; base += (n + 1) * 4;
    s.t = (int32_t *)base; base += (n + 1) * 4;
    s.ord = (int32_t *)base; base += (n + 1) * 4;
    s.sortws = (int32_t *)base;
}

// walk one chain back from z[k] until an anchor already used or the score drops by more than max_drop
TELR_HD int32_t bk_end(int max_drop, const Anchor *z, const int32_t *f, const int32_t *p, int32_t *t, int k)
{
    int32_t i = (int32_t)z[k].y, end_i = -1, max_i = i, max_s = 0;
    if (i < 0 || t[i] != 0) return i;
    do {
        t[i] = 2;
        end_i = i = p[i];
        int32_t s = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
        if (s > max_s) max_s = s, max_i = i;
        else if (max_s - s > max_drop) break;
    } while (i >= 0 && t[i] == 0);
    for (i = (int32_t)z[k].y; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
    return max_i;
}

// chains from (f,p): u[] = score<<32|count, v[] = anchor indices (each chain listed end -> start)
TELR_HDN void chain_backtrack(int n, ChainScratch &s, int min_cnt, int min_sc, int max_drop, int *n_u_, int *n_v_)
{
    const int32_t *f = s.f, *p = s.p;
    int32_t *v = s.v, *t = s.t;
    Anchor *z = s.z;
    int n_z = 0, n_u = 0, n_v = 0;
    *n_u_ = *n_v_ = 0;
    for (int i = 0; i < n; ++i)
        if (f[i] >= min_sc) z[n_z].x = (uint64_t)f[i], z[n_z++].y = (uint64_t)i;
    if (n_z == 0) return;
    rs_sort_emul(z, n_z, KeyX(), s.sortws);
    for (int i = 0; i < n; ++i) t[i] = 0;
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
    *n_u_ = n_u, *n_v_ = n_v;
}

// reorder anchors chain by chain (start -> end inside a chain, chains by target position); a[] is overwritten
TELR_HDN void chain_compact(int n_u, int n_v, ChainScratch &s, Anchor *a)
{
    Anchor *b = s.b, *w = s.z;
    uint64_t *u = s.u, *u2 = s.u2;
    int k = 0;
    for (int i = 0; i < n_u; ++i) {
        int k0 = k, ni = (int32_t)u[i];
        for (int j = 0; j < ni; ++j) b[k++] = a[s.v[k0 + (ni - j - 1)]];
    }
    k = 0;
    for (int i = 0; i < n_u; ++i) {
        w[i].x = b[k].x, w[i].y = (uint64_t)k << 32 | (uint32_t)i;
        k += (int32_t)u[i];
    }
    rs_sort_emul(w, n_u, KeyX(), s.sortws);
    k = 0;
    for (int i = 0; i < n_u; ++i) {
        int j = (int32_t)w[i].y, n = (int32_t)u[j];
        u2[i] = u[j];
        const Anchor *src = &b[w[i].y >> 32];
        for (int c = 0; c < n; ++c) a[k + c] = src[c];
        k += n;
    }
    for (int i = 0; i < n_u; ++i) u[i] = u2[i];
    (void)n_v;
}

// first-pass chaining DP, sequential form (tests/emu); k_chain.cuh holds the warp-parallel form
TELR_HDN void chain_dp_seq(const Opt &o, int n, const Anchor *a, ChainScratch &s)
{
    int32_t *f = s.f, *p = s.p, *v = s.v, *t = s.t;
    int max_dist_x = tmax(o.max_gap, o.bw), max_dist_y = tmax(o.max_gap, o.bw);
    int st = 0, max_ii = -1;
    for (int i = 0; i < n; ++i) t[i] = 0;
    for (int i = 0; i < n; ++i) {
        int max_j = -1, end_j, j, n_skip = 0;
        int32_t max_f = (int32_t)(a[i].y >> 32 & 0xff);
        while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + (uint64_t)max_dist_x)) ++st;
        if (i - st > o.max_chain_iter) st = i - o.max_chain_iter;
        for (j = i - 1; j >= st; --j) {
            int32_t sc = link_score(a[i], a[j], max_dist_x, max_dist_y, o.bw, o.chn_pen_gap, o.chn_pen_skip);
            if (sc == INT32_MIN) continue;
            sc += f[j];
            if (sc > max_f) {
                max_f = sc, max_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == i) {
                if (++n_skip > o.max_chain_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = i;
        }
        end_j = j;
        if (max_ii < 0 || a[i].x - a[max_ii].x > (uint64_t)(int64_t)max_dist_x) {
            int32_t mx = INT32_MIN;
            max_ii = -1;
            for (j = i - 1; j >= st; --j)
                if (mx < f[j]) mx = f[j], max_ii = j;
        }
        if (max_ii >= 0 && max_ii < end_j) {
            int32_t tmp = link_score(a[i], a[max_ii], max_dist_x, max_dist_y, o.bw, o.chn_pen_gap, o.chn_pen_skip);
            if (tmp != INT32_MIN && max_f < tmp + f[max_ii]) max_f = tmp + f[max_ii], max_j = max_ii;
        }
        f[i] = max_f, p[i] = max_j;
        v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
        if (max_ii < 0 || (a[i].x - a[max_ii].x <= (uint64_t)(int64_t)max_dist_x && f[max_ii] < f[i])) max_ii = i;
    }
}

// RMQ priority of a finished anchor (smaller = better); double arithmetic without contraction
TELR_HD double rmq_pri(const Anchor &aj, int32_t fj, float pen_gap)
{
    double w = TELR_DMUL(TELR_DMUL(0.5, (double)pen_gap), (double)((int32_t)aj.x + (int32_t)aj.y));
    return -TELR_DADD((double)fj, w);
}

// true if key (y1, j1) should replace (y0, j0) as RMQ answer given their priorities
TELR_HD bool rmq_better(double pri1, int32_t y1, int j1, double pri0, int32_t y0, int j0)
{
    if (pri1 < pri0) return true;
    if (pri1 > pri0) return false;
    return y1 > y0 || (y1 == y0 && j1 > j0);
}

// one anchor of the re-chaining pass given the outer RMQ answer `best` (-1 = none); sequential inner scan.
TELR_HD void rmq_step(const Opt &o, int n, int i, const Anchor *a, ChainScratch &s, int best, int st_inner, int i0,
                      int max_dist_inner)
{
    int32_t *f = s.f, *p = s.p, *v = s.v, *t = s.t, *ord = s.ord;
    int max_j = -1;
    int32_t max_f = (int32_t)(a[i].y >> 32 & 0xff);
    if (best >= 0) {
        int exact, width, n_skip = 0, j = best;
        int32_t sc = f[j] + link_score_simple(a[i], a[j], o.chn_pen_gap, o.chn_pen_skip, &exact, &width);
        if (width <= o.bw_long && sc > max_f) max_f = sc, max_j = j;
        if (!exact && max_dist_inner > 0 && (int32_t)a[i].y > 0) {
            int m = 0;
            int32_t yi = (int32_t)a[i].y;
            for (j = st_inner < i0 ? st_inner : i0; j < i0; ++j)
                if ((int32_t)a[j].y <= yi - 1 && (int32_t)a[j].y >= yi - max_dist_inner) ord[m++] = j;
            for (int c = 1; c < m; ++c) {        // by (y, j) descending
                int tj = ord[c], d = c;
                while (d > 0) {
                    int oo = ord[d - 1];
                    int32_t yo = (int32_t)a[oo].y, yt = (int32_t)a[tj].y;
                    if (yo > yt || (yo == yt && oo > tj)) break;
                    ord[d] = oo; --d;
                }
                ord[d] = tj;
            }
            for (int c = 0; c < m; ++c) {
                int width2;
                j = ord[c];
                sc = f[j] + link_score_simple(a[i], a[j], o.chn_pen_gap, o.chn_pen_skip, 0, &width2);
                if (width2 <= o.bw_long) {
                    if (sc > max_f) {
                        max_f = sc, max_j = j;
                        if (n_skip > 0) --n_skip;
                    } else if (t[j] == i) {
                        if (++n_skip > o.max_chain_skip) break;
                    }
                    if (p[j] >= 0) t[p[j]] = i;
                }
            }
        }
    }
    f[i] = max_f, p[i] = max_j;
    v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
    (void)n;
}

// ---- high-occurrence seeds of one read (minimap2 seed.c mm_seed_select + the rep_len bookkeeping of mm_collect_matches) ----
// ks_heapdown / ks_heapmake of klib's ksort.h on uint64 (max-heap)
TELR_HD void heapdown_u64(int i, int n, uint64_t *l)
{
    int k = i;
    uint64_t tmp = l[i];
    while ((k = (k << 1) + 1) < n) {
        if (k != n - 1 && l[k] < l[k + 1]) ++k;
        if (l[k] < tmp) break;
        l[i] = l[k]; i = k;
    }
    l[i] = tmp;
}
// kept[0..nk): the read's minimizers after the query-side filter; tarr[j] = occurrences of kept[j] in the contig index.
// Marks every seed upstream would filter by NEGATING tarr[j] and returns rep_len, the query bases under filtered seeds.
// Inside a streak of consecutive high-occurrence seeds (among the seeds that hit the index) the (streak span / occ_dist)
// seeds with the fewest occurrences survive; seeds above max_max_occ never do.  cidx: scratch, nk words.  Sequential.
TELR_HDN int seeds_filter(int max_occ, int max_max_occ, int occ_dist, int nk, const int32_t *kept, const uint64_t *qx, const uint32_t *qy,
                          int32_t *tarr, int32_t *cidx, int qlen)
{
    int n = 0;
    for (int j = 0; j < nk; ++j) if (tarr[j] > 0) cidx[n++] = j;
#define SF_T(i) tarr[cidx[i]]
#define SF_QPOS(i) ((int32_t)(qy[kept[cidx[i]]] >> 1))
    if (occ_dist > 0 && max_max_occ > max_occ) {
        if (n > 1) {
            uint64_t b[128];
            int last0 = -1;
            for (int i = 0; i <= n; ++i) {
                if (i == n || SF_T(i) <= max_occ) {
                    if (i - last0 > 1) {
                        const int ps = last0 < 0 ? 0 : SF_QPOS(last0), pe = i == n ? qlen : SF_QPOS(i);
                        const int st = last0 + 1, en = i;
                        int j, k, max_high_occ = (int)((double)(pe - ps) / occ_dist + .499);
                        if (max_high_occ > 0) {
                            if (max_high_occ > 128) max_high_occ = 128;
                            for (j = st, k = 0; j < en && k < max_high_occ; ++j, ++k) b[k] = (uint64_t)(uint32_t)SF_T(j) << 32 | (uint32_t)j;
                            for (int h = (k >> 1) - 1; h >= 0; --h) heapdown_u64(h, k, b);
                            for (; j < en; ++j)
                                if (SF_T(j) < (int32_t)(b[0] >> 32)) { b[0] = (uint64_t)(uint32_t)SF_T(j) << 32 | (uint32_t)j; heapdown_u64(0, k, b); }
                            for (j = 0; j < k; ++j) SF_T((uint32_t)b[j]) = -SF_T((uint32_t)b[j]);          // flt = 1 on the chosen ...
                        }
                        for (j = st; j < en; ++j) SF_T(j) = -SF_T(j);                                        // ... then flt ^= 1 over the streak
                        for (j = st; j < en; ++j) { const int t = SF_T(j) < 0 ? -SF_T(j) : SF_T(j); if (t > max_max_occ) SF_T(j) = -t; }
                    }
                    last0 = i;
                }
            }
        }
    } else {
        for (int i = 0; i < n; ++i) if (SF_T(i) > max_occ) SF_T(i) = -SF_T(i);
    }
    int rep_st = 0, rep_en = 0, rep_len = 0;
    for (int i = 0; i < n; ++i)
        if (SF_T(i) < 0) {
            const int en = SF_QPOS(i) + 1, st = en - (int)(qx[kept[cidx[i]]] & 0xff);
            if (st > rep_en) { rep_len += rep_en - rep_st; rep_st = st; rep_en = en; }
            else rep_en = en;
        }
#undef SF_T
#undef SF_QPOS
    return rep_len + (rep_en - rep_st);
}

struct RmqWin { int st, st_inner, i0, max_dist, max_dist_inner; };
TELR_HD void rmq_win_init(RmqWin &w, const Opt &o)
{
    w.st = w.st_inner = w.i0 = 0;
    w.max_dist = tmax(o.max_gap, o.bw_long);
    w.max_dist_inner = tmin(tmax(o.rmq_inner_dist, 0), w.max_dist);
}
// advance the active windows for anchor i
TELR_HD void rmq_win_advance(RmqWin &w, const Opt &o, int i, const Anchor *a)
{
    if (w.i0 < i && a[w.i0].x != a[i].x) w.i0 = i;
    while (w.st < i && (a[i].x >> 32 != a[w.st].x >> 32 || a[i].x > a[w.st].x + (uint64_t)w.max_dist || w.i0 - w.st > o.rmq_size_cap)) ++w.st;
    if (w.max_dist_inner > 0)
        while (w.st_inner < i && (a[i].x >> 32 != a[w.st_inner].x >> 32 || a[i].x > a[w.st_inner].x + (uint64_t)w.max_dist_inner ||
                                  w.i0 - w.st_inner > o.rmq_size_cap)) ++w.st_inner;
}
TELR_HD bool rmq_in_range(const Anchor &aj, int j, int32_t yi, int max_dist)
{
    int32_t yj = (int32_t)aj.y, lo_y = yi - max_dist;
    if (yj < lo_y || yj == lo_y) return false;            // lo = (lo_y, INT32_MAX): j < INT32_MAX always
    if (yj > yi || (yj == yi && j > 0)) return false;     // hi = (yi, 0)
    return true;
}

// sequential re-chaining pass (tests/emu)
TELR_HDN void chain_rmq_seq(const Opt &o, int n, const Anchor *a, ChainScratch &s)
{
    RmqWin w;
    rmq_win_init(w, o);
    for (int i = 0; i < n; ++i) s.t[i] = -1, s.v[i] = 0;
    for (int i = 0; i < n; ++i) {
        rmq_win_advance(w, o, i, a);
        int best = -1;
        double bp = 0.0;
        for (int j = w.st < w.i0 ? w.st : w.i0; j < w.i0; ++j) {
            if (!rmq_in_range(a[j], j, (int32_t)a[i].y, w.max_dist)) continue;
            double pri = rmq_pri(a[j], s.f[j], o.chn_pen_gap);
            if (best < 0 || rmq_better(pri, (int32_t)a[j].y, j, bp, (int32_t)a[best].y, best)) best = j, bp = pri;
        }
        rmq_step(o, n, i, a, s, best, w.st_inner, w.i0, w.max_dist_inner);
    }
}

}  // namespace telr
