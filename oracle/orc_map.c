/*
 * oracle/orc_map.c — seeding, chaining, region bookkeeping and base-level alignment of ONE read
 * against ONE contig strand.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see orc.h).
 *
 * Restates, for the single-segment long-read code path that `minimap2 -a -x map-ont|map-pb contig
 * reads` (TELR_te.py:505) exercises, these [UP] minimap2 2.22 routines:
 *   index.c   mm_idx_gen (bucket sort of contig minimizers), mm_idx_get, mm_idx_cal_max_occ
 *   seed.c    mm_seed_mz_flt, mm_seed_collect_all, mm_collect_matches
 *   map.c     collect_seed_hits, mm_map_frag, chain_post, align_regs
 *   lchain.c  mg_lchain_dp, mg_chain_backtrack, mg_lchain_rmq, compact_a
 *   hit.c     mm_gen_regs, mm_set_parent, mm_select_sub, mm_sync_regs, mm_filter_regs, mm_hit_sort,
 *             mm_set_sam_pri, mm_split_reg, mm_squeeze_a, mm_update_dp_max
 *   align.c   mm_align_skeleton, mm_align1, mm_fix_bad_ends, mm_filter_bad_seeds(_alt),
 *             mm_adjust_minier, mm_test_zdrop, mm_fix_cigar, mm_update_extra, mm_align1_inv
 *   seed.c    mm_seed_select;  hit.c mm_set_mapq, mm_set_inv_mapq
 * Not restated (does not influence coordinates/CIGAR/secondary status/MAPQ): mm_est_err (dv:f tag only).
 * Stated deviations: (2) krmq ties in the RMQ priority are resolved to the largest (y,i) key instead of by
 * AVL shape; (3) clean DP band (orc_ksw.c); logf in mm_set_mapq is the correctly rounded float logarithm.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <assert.h>
#include "orc.h"

#define SEED_LONG_JOIN (1ULL << 40)
#define SEED_IGNORE    (1ULL << 41)
#define SEED_TANDEM    (1ULL << 42)
#define PARENT_UNSET   (-1)
#define PARENT_TMP_PRI (-2)

/* ---------------- tiny arena so that struct copies of regs can share cigar storage ---------------- */
typedef struct ablk { struct ablk *next; } ablk_t;
typedef struct { ablk_t *head; } arena_t;
static void *aalloc(arena_t *A, size_t n)
{
    ablk_t *b = (ablk_t *)malloc(sizeof(ablk_t) + 8 + n);
    b->next = A->head; A->head = b;
    return (char *)b + sizeof(ablk_t) + 8 - (sizeof(ablk_t) % 8);
}
static void afree_all(arena_t *A)
{
    while (A->head) { ablk_t *n = A->head->next; free(A->head); A->head = n; }
}

typedef struct {
    int32_t id, cnt, rid, score, qs, qe, rs, re, parent, subsc, as, mlen, blen, n_sub, score0;
    uint32_t hash;
    int32_t rev, inv, sam_pri, split, split_inv, strand_retained;
    /* mm_extra_t */
    int32_t has_p, dp_score, dp_max, dp_max2, n_ambi, n_cigar, m_cigar;
    uint32_t *cigar;
} reg_t;

/* ------------------------------- index ------------------------------- */
typedef struct {
    int b;
    int64_t n;
    orc128_t *a;      /* minimizers grouped by bucket, each bucket sorted by radix_sort_128x */
    int64_t *boff;    /* [2^b + 1] */
    int32_t mid_occ, n_keys;
} idx_t;

int64_t orc_dev[8];
void orc_dev_counters(int64_t out[8], int reset)
{
    int i;
    for (i = 0; i < 8; ++i) { if (out) out[i] = orc_dev[i]; if (reset) orc_dev[i] = 0; }
}
#define DEV_COUNT(k) do { _Pragma("omp atomic") ++orc_dev[k]; } while (0)

static void idx_build(idx_t *mi, const orc_opt_t *opt, const uint8_t *seq, int32_t len)
{
    int64_t cap = len + 16, n, i;
    uint64_t *x = (uint64_t *)malloc((size_t)cap * 8), *y = (uint64_t *)malloc((size_t)cap * 8);
    int nb = 1 << 14, mask = nb - 1;
    n = orc_sketch(seq, len, opt->w, opt->k, opt->hpc, x, y, cap);
    assert(n <= cap);
    mi->b = 14; mi->n = n;
    mi->a = (orc128_t *)malloc((size_t)(n + 1) * 16);
    mi->boff = (int64_t *)calloc((size_t)nb + 1, 8);
    for (i = 0; i < n; ++i) ++mi->boff[((x[i] >> 8) & mask) + 1];
    for (i = 0; i < nb; ++i) mi->boff[i + 1] += mi->boff[i];
    {
        int64_t *fill = (int64_t *)malloc((size_t)nb * 8);
        memcpy(fill, mi->boff, (size_t)nb * 8);
        for (i = 0; i < n; ++i) {       /* mm_idx_add: appended in sketch order */
            int64_t p = fill[(x[i] >> 8) & mask]++;
            mi->a[p].x = x[i], mi->a[p].y = y[i];
        }
        free(fill);
    }
    for (i = 0; i < nb; ++i)            /* worker_post: radix_sort_128x per bucket */
        if (mi->boff[i + 1] - mi->boff[i] > 1) {
            if (mi->boff[i + 1] - mi->boff[i] > 64) DEV_COUNT(4);
            orc_radix_sort_128x(mi->a + mi->boff[i], mi->a + mi->boff[i + 1]);
        }
    free(x); free(y);
    /* mm_idx_cal_max_occ + mm_mapopt_update */
    {
        int64_t nk = 0, j;
        uint32_t *cnt = (uint32_t *)malloc((size_t)(n + 1) * 4);
        for (i = 0; i < nb; ++i) {
            int64_t s = mi->boff[i], e = mi->boff[i + 1], st;
            for (st = s, j = s + 1; j <= e; ++j)
                if (j == e || mi->a[j].x >> 8 != mi->a[st].x >> 8) {
                    if (e > s) cnt[nk++] = (uint32_t)(j - st);
                    st = j;
                }
        }
        mi->n_keys = (int32_t)nk;
        if (nk == 0 || opt->mid_occ_frac <= 0.f) mi->mid_occ = INT32_MAX;
        else {
            uint32_t kk = (uint32_t)((1. - opt->mid_occ_frac) * nk);
            /* ks_ksmall: k-th smallest (0-based) */
            int64_t c;
            uint32_t best = 0;
            for (c = 1;; ++c) {     /* counts are small: find smallest c with #(cnt<=c) > kk */
                int64_t le = 0;
                for (j = 0; j < nk; ++j) le += cnt[j] <= (uint32_t)c;
                if (le > kk) { best = (uint32_t)c; break; }
            }
            mi->mid_occ = (int32_t)(best + 1);
        }
        if (mi->mid_occ < opt->min_mid_occ) mi->mid_occ = opt->min_mid_occ;
        if (opt->max_mid_occ > opt->min_mid_occ && mi->mid_occ > opt->max_mid_occ) mi->mid_occ = opt->max_mid_occ;
        free(cnt);
    }
}
static void idx_free(idx_t *mi) { free(mi->a); free(mi->boff); }

static const orc128_t *idx_get(const idx_t *mi, uint64_t minier, int *n)
{
    int64_t s = mi->boff[minier & ((1 << mi->b) - 1)], e = mi->boff[(minier & ((1 << mi->b) - 1)) + 1], i, j;
    *n = 0;
    for (i = s; i < e; ++i)
        if (mi->a[i].x >> 8 == minier) {
            for (j = i; j < e && mi->a[j].x >> 8 == minier; ++j) {}
            *n = (int)(j - i);
            return &mi->a[i];
        }
    return 0;
}

/* ------------------------------- seeding ------------------------------- */
/* [UP] seed.c mm_seed_mz_flt */
static int64_t seed_mz_flt(int64_t n, uint64_t *mx, uint64_t *my, int32_t q_occ_max, float q_occ_frac)
{
    orc128_t *a;
    int64_t i, j, st;
    if (n <= q_occ_max || q_occ_frac <= 0.0f || q_occ_max <= 0) return n;
    a = (orc128_t *)malloc((size_t)n * 16);
    for (i = 0; i < n; ++i) a[i].x = mx[i], a[i].y = (uint64_t)i;
    orc_radix_sort_128x(a, a + n);
    for (st = 0, i = 1; i <= n; ++i) {
        if (i == n || a[i].x != a[st].x) {
            int32_t cnt = (int32_t)(i - st);
            if (cnt > q_occ_max && cnt > n * q_occ_frac)
                for (j = st; j < i; ++j) mx[a[j].y] = 0;
            st = i;
        }
    }
    free(a);
    for (i = j = 0; i < n; ++i)
        if (mx[i] != 0) mx[j] = mx[i], my[j] = my[i], ++j;
    return j;
}

/* [UP] seed.c mm_seed_select: inside every streak of consecutive high-occurrence seeds keep the (streak length / dist)
 * seeds with the fewest occurrences (a max-heap of n<<32|index, ksort.h ks_heapmake / ks_heapdown), drop the rest */
#define MAX_MAX_HIGH_OCC 128
typedef struct { uint32_t q_pos, q_span; int32_t n, flt, is_tandem; const orc128_t *cr; } seed_t;
static void heapdown_u64(size_t i, size_t n, uint64_t *l)
{
    size_t k = i;
    uint64_t tmp = l[i];
    while ((k = (k << 1) + 1) < n) {
        if (k != n - 1 && l[k] < l[k + 1]) ++k;
        if (l[k] < tmp) break;
        l[i] = l[k]; i = k;
    }
    l[i] = tmp;
}
static void heapmake_u64(size_t n, uint64_t *l)
{
    size_t i;
    for (i = (n >> 1) - 1; i != (size_t)(-1); --i) heapdown_u64(i, n, l);
}
static void seed_select(int32_t n, seed_t *a, int len, int max_occ, int max_max_occ, int dist)
{
    int32_t i, last0, m;
    uint64_t b[MAX_MAX_HIGH_OCC];
    if (n == 0 || n == 1) return;
    for (i = m = 0; i < n; ++i)
        if (a[i].n > max_occ) ++m;
    if (m == 0) return;
    for (i = 0, last0 = -1; i <= n; ++i) {
        if (i == n || a[i].n <= max_occ) {
            if (i - last0 > 1) {
                int32_t ps = last0 < 0 ? 0 : (int32_t)(a[last0].q_pos >> 1);
                int32_t pe = i == n ? len : (int32_t)(a[i].q_pos >> 1);
                int32_t j, k, st = last0 + 1, en = i;
                int32_t max_high_occ = (int32_t)((double)(pe - ps) / dist + .499);
                if (max_high_occ > 0) {
                    if (max_high_occ > MAX_MAX_HIGH_OCC) max_high_occ = MAX_MAX_HIGH_OCC;
                    for (j = st, k = 0; j < en && k < max_high_occ; ++j, ++k) b[k] = (uint64_t)a[j].n << 32 | (uint32_t)j;
                    heapmake_u64((size_t)k, b);
                    for (; j < en; ++j)
                        if (a[j].n < (int32_t)(b[0] >> 32)) {
                            b[0] = (uint64_t)a[j].n << 32 | (uint32_t)j;
                            heapdown_u64(0, (size_t)k, b);
                        }
                    for (j = 0; j < k; ++j) a[(uint32_t)b[j]].flt = 1;
                }
                for (j = st; j < en; ++j) a[j].flt ^= 1;
                for (j = st; j < en; ++j)
                    if (a[j].n > max_max_occ) a[j].flt = 1;
            }
            last0 = i;
        }
    }
}

/* [UP] mm_seed_collect_all + mm_collect_matches + collect_seed_hits; *rep_len_ = query bases under filtered seeds */
static orc128_t *collect_seed_hits(const orc_opt_t *opt, const idx_t *mi, int qlen, int64_t n_mz, const uint64_t *mx, const uint64_t *my,
                                   int64_t *n_a_, int32_t *rep_len_)
{
    int64_t i, n_a = 0, cap = 0;
    int32_t n_m = 0, rep_st = 0, rep_en = 0, rep_len = 0;
    orc128_t *a;
    seed_t *m = (seed_t *)malloc((size_t)(n_mz + 1) * sizeof(seed_t));
    for (i = 0; i < n_mz; ++i) {
        int t;
        const orc128_t *cr = idx_get(mi, mx[i] >> 8, &t);
        seed_t *q;
        if (t == 0) continue;
        q = &m[n_m++];
        q->q_pos = (uint32_t)my[i], q->q_span = (uint32_t)(mx[i] & 0xff), q->cr = cr, q->n = t, q->flt = 0, q->is_tandem = 0;
        if (i > 0 && mx[i] >> 8 == mx[i - 1] >> 8) q->is_tandem = 1;
        if (i < n_mz - 1 && mx[i] >> 8 == mx[i + 1] >> 8) q->is_tandem = 1;
        if (t > mi->mid_occ) DEV_COUNT(0);      /* counts the high-occurrence seeds (mm_seed_select decides their fate) */
    }
    if (opt->occ_dist > 0 && opt->max_max_occ > mi->mid_occ) seed_select(n_m, m, qlen, mi->mid_occ, opt->max_max_occ, opt->occ_dist);
    else for (i = 0; i < n_m; ++i) if (m[i].n > mi->mid_occ) m[i].flt = 1;
    for (i = 0; i < n_m; ++i) {
        const seed_t *q = &m[i];
        if (q->flt) {
            int en = (int)(q->q_pos >> 1) + 1, st = en - (int)q->q_span;
            if (st > rep_en) { rep_len += rep_en - rep_st; rep_st = st, rep_en = en; }
            else rep_en = en;
        } else cap += q->n;
    }
    rep_len += rep_en - rep_st;
    a = (orc128_t *)malloc((size_t)(cap + 1) * 16);
    for (i = 0; i < n_m; ++i) {
        const seed_t *q = &m[i];
        int k;
        if (q->flt) continue;
        for (k = 0; k < q->n; ++k) {
            uint64_t r = q->cr[k].y;
            int32_t rpos = (uint32_t)r >> 1;
            orc128_t *p = &a[n_a++];
            if ((r & 1) == (q->q_pos & 1)) {
                p->x = (r & 0xffffffff00000000ULL) | (uint64_t)rpos;
                p->y = (uint64_t)q->q_span << 32 | q->q_pos >> 1;
            } else {
                p->x = 1ULL << 63 | (r & 0xffffffff00000000ULL) | (uint64_t)rpos;
                p->y = (uint64_t)q->q_span << 32 | (uint32_t)(qlen - ((int32_t)(q->q_pos >> 1) + 1 - (int32_t)q->q_span) - 1);
            }
            if (q->is_tandem) p->y |= SEED_TANDEM;
        }
    }
    free(m);
    orc_radix_sort_128x(a, a + n_a);
    *n_a_ = n_a;
    *rep_len_ = rep_len;
    return a;
}

/* ------------------------------- chaining ------------------------------- */
/* [UP] mmpriv.h mg_log2 */
static inline float mg_log2(float x)
{
    union { float f; uint32_t i; } z;
    float log_2;
    z.f = x;
    log_2 = (float)(((z.i >> 23) & 255) - 128);
    z.i &= ~(255u << 23);
    z.i += 127u << 23;
    log_2 += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
    return log_2;
}

/* [UP] lchain.c comput_sc (single segment, not cDNA) */
static inline int32_t comput_sc(const orc128_t *ai, const orc128_t *aj, int32_t max_dist_x, int32_t max_dist_y,
                                int32_t bw, float chn_pen_gap, float chn_pen_skip)
{
    int32_t dq = (int32_t)ai->y - (int32_t)aj->y, dr, dd, dg, q_span, sc;
    if (dq <= 0 || dq > max_dist_x) return INT32_MIN;
    dr = (int32_t)(ai->x - aj->x);
    if (dr == 0 || dq > max_dist_y) return INT32_MIN;
    dd = dr > dq ? dr - dq : dq - dr;
    if (dd > bw) return INT32_MIN;
    dg = dr < dq ? dr : dq;
    q_span = aj->y >> 32 & 0xff;
    sc = q_span < dg ? q_span : dg;
    if (dd || dg > q_span) {
        float lin_pen, log_pen;
        lin_pen = chn_pen_gap * (float)dd + chn_pen_skip * (float)dg;
        log_pen = dd >= 1 ? mg_log2((float)(dd + 1)) : 0.0f;
        sc -= (int)(lin_pen + .5f * log_pen);
    }
    return sc;
}

/* [UP] lchain.c mg_chain_bk_end */
static int64_t chain_bk_end(int32_t max_drop, const orc128_t *z, const int32_t *f, const int64_t *p, int32_t *t, int64_t k)
{
    int64_t i = (int64_t)z[k].y, end_i = -1, max_i = i;
    int32_t max_s = 0;
    if (i < 0 || t[i] != 0) return i;
    do {
        int32_t s;
        t[i] = 2;
        end_i = i = p[i];
        s = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
        if (s > max_s) max_s = s, max_i = i;
        else if (max_s - s > max_drop) break;
    } while (i >= 0 && t[i] == 0);
    for (i = (int64_t)z[k].y; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
    return max_i;
}

/* [UP] lchain.c mg_chain_backtrack (second, populating pass; the first pass only sizes u[]) */
static uint64_t *chain_backtrack(int64_t n, const int32_t *f, const int64_t *p, int32_t *v, int32_t *t,
                                 int32_t min_cnt, int32_t min_sc, int32_t max_drop, int32_t *n_u_, int32_t *n_v_)
{
    orc128_t *z;
    uint64_t *u;
    int64_t i, k, n_z, n_v;
    int32_t n_u;
    *n_u_ = *n_v_ = 0;
    for (i = 0, n_z = 0; i < n; ++i)
        if (f[i] >= min_sc) ++n_z;
    if (n_z == 0) return 0;
    z = (orc128_t *)malloc((size_t)n_z * 16);
    for (i = 0, k = 0; i < n; ++i)
        if (f[i] >= min_sc) z[k].x = (uint64_t)f[i], z[k++].y = (uint64_t)i;
    orc_radix_sort_128x(z, z + n_z);
    u = (uint64_t *)malloc((size_t)n_z * 8);
    memset(t, 0, (size_t)n * 4);
    for (k = n_z - 1, n_v = n_u = 0; k >= 0; --k) {
        if (t[z[k].y] == 0) {
            int64_t n_v0 = n_v, end_i;
            int32_t sc;
            end_i = chain_bk_end(max_drop, z, f, p, t, k);
            for (i = (int64_t)z[k].y; i != end_i; i = p[i]) v[n_v++] = (int32_t)i, t[i] = 1;
            sc = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
            if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) u[n_u++] = (uint64_t)sc << 32 | (uint64_t)(n_v - n_v0);
            else n_v = n_v0;
        }
    }
    free(z);
    *n_u_ = n_u, *n_v_ = (int32_t)n_v;
    return u;
}

/* [UP] lchain.c compact_a */
static orc128_t *compact_a(int32_t n_u, uint64_t *u, int32_t n_v, int32_t *v, orc128_t *a)
{
    orc128_t *b, *w;
    uint64_t *u2;
    int64_t i, j, k;
    b = (orc128_t *)malloc((size_t)(n_v + 1) * 16);
    for (i = 0, k = 0; i < n_u; ++i) {
        int32_t k0 = (int32_t)k, ni = (int32_t)u[i];
        for (j = 0; j < ni; ++j) b[k++] = a[v[k0 + (ni - j - 1)]];
    }
    w = (orc128_t *)malloc((size_t)n_u * 16);
    for (i = k = 0; i < n_u; ++i) {
        w[i].x = b[k].x, w[i].y = (uint64_t)k << 32 | (uint64_t)i;
        k += (int32_t)u[i];
    }
    orc_radix_sort_128x(w, w + n_u);
    u2 = (uint64_t *)malloc((size_t)n_u * 8);
    for (i = k = 0; i < n_u; ++i) {
        int32_t jj = (int32_t)w[i].y, n = (int32_t)u[jj];
        u2[i] = u[jj];
        memcpy(&a[k], &b[w[i].y >> 32], (size_t)n * 16);
        k += n;
    }
    memcpy(u, u2, (size_t)n_u * 8);
    memcpy(b, a, (size_t)k * 16);
    free(a); free(w); free(u2);
    return b;
}

/* [UP] lchain.c mg_lchain_dp (n_seg = 1, is_cdna = 0) */
static orc128_t *lchain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                           float chn_pen_gap, float chn_pen_skip, int64_t n, orc128_t *a, int *n_u_, uint64_t **_u)
{
    int32_t *f, *t, *v, n_u, n_v, max_drop = bw;
    int64_t *p, i, j, max_ii, st = 0;
    uint64_t *u;
    *_u = 0, *n_u_ = 0;
    if (n == 0 || a == 0) { free(a); return 0; }
    if (max_dist_x < bw) max_dist_x = bw;
    if (max_dist_y < bw) max_dist_y = bw;
    p = (int64_t *)malloc((size_t)n * 8);
    f = (int32_t *)malloc((size_t)n * 4);
    v = (int32_t *)malloc((size_t)n * 4);
    t = (int32_t *)calloc((size_t)n, 4);
    for (i = 0, max_ii = -1; i < n; ++i) {
        int64_t max_j = -1, end_j;
        int32_t max_f = a[i].y >> 32 & 0xff, n_skip = 0;
        while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + (uint64_t)max_dist_x)) ++st;
        if (i - st > max_iter) st = i - max_iter;
        for (j = i - 1; j >= st; --j) {
            int32_t sc = comput_sc(&a[i], &a[j], max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip);
            if (sc == INT32_MIN) continue;
            sc += f[j];
            if (sc > max_f) {
                max_f = sc, max_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == (int32_t)i) {
                if (++n_skip > max_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = (int32_t)i;
        }
        end_j = j;
        if (max_ii < 0 || a[i].x - a[max_ii].x > (uint64_t)(int64_t)max_dist_x) {
            int32_t max = INT32_MIN;
            max_ii = -1;
            for (j = i - 1; j >= st; --j)
                if (max < f[j]) max = f[j], max_ii = j;
        }
        if (max_ii >= 0 && max_ii < end_j) {
            int32_t tmp = comput_sc(&a[i], &a[max_ii], max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip);
            if (tmp != INT32_MIN && max_f < tmp + f[max_ii]) max_f = tmp + f[max_ii], max_j = max_ii;
        }
        f[i] = max_f, p[i] = max_j;
        v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
        if (max_ii < 0 || (a[i].x - a[max_ii].x <= (uint64_t)(int64_t)max_dist_x && f[max_ii] < f[i])) max_ii = i;
    }
    u = chain_backtrack(n, f, p, v, t, min_cnt, min_sc, max_drop, &n_u, &n_v);
    *n_u_ = n_u, *_u = u;
    free(p); free(f); free(t);
    if (n_u == 0) { free(a); free(v); free(u); *_u = 0; return 0; }
    a = compact_a(n_u, u, n_v, v, a);
    free(v);
    return a;
}

/* [UP] lchain.c comput_sc_simple */
static inline int32_t comput_sc_simple(const orc128_t *ai, const orc128_t *aj, float chn_pen_gap, float chn_pen_skip,
                                       int32_t *exact, int32_t *width)
{
    int32_t dq = (int32_t)ai->y - (int32_t)aj->y, dr, dd, dg, q_span, sc;
    dr = (int32_t)(ai->x - aj->x);
    *width = dd = dr > dq ? dr - dq : dq - dr;
    dg = dr < dq ? dr : dq;
    q_span = aj->y >> 32 & 0xff;
    sc = q_span < dg ? q_span : dg;
    if (exact) *exact = (dd == 0 && dg <= q_span);
    if (dd || dq > q_span) {
        float lin_pen, log_pen;
        lin_pen = chn_pen_gap * (float)dd + chn_pen_skip * (float)dg;
        log_pen = dd >= 1 ? mg_log2((float)(dd + 1)) : 0.0f;
        sc -= (int)(lin_pen + .5f * log_pen);
    }
    return sc;
}

/* [UP] lchain.c mg_lchain_rmq.  The two krmq (AVL) trees are restated as their contents:
 *   outer = { j in [st, i0) }, inner = { j in [st_inner, i0) }  (inserted in batches when x changes),
 * with key order (y, j) and priority pri(j) = -(f[j] + 0.5*chn_pen_gap*(x_j + y_j)) in double.
 * krmq_rmq(lo, hi) = element of minimum pri among keys in the closed interval; deviation (2): ties
 * go to the largest key.  The size cap (rmq_size_cap = 100000 elements) is honoured through st. */
static orc128_t *lchain_rmq(int max_dist, int max_dist_inner, int bw, int max_chn_skip, int cap_rmq_size, int min_cnt,
                            int min_sc, float chn_pen_gap, float chn_pen_skip, int64_t n, orc128_t *a, int *n_u_,
                            uint64_t **_u)
{
    int32_t *f, *t, *v, n_u, n_v, max_drop = bw;
    int64_t *p, i, i0, st = 0, st_inner = 0, j;
    uint64_t *u;
    int32_t *ord;   /* scratch: inner candidates sorted by key descending */
    *_u = 0, *n_u_ = 0;
    if (n == 0 || a == 0) { free(a); return 0; }
    if (max_dist < bw) max_dist = bw;
    if (max_dist_inner < 0) max_dist_inner = 0;
    if (max_dist_inner > max_dist) max_dist_inner = max_dist;
    p = (int64_t *)malloc((size_t)n * 8);
    f = (int32_t *)malloc((size_t)n * 4);
    t = (int32_t *)malloc((size_t)n * 4);
    v = (int32_t *)calloc((size_t)n, 4);
    ord = (int32_t *)malloc((size_t)n * 4);
    for (i = 0; i < n; ++i) t[i] = -1;   /* upstream leaves t[] uninitialised; it is only compared with i after being set */
    for (i = i0 = 0; i < n; ++i) {
        int64_t max_j = -1;
        int32_t q_span = a[i].y >> 32 & 0xff, max_f = q_span;
        if (i0 < i && a[i0].x != a[i].x) i0 = i;         /* batch insert [old i0, i) into both trees */
        while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + (uint64_t)max_dist || i0 - st > cap_rmq_size)) ++st;
        if (max_dist_inner > 0)
            while (st_inner < i && (a[i].x >> 32 != a[st_inner].x >> 32 || a[i].x > a[st_inner].x + (uint64_t)max_dist_inner ||
                                    i0 - st_inner > cap_rmq_size)) ++st_inner;
        {   /* krmq_rmq over the outer tree */
            int32_t lo_y = (int32_t)a[i].y - max_dist, hi_y = (int32_t)a[i].y;
            int64_t best = -1;
            double best_pri = 0.0;
            int tied = 0;
            for (j = st < i0 ? st : i0; j < i0; ++j) {
                int32_t yj = (int32_t)a[j].y;
                double pri;
                /* lo = (lo_y, INT32_MAX) <= (yj, j) <= hi = (hi_y, 0) */
                if (yj < lo_y || (yj == lo_y && j < INT32_MAX)) continue;
                if (yj > hi_y || (yj == hi_y && j > 0)) continue;
                pri = -(f[j] + 0.5 * chn_pen_gap * ((int32_t)a[j].x + (int32_t)a[j].y));
                if (best >= 0 && pri == best_pri) tied = 1; else if (best < 0 || pri < best_pri) tied = 0;
                if (best < 0 || pri < best_pri ||
                    (pri == best_pri && (yj > (int32_t)a[best].y || (yj == (int32_t)a[best].y && j > best))))
                    best = j, best_pri = pri;
            }
            if (best >= 0) {
                int32_t sc, exact, width, n_skip = 0;
                DEV_COUNT(6);
                if (tied) DEV_COUNT(1);
                j = best;
                sc = f[j] + comput_sc_simple(&a[i], &a[j], chn_pen_gap, chn_pen_skip, &exact, &width);
                if (width <= bw && sc > max_f) max_f = sc, max_j = j;
                if (!exact && max_dist_inner > 0 && (int32_t)a[i].y > 0) {
                    /* krmq_interval(root_inner, (y_i - 1, n)) -> largest key <= that; iterate downwards */
                    int64_t m = 0, c;
                    int32_t yi = (int32_t)a[i].y;
                    for (j = st_inner < i0 ? st_inner : i0; j < i0; ++j)
                        if ((int32_t)a[j].y <= yi - 1) ord[m++] = (int32_t)j;
                    /* sort by key (y, j) descending: insertion sort (m is small) */
                    for (c = 1; c < m; ++c) {
                        int32_t tmpj = ord[c];
                        int64_t d = c;
                        while (d > 0) {
                            int32_t o = ord[d - 1];
                            int32_t yo = (int32_t)a[o].y, yt = (int32_t)a[tmpj].y;
                            if (yo > yt || (yo == yt && o > tmpj)) break;
                            ord[d] = o; --d;
                        }
                        ord[d] = tmpj;
                    }
                    for (c = 0; c < m; ++c) {
                        int32_t width2;
                        j = ord[c];
                        if ((int32_t)a[j].y < yi - max_dist_inner) break;
                        sc = f[j] + comput_sc_simple(&a[i], &a[j], chn_pen_gap, chn_pen_skip, 0, &width2);
                        if (width2 <= bw) {
                            if (sc > max_f) {
                                max_f = sc, max_j = j;
                                if (n_skip > 0) --n_skip;
                            } else if (t[j] == (int32_t)i) {
                                if (++n_skip > max_chn_skip) break;
                            }
                            if (p[j] >= 0) t[p[j]] = (int32_t)i;
                        }
                    }
                }
            }
        }
        f[i] = max_f, p[i] = max_j;
        v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
    }
    free(ord);
    u = chain_backtrack(n, f, p, v, t, min_cnt, min_sc, max_drop, &n_u, &n_v);
    *n_u_ = n_u, *_u = u;
    free(p); free(f); free(t);
    if (n_u == 0) { free(a); free(v); free(u); *_u = 0; return 0; }
    a = compact_a(n_u, u, n_v, v, a);
    free(v);
    return a;
}

/* ------------------------------- hit.c ------------------------------- */
static inline uint64_t hash64(uint64_t key)
{
    key = ~key + (key << 21);
    key = key ^ key >> 24;
    key = (key + (key << 3)) + (key << 8);
    key = key ^ key >> 14;
    key = (key + (key << 2)) + (key << 4);
    key = key ^ key >> 28;
    key = key + (key << 31);
    return key;
}
static inline uint32_t wang_hash(uint32_t key)
{
    key += ~(key << 15);
    key ^= (key >> 10);
    key += (key << 3);
    key ^= (key >> 6);
    key += ~(key << 11);
    key ^= (key >> 16);
    return key;
}

static void reg_set_coor(reg_t *r, int32_t qlen, const orc128_t *a)
{
    int32_t k = r->as, q_span = (int32_t)(a[k].y >> 32 & 0xff);
    r->rev = (int32_t)(a[k].x >> 63);
    r->rid = (int32_t)(a[k].x << 1 >> 33);
    r->rs = (int32_t)a[k].x + 1 > q_span ? (int32_t)a[k].x + 1 - q_span : 0;
    r->re = (int32_t)a[k + r->cnt - 1].x + 1;
    if (!r->rev) {
        r->qs = (int32_t)a[k].y + 1 - q_span;
        r->qe = (int32_t)a[k + r->cnt - 1].y + 1;
    } else {
        r->qs = qlen - ((int32_t)a[k + r->cnt - 1].y + 1);
        r->qe = qlen - ((int32_t)a[k].y + 1 - q_span);
    }
}
static void cal_fuzzy_len(reg_t *r, const orc128_t *a)
{
    int i;
    r->mlen = r->blen = 0;
    if (r->cnt <= 0) return;
    r->mlen = r->blen = a[r->as].y >> 32 & 0xff;
    for (i = r->as + 1; i < r->as + r->cnt; ++i) {
        int span = a[i].y >> 32 & 0xff;
        int tl = (int32_t)a[i].x - (int32_t)a[i - 1].x;
        int ql = (int32_t)a[i].y - (int32_t)a[i - 1].y;
        r->blen += tl > ql ? tl : ql;
        r->mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
    }
}

static reg_t *gen_regs(uint32_t hash, int qlen, int n_u, uint64_t *u, orc128_t *a)
{
    orc128_t *z, tmp;
    reg_t *r;
    int i, k;
    if (n_u == 0) return 0;
    z = (orc128_t *)malloc((size_t)n_u * 16);
    for (i = k = 0; i < n_u; ++i) {
        uint32_t h = (uint32_t)hash64((hash64(a[k].x) + hash64(a[k].y)) ^ hash);
        z[i].x = u[i] ^ h;
        z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
        k += (int32_t)u[i];
    }
    orc_radix_sort_128x(z, z + n_u);
    for (i = 0; i < n_u >> 1; ++i) tmp = z[i], z[i] = z[n_u - 1 - i], z[n_u - 1 - i] = tmp;
    r = (reg_t *)calloc((size_t)n_u, sizeof(reg_t));
    for (i = 0; i < n_u; ++i) {
        reg_t *ri = &r[i];
        ri->id = i;
        ri->parent = PARENT_UNSET;
        ri->score = ri->score0 = (int32_t)(z[i].x >> 32);
        ri->hash = (uint32_t)z[i].x;
        ri->cnt = (int32_t)z[i].y;
        ri->as = (int32_t)(z[i].y >> 32);
        reg_set_coor(ri, qlen, a);
        cal_fuzzy_len(ri, a);
    }
    free(z);
    return r;
}

static void set_sam_pri(int n, reg_t *r)
{
    int i, n_pri = 0;
    for (i = 0; i < n; ++i)
        if (r[i].id == r[i].parent) {
            ++n_pri;
            r[i].sam_pri = (n_pri == 1);
        } else r[i].sam_pri = 0;
}

static void sync_regs(int n_regs, reg_t *regs)
{
    int *tmp, i, max_id = -1, n_tmp;
    if (n_regs <= 0) return;
    for (i = 0; i < n_regs; ++i) max_id = max_id > regs[i].id ? max_id : regs[i].id;
    n_tmp = max_id + 1;
    tmp = (int *)malloc((size_t)(n_tmp + 1) * sizeof(int));
    for (i = 0; i < n_tmp; ++i) tmp[i] = -1;
    for (i = 0; i < n_regs; ++i)
        if (regs[i].id >= 0) tmp[regs[i].id] = i;
    for (i = 0; i < n_regs; ++i) {
        reg_t *r = &regs[i];
        r->id = i;
        if (r->parent == PARENT_TMP_PRI) r->parent = i;
        else if (r->parent >= 0 && r->parent < n_tmp && tmp[r->parent] >= 0) r->parent = tmp[r->parent];
        else r->parent = PARENT_UNSET;
    }
    free(tmp);
    set_sam_pri(n_regs, regs);
}

static void set_parent(float mask_level, int mask_len, int n, reg_t *r, int sub_diff)
{
    int i, j, k, *w;
    uint64_t *cov;
    if (n <= 0) return;
    for (i = 0; i < n; ++i) r[i].id = i;
    cov = (uint64_t *)malloc((size_t)n * 8);
    w = (int *)malloc((size_t)n * sizeof(int));
    w[0] = 0, r[0].parent = 0;
    for (i = 1, k = 1; i < n; ++i) {
        reg_t *ri = &r[i];
        int si = ri->qs, ei = ri->qe, n_cov = 0, uncov_len = 0;
        for (j = 0; j < k; ++j) {
            reg_t *rp = &r[w[j]];
            int sj = rp->qs, ej = rp->qe;
            if (ej <= si || sj >= ei) continue;
            if (sj < si) sj = si;
            if (ej > ei) ej = ei;
            cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
        }
        if (n_cov == 0) {
            goto set_parent_test;
        } else {
            int jj, x = si;
            orc_radix_sort_64(cov, cov + n_cov);
            for (jj = 0; jj < n_cov; ++jj) {
                if ((int)(cov[jj] >> 32) > x) uncov_len += (int)(cov[jj] >> 32) - x;
                x = (int32_t)cov[jj] > x ? (int32_t)cov[jj] : x;
            }
            if (ei > x) uncov_len += ei - x;
        }
        for (j = 0; j < k; ++j) {
            reg_t *rp = &r[w[j]];
            int sj = rp->qs, ej = rp->qe, min, max, ol;
            if (ej <= si || sj >= ei) continue;
            min = ej - sj < ei - si ? ej - sj : ei - si;
            max = ej - sj > ei - si ? ej - sj : ei - si;
            ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
            if ((float)ol / min - (float)uncov_len / max > mask_level && uncov_len <= mask_len) {
                int cnt_sub = 0, sci = ri->score;
                ri->parent = rp->parent;
                rp->subsc = rp->subsc > sci ? rp->subsc : sci;
                if (ri->cnt >= rp->cnt) cnt_sub = 1;
                if (rp->has_p && ri->has_p && (rp->rid != ri->rid || rp->rs != ri->rs || rp->re != ri->re || ol != min)) {
                    sci = ri->dp_max;
                    rp->dp_max2 = rp->dp_max2 > sci ? rp->dp_max2 : sci;
                    if (rp->dp_max - ri->dp_max <= sub_diff) cnt_sub = 1;
                }
                if (cnt_sub) ++rp->n_sub;
                break;
            }
        }
set_parent_test:
        if (j == k) w[k++] = i, ri->parent = i, ri->n_sub = 0;
    }
    free(cov); free(w);
}

static void select_sub(float pri_ratio, int min_diff, int best_n, int check_strand, int min_strand_sc, int *n_, reg_t *r)
{
    if (pri_ratio > 0.0f && *n_ > 0) {
        int i, k, n = *n_, n_2nd = 0;
        for (i = k = 0; i < n; ++i) {
            int p = r[i].parent;
            if (p == i || r[i].inv) {
                r[k++] = r[i];
            } else if ((r[i].score >= r[p].score * pri_ratio || r[i].score + min_diff >= r[p].score) && n_2nd < best_n) {
                if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rid == r[p].rid && r[i].rs == r[p].rs && r[i].re == r[p].re))
                    r[k++] = r[i], ++n_2nd;
            } else if (check_strand && n_2nd < best_n && r[i].score > min_strand_sc && r[p].rev != r[i].rev) {
                r[i].strand_retained = 1;
                r[k++] = r[i], ++n_2nd;
            }
        }
        if (k != n) sync_regs(k, r);
        *n_ = k;
    }
}

/* [UP] hit.c mm_set_inv_mapq + mm_set_mapq (long reads: is_sr = 0).  logf is taken as the correctly rounded float
 * logarithm, (float)log((double)x), so that the CUDA path can reproduce it bit for bit. */
static inline float logf_cr(float x) { return (float)log((double)x); }
static void set_mapq(int n_regs, reg_t *regs, int min_chain_sc, int match_sc, int rep_len, int32_t *mapq)
{
    static const float q_coef = 40.0f;
    int64_t sum_sc = 0;
    float uniq_ratio;
    int i;
    if (n_regs == 0) return;
    for (i = 0; i < n_regs; ++i)
        if (regs[i].parent == regs[i].id) sum_sc += regs[i].score;
    uniq_ratio = (float)sum_sc / (sum_sc + rep_len);
    for (i = 0; i < n_regs; ++i) {
        reg_t *r = &regs[i];
        if (r->inv) mapq[i] = 0;
        else if (r->parent == r->id) {
            int mq, subsc;
            float pen_s1 = (r->score > 100 ? 1.0f : 0.01f * r->score) * uniq_ratio;
            float pen_cm = r->cnt > 10 ? 1.0f : 0.1f * r->cnt;
            pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
            subsc = r->subsc > min_chain_sc ? r->subsc : min_chain_sc;
            if (r->has_p && r->dp_max2 > 0 && r->dp_max > 0) {
                float identity = (float)r->mlen / r->blen;
                float x = (float)r->dp_max2 * subsc / r->dp_max / r->score0;
                int mq_alt;
                mq = (int)(identity * pen_cm * q_coef * (1.0f - x * x) * logf_cr((float)r->dp_max / match_sc));
                mq_alt = (int)(6.02f * identity * identity * (r->dp_max - r->dp_max2) / match_sc + .499f);
                mq = mq < mq_alt ? mq : mq_alt;
            } else {
                float x = (float)subsc / r->score0;
                if (r->has_p) {
                    float identity = (float)r->mlen / r->blen;
                    mq = (int)(identity * pen_cm * q_coef * (1.0f - x) * logf_cr((float)r->dp_max / match_sc));
                } else mq = (int)(pen_cm * q_coef * (1.0f - x) * logf_cr((float)r->score));
            }
            mq -= (int)(4.343f * logf_cr((float)(r->n_sub + 1)) + .499f);
            mq = mq > 0 ? mq : 0;
            mapq[i] = mq < 60 ? mq : 60;
            if (r->has_p && r->dp_max > r->dp_max2 && mapq[i] == 0) mapq[i] = 1;
        } else mapq[i] = 0;
    }
    /* mm_set_inv_mapq: an inversion piece takes the smaller MAPQ of its two neighbours on the target */
    if (n_regs >= 3) {
        for (i = 0; i < n_regs; ++i) if (regs[i].inv) break;
        if (i < n_regs) {
            orc128_t *aux = (orc128_t *)malloc((size_t)n_regs * 16);
            int n_aux = 0;
            for (i = 0; i < n_regs; ++i)
                if (regs[i].parent == i || regs[i].parent < 0) aux[n_aux].y = (uint64_t)i, aux[n_aux++].x = (uint64_t)regs[i].rid << 32 | (uint32_t)regs[i].rs;
            orc_radix_sort_128x(aux, aux + n_aux);
            for (i = 1; i < n_aux - 1; ++i)
                if (regs[aux[i].y].inv) {
                    int32_t l = mapq[aux[i - 1].y], rr = mapq[aux[i + 1].y];
                    mapq[aux[i].y] = l < rr ? l : rr;
                }
            free(aux);
        }
    }
}

static void filter_regs(const orc_opt_t *opt, int qlen, int *n_regs, reg_t *regs)
{
    int i, k;
    for (i = k = 0; i < *n_regs; ++i) {
        reg_t *r = &regs[i];
        int flt = 0;
        if (!r->inv && r->cnt < opt->min_cnt) flt = 1;
        if (r->has_p) {
            if (r->mlen < opt->min_chain_score) flt = 1;
            else if (r->dp_max < opt->min_dp_max) flt = 1;
            else if (r->qs > qlen * opt->max_clip_ratio && qlen - r->qe > qlen * opt->max_clip_ratio) flt = 1;
        }
        if (!flt) {
            if (k < i) regs[k++] = regs[i];
            else ++k;
        }
    }
    *n_regs = k;
}

static void hit_sort(int *n_regs, reg_t *r)
{
    int32_t i, n_aux, n = *n_regs;
    orc128_t *aux;
    reg_t *t;
    if (n <= 1) return;
    aux = (orc128_t *)malloc((size_t)n * 16);
    t = (reg_t *)malloc((size_t)n * sizeof(reg_t));
    for (i = n_aux = 0; i < n; ++i) {
        if (r[i].inv || r[i].cnt > 0) {
            int score = r[i].has_p ? r[i].dp_max : r[i].score;
            aux[n_aux].x = (uint64_t)(uint32_t)score << 32 | r[i].hash;
            aux[n_aux++].y = (uint64_t)i;
        }
    }
    orc_radix_sort_128x(aux, aux + n_aux);
    for (i = n_aux - 1; i >= 0; --i) t[n_aux - 1 - i] = r[aux[i].y];
    memcpy(r, t, sizeof(reg_t) * (size_t)n_aux);
    *n_regs = n_aux;
    free(aux); free(t);
}

static int squeeze_a(int n_regs, reg_t *regs, orc128_t *a)
{
    int i, as = 0;
    uint64_t *aux = (uint64_t *)malloc((size_t)(n_regs + 1) * 8);
    for (i = 0; i < n_regs; ++i) aux[i] = (uint64_t)regs[i].as << 32 | (uint32_t)i;
    orc_radix_sort_64(aux, aux + n_regs);
    for (i = 0; i < n_regs; ++i) {
        reg_t *r = &regs[(int32_t)aux[i]];
        if (r->as != as) {
            memmove(&a[as], &a[r->as], (size_t)r->cnt * 16);
            r->as = as;
        }
        as += r->cnt;
    }
    free(aux);
    return as;
}

static void split_reg(reg_t *r, reg_t *r2, int n, int qlen, orc128_t *a)
{
    if (n <= 0 || n >= r->cnt) return;
    *r2 = *r;
    r2->id = -1;
    r2->sam_pri = 0;
    r2->has_p = 0, r2->cigar = 0, r2->n_cigar = r2->m_cigar = 0, r2->dp_score = r2->dp_max = r2->dp_max2 = r2->n_ambi = 0;
    r2->split_inv = 0;
    r2->cnt = r->cnt - n;
    r2->score = (int32_t)(r->score * ((float)r2->cnt / r->cnt) + .499);
    r2->as = r->as + n;
    if (r->parent == r->id) r2->parent = PARENT_TMP_PRI;
    reg_set_coor(r2, qlen, a);
    r->cnt -= r2->cnt;
    r->score -= r2->score;
    reg_set_coor(r, qlen, a);
    r->split |= 1, r2->split |= 2;
}

/* [UP] align.c mm_recal_max_dp / mm_update_dp_max */
static void update_dp_max(int qlen, int n_regs, reg_t *regs, float frac, int a, int b)
{
    int32_t max = -1, max2 = -1, i, max_i = -1;
    double div, b2;
    if (n_regs < 2) return;
    for (i = 0; i < n_regs; ++i) {
        reg_t *r = &regs[i];
        if (!r->has_p) continue;
        if (r->dp_max > max) max2 = max, max = r->dp_max, max_i = i;
        else if (r->dp_max > max2) max2 = r->dp_max;
    }
    if (max_i < 0 || max < 0 || max2 < 0) return;
    if (regs[max_i].qe - regs[max_i].qs < (double)qlen * frac) return;
    if (max2 < (double)max * frac) return;
    div = 1. - (double)regs[max_i].mlen / regs[max_i].blen;
    if (div < 0.02) div = 0.02;
    b2 = 0.5 / div;
    if (b2 * a < b) b2 = (double)a / b;
    for (i = 0; i < n_regs; ++i) {
        reg_t *r = &regs[i];
        uint32_t k;
        int32_t n_gap = 0, n_mis;
        double gap_cost = 0.0;
        if (!r->has_p) continue;
        for (k = 0; k < (uint32_t)r->n_cigar; ++k) {
            int32_t op = r->cigar[k] & 0xf, len = r->cigar[k] >> 4;
            if (op == 1 || op == 2) {
                gap_cost += b2 + (double)mg_log2(1.0f + (float)len);
                n_gap += len;
            }
        }
        n_mis = r->blen + r->n_ambi - r->mlen - n_gap;
        r->dp_max = (int32_t)(a * (r->mlen - b2 * n_mis - gap_cost) + .499);
        if (r->dp_max < 0) r->dp_max = 0;
    }
}

/* ------------------------------- align.c ------------------------------- */
typedef struct {
    const orc_opt_t *opt;
    const uint8_t *tseq0; int32_t tlen0;
    const uint8_t *qseq0[2]; int32_t qlen;
    arena_t *A;
    int64_t cells, n_tasks;
    orc_ez_t ez;
} actx_t;

static void append_cigar(actx_t *c, reg_t *r, uint32_t n_cigar, const uint32_t *cigar)
{
    if (n_cigar == 0) return;
    if (!r->has_p) {
        r->has_p = 1; r->dp_score = r->dp_max = r->dp_max2 = r->n_ambi = 0;
        r->n_cigar = 0; r->m_cigar = 0; r->cigar = 0;
    }
    if (r->n_cigar + (int32_t)n_cigar > r->m_cigar) {
        int32_t m = (r->n_cigar + (int32_t)n_cigar) * 2 + 8;
        uint32_t *nc = (uint32_t *)aalloc(c->A, (size_t)m * 4);
        if (r->n_cigar) memcpy(nc, r->cigar, (size_t)r->n_cigar * 4);
        r->cigar = nc; r->m_cigar = m;
    }
    if (r->n_cigar > 0 && (r->cigar[r->n_cigar - 1] & 0xf) == (cigar[0] & 0xf)) {
        r->cigar[r->n_cigar - 1] += (cigar[0] >> 4) << 4;
        if (n_cigar > 1) memcpy(r->cigar + r->n_cigar, cigar + 1, (size_t)(n_cigar - 1) * 4);
        r->n_cigar += (int32_t)n_cigar - 1;
    } else {
        memcpy(r->cigar + r->n_cigar, cigar, (size_t)n_cigar * 4);
        r->n_cigar += (int32_t)n_cigar;
    }
}

int64_t orc_cell_stats[8];   /* debugging: cells by task class */
int64_t orc_size_hist[4][16];
int64_t orc_tlen_hist[64], orc_qlen_hist[64];
int64_t orc_fill_model[4];   /* debugging: fill cells; cell slots of the 8-column systolic mapping; of a flexible-width mapping */
static void align_pair(actx_t *c, int qlen, const uint8_t *qseq, int tlen, const uint8_t *tseq, int w, int end_bonus,
                       int zdrop, int flag)
{
    const orc_opt_t *o = c->opt;
    orc_ez_t *ez = &c->ez;
    if (o->max_sw_mat > 0 && (int64_t)tlen * qlen > o->max_sw_mat) {
        ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
        ez->max = 0; ez->score = ez->mqe = ez->mte = -0x40000000;
        ez->n_cigar = 0; ez->reach_end = 0;
        ez->zdropped = 1;
        return;
    }
    orc_ksw_extd2(qlen, qseq, tlen, tseq, o->a, o->b, o->sc_ambi, o->q, o->e, o->q2, o->e2, w, zdrop, end_bonus, flag, ez);
    c->cells += ez->cells;
    c->n_tasks++;
    {
        int cls = (flag & ORC_KSW_APPROX_MAX) ? 0 : (flag & ORC_KSW_EXTZ_ONLY) ? 2 : 1, b = 0;
        int mx = qlen > tlen ? qlen : tlen;
        while ((1 << (b + 5)) < mx && b < 15) ++b;
#pragma omp atomic
        orc_cell_stats[cls] += ez->cells;
#pragma omp atomic
        orc_cell_stats[4 + cls] += 1;
#pragma omp atomic
        orc_size_hist[cls][b] += ez->cells;
        if (cls == 0) {
            int tb = tlen / 32 < 63 ? tlen / 32 : 63, qb = qlen / 32 < 63 ? qlen / 32 : 63;
#pragma omp atomic
            orc_tlen_hist[tb] += ez->cells;
#pragma omp atomic
            orc_qlen_hist[qb] += ez->cells;
            {
                int64_t npairs = (qlen + 1) / 2, stepsA = 0, slotsB;
                for (int t0 = 0; t0 < tlen; t0 += 256) {
                    int rem = tlen - t0, nlive = rem >= 256 ? 32 : (rem + 7) / 8;
                    stepsA += npairs + nlive - 1;
                }
                int fc = (tlen + 31) / 32; fc += fc & 1; if (fc < 2) fc = 2;
                if (fc <= 16) { int nlive = (tlen + fc - 1) / fc; slotsB = (npairs + nlive - 1) * 32 * (fc + 1) * 2; }
                else { int np = (tlen + 511) / 512; slotsB = (int64_t)np * (npairs + 31) * 32 * 17 * 2; }
#pragma omp atomic
                orc_fill_model[0] += ez->cells;
#pragma omp atomic
                orc_fill_model[1] += stepsA * 32 * 18;
#pragma omp atomic
                orc_fill_model[2] += slotsB;
            }
        }
    }
}

static inline void update_max_zdrop(int32_t score, int i, int j, int32_t *max, int *max_i, int *max_j, int e,
                                    int *max_zdrop, int pos[2][2])
{
    if (score < *max) {
        int li = i - *max_i, lj = j - *max_j;
        int diff = li > lj ? li - lj : lj - li;
        int z = *max - score - diff * e;
        if (z > *max_zdrop) {
            *max_zdrop = z;
            pos[0][0] = *max_i, pos[0][1] = *max_j;
            pos[1][0] = i, pos[1][1] = j;
        }
    } else *max = score, *max_i = i, *max_j = j;
}

static int test_zdrop(actx_t *c, const uint8_t *qseq, const uint8_t *tseq, uint32_t n_cigar, const uint32_t *cigar)
{
    const orc_opt_t *o = c->opt;
    uint32_t k;
    int32_t score = 0, max = INT32_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
    int pos[2][2] = {{-1, -1}, {-1, -1}}, q_len, t_len;
    for (k = 0, score = 0; k < n_cigar; ++k) {
        uint32_t l, op = cigar[k] & 0xf, len = cigar[k] >> 4;
        if (op == 0) {
            for (l = 0; l < len; ++l) {
                int tc = tseq[i + l], qc = qseq[j + l];
                score += (tc > 3 || qc > 3) ? -o->sc_ambi : tc == qc ? o->a : -o->b;
                update_max_zdrop(score, i + l, j + l, &max, &max_i, &max_j, o->e, &max_zdrop, pos);
            }
            i += len, j += len;
        } else if (op == 1 || op == 2) {
            score -= o->q + o->e * len;
            if (op == 1) j += len;
            else i += len;
            update_max_zdrop(score, i, j, &max, &max_i, &max_j, o->e, &max_zdrop, pos);
        }
    }
    q_len = pos[1][1] - pos[0][1], t_len = pos[1][0] - pos[0][0];
    if (max_zdrop > o->zdrop_inv && q_len < o->max_gap && t_len < o->max_gap) {
        uint8_t *qseq2 = (uint8_t *)malloc((size_t)(q_len > 0 ? q_len : 1));
        int qe, te;
        for (i = 0; i < q_len; ++i) {
            int cc = qseq[pos[1][1] - i - 1];
            qseq2[i] = cc >= 4 ? 4 : 3 - cc;
        }
        score = orc_ksw_ll(q_len, qseq2, t_len, tseq + pos[0][0], o->a, o->b, o->sc_ambi, o->q, o->e, &qe, &te);
        c->cells += (int64_t)(q_len > 0 ? q_len : 0) * (t_len > 0 ? t_len : 0);
        free(qseq2);
        if (score >= o->min_chain_score * o->a && score >= o->min_dp_max) return 2;
    }
    return max_zdrop > o->zdrop ? 1 : 0;
}

static void fix_cigar(reg_t *r, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift)
{
    int32_t toff = 0, qoff = 0, to_shrink = 0;
    int32_t k;
    *qshift = *tshift = 0;
    if (r->n_cigar <= 1) return;
    for (k = 0; k < r->n_cigar; ++k) {
        uint32_t op = r->cigar[k] & 0xf, len = r->cigar[k] >> 4;
        if (len == 0) to_shrink = 1;
        if (op == 0) {
            toff += len, qoff += len;
        } else if (op == 1 || op == 2) {
            if (k > 0 && k < r->n_cigar - 1 && (r->cigar[k - 1] & 0xf) == 0 && (r->cigar[k + 1] & 0xf) == 0) {
                int l, prev_len = r->cigar[k - 1] >> 4;
                if (op == 1) {
                    for (l = 0; l < prev_len; ++l)
                        if (qseq[qoff - 1 - l] != qseq[qoff + len - 1 - l]) break;
                } else {
                    for (l = 0; l < prev_len; ++l)
                        if (tseq[toff - 1 - l] != tseq[toff + len - 1 - l]) break;
                }
                if (l > 0) r->cigar[k - 1] -= (uint32_t)l << 4, r->cigar[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
                if (l == prev_len) to_shrink = 1;
            }
            if (op == 1) qoff += len;
            else toff += len;
        }
    }
    for (k = 0; k < r->n_cigar - 2; ++k) {
        if ((r->cigar[k] & 0xf) > 0 && (r->cigar[k] & 0xf) + (r->cigar[k + 1] & 0xf) == 3) {
            int32_t l;
            uint32_t s[3] = {0, 0, 0};
            for (l = k; l < r->n_cigar; ++l) {
                uint32_t op = r->cigar[l] & 0xf;
                if (op == 1 || op == 2 || r->cigar[l] >> 4 == 0) s[op] += r->cigar[l] >> 4;
                else break;
            }
            if (s[1] > 0 && s[2] > 0 && l - k > 2) {
                r->cigar[k] = s[1] << 4 | 1;
                r->cigar[k + 1] = s[2] << 4 | 2;
                for (k += 2; k < l; ++k) r->cigar[k] &= 0xf;
                to_shrink = 1;
            }
            k = l;
        }
    }
    if (to_shrink) {
        int32_t l = 0;
        for (k = 0; k < r->n_cigar; ++k)
            if (r->cigar[k] >> 4 != 0) r->cigar[l++] = r->cigar[k];
        r->n_cigar = l;
        for (k = l = 0; k < r->n_cigar; ++k)
            if (k == r->n_cigar - 1 || (r->cigar[k] & 0xf) != (r->cigar[k + 1] & 0xf)) r->cigar[l++] = r->cigar[k];
            else r->cigar[k + 1] += r->cigar[k] >> 4 << 4;
        r->n_cigar = l;
    }
    if ((r->cigar[0] & 0xf) == 1 || (r->cigar[0] & 0xf) == 2) {
        int32_t l = r->cigar[0] >> 4;
        if ((r->cigar[0] & 0xf) == 1) {
            if (r->rev) r->qe -= l;
            else r->qs += l;
            *qshift = l;
        } else r->rs += l, *tshift = l;
        --r->n_cigar;
        memmove(r->cigar, r->cigar + 1, (size_t)r->n_cigar * 4);
    }
}

static void update_extra(const orc_opt_t *o, reg_t *r, const uint8_t *qseq, const uint8_t *tseq)
{
    uint32_t k, l;
    int32_t qshift, tshift, toff = 0, qoff = 0;
    double s = 0.0, max = 0.0;
    if (!r->has_p) return;
    fix_cigar(r, qseq, tseq, &qshift, &tshift);
    qseq += qshift, tseq += tshift;
    r->blen = r->mlen = 0;
    for (k = 0; k < (uint32_t)r->n_cigar; ++k) {
        uint32_t op = r->cigar[k] & 0xf, len = r->cigar[k] >> 4;
        if (op == 0) {
            int n_ambi = 0, diff = 0;
            for (l = 0; l < len; ++l) {
                int cq = qseq[qoff + l], ct = tseq[toff + l];
                if (ct > 3 || cq > 3) ++n_ambi;
                else if (ct != cq) ++diff;
                s += (ct > 3 || cq > 3) ? -o->sc_ambi : ct == cq ? o->a : -o->b;
                if (s < 0) s = 0;
                else max = max > s ? max : s;
            }
            r->blen += len - n_ambi, r->mlen += len - (n_ambi + diff), r->n_ambi += n_ambi;
            toff += len, qoff += len;
        } else if (op == 1) {
            int n_ambi = 0;
            for (l = 0; l < len; ++l)
                if (qseq[qoff + l] > 3) ++n_ambi;
            r->blen += len - n_ambi, r->n_ambi += n_ambi;
            s -= o->q + o->e * len;
            if (s < 0) s = 0;
            qoff += len;
        } else if (op == 2) {
            int n_ambi = 0;
            for (l = 0; l < len; ++l)
                if (tseq[toff + l] > 3) ++n_ambi;
            r->blen += len - n_ambi, r->n_ambi += n_ambi;
            s -= o->q + o->e * len;
            if (s < 0) s = 0;
            toff += len;
        }
    }
    r->dp_max = (int32_t)(max + .499);
}

static void fix_bad_ends(const reg_t *r, const orc128_t *a, int bw, int min_match, int32_t *as, int32_t *cnt)
{
    int32_t i, l, m;
    *as = r->as, *cnt = r->cnt;
    if (r->cnt < 3) return;
    m = l = a[r->as].y >> 32 & 0xff;
    for (i = r->as + 1; i < r->as + r->cnt - 1; ++i) {
        int32_t lq, lr, min, max;
        int32_t q_span = a[i].y >> 32 & 0xff;
        if (a[i].y & SEED_LONG_JOIN) break;
        lr = (int32_t)a[i].x - (int32_t)a[i - 1].x;
        lq = (int32_t)a[i].y - (int32_t)a[i - 1].y;
        min = lr < lq ? lr : lq;
        max = lr > lq ? lr : lq;
        if (max - min > l >> 1) *as = i;
        l += min;
        m += min < q_span ? min : q_span;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
    }
    *cnt = r->as + r->cnt - *as;
    m = l = a[r->as + r->cnt - 1].y >> 32 & 0xff;
    for (i = r->as + r->cnt - 2; i > *as; --i) {
        int32_t lq, lr, min, max;
        int32_t q_span = a[i + 1].y >> 32 & 0xff;
        if (a[i + 1].y & SEED_LONG_JOIN) break;
        lr = (int32_t)a[i + 1].x - (int32_t)a[i].x;
        lq = (int32_t)a[i + 1].y - (int32_t)a[i].y;
        min = lr < lq ? lr : lq;
        max = lr > lq ? lr : lq;
        if (max - min > l >> 1) *cnt = i + 1 - *as;
        l += min;
        m += min < q_span ? min : q_span;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
    }
}

static int *collect_long_gaps(int as1, int cnt1, orc128_t *a, int min_gap, int *n_)
{
    int i, n, *K;
    *n_ = 0;
    for (i = 1, n = 0; i < cnt1; ++i) {
        int gap = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - (int32_t)a[as1 + i - 1].x);
        if (gap < -min_gap || gap > min_gap) ++n;
    }
    if (n <= 1) return 0;
    K = (int *)malloc((size_t)n * sizeof(int));
    for (i = 1, n = 0; i < cnt1; ++i) {
        int gap = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - (int32_t)a[as1 + i - 1].x);
        if (gap < -min_gap || gap > min_gap) K[n++] = i;
    }
    *n_ = n;
    return K;
}

static void filter_bad_seeds(int as1, int cnt1, orc128_t *a, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt)
{
    int max_st, max_en, n, i, k, max, *K;
    K = collect_long_gaps(as1, cnt1, a, min_gap, &n);
    if (K == 0) return;
    max = 0, max_st = max_en = -1;
    for (k = 0;; ++k) {
        int gap, l, n_ins = 0, n_del = 0, qs, rs, max_diff = 0, max_diff_l = -1;
        if (k == n || k >= max_en) {
            if (max_en > 0)
                for (i = K[max_st]; i < K[max_en]; ++i) a[as1 + i].y |= SEED_IGNORE;
            max = 0, max_st = max_en = -1;
            if (k == n) break;
        }
        i = K[k];
        gap = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - (int32_t)a[as1 + i - 1].x);
        if (gap > 0) n_ins += gap;
        else n_del += -gap;
        qs = (int32_t)a[as1 + i - 1].y;
        rs = (int32_t)a[as1 + i - 1].x;
        for (l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
            int j = K[l], diff;
            if ((int32_t)a[as1 + j].y - qs > max_ext_len || (int32_t)a[as1 + j].x - rs > max_ext_len) break;
            gap = ((int32_t)a[as1 + j].y - (int32_t)a[as1 + j - 1].y) - ((int32_t)a[as1 + j].x - (int32_t)a[as1 + j - 1].x);
            if (gap > 0) n_ins += gap;
            else n_del += -gap;
            diff = n_ins + n_del - abs(n_ins - n_del);
            if (max_diff < diff) max_diff = diff, max_diff_l = l;
        }
        if (max_diff > diff_thres && max_diff > max) max = max_diff, max_st = k, max_en = max_diff_l;
    }
    free(K);
}

static void filter_bad_seeds_alt(int as1, int cnt1, orc128_t *a, int min_gap, int max_ext)
{
    int n, k, *K;
    K = collect_long_gaps(as1, cnt1, a, min_gap, &n);
    if (K == 0) return;
    for (k = 0; k < n;) {
        int i = K[k], l;
        int gap1 = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - (int32_t)a[as1 + i - 1].x);
        int re1 = (int32_t)a[as1 + i].x;
        int qe1 = (int32_t)a[as1 + i].y;
        gap1 = gap1 > 0 ? gap1 : -gap1;
        for (l = k + 1; l < n; ++l) {
            int j = K[l], gap2, q_span_pre, rs2, qs2, m;
            if ((int32_t)a[as1 + j].y - qe1 > max_ext || (int32_t)a[as1 + j].x - re1 > max_ext) break;
            gap2 = ((int32_t)a[as1 + j].y - (int32_t)a[as1 + j - 1].y) - ((int32_t)a[as1 + j].x - (int32_t)a[as1 + j - 1].x);
            q_span_pre = a[as1 + j - 1].y >> 32 & 0xff;
            rs2 = (int32_t)a[as1 + j - 1].x + q_span_pre;
            qs2 = (int32_t)a[as1 + j - 1].y + q_span_pre;
            m = rs2 - re1 < qs2 - qe1 ? rs2 - re1 : qs2 - qe1;
            gap2 = gap2 > 0 ? gap2 : -gap2;
            if (m > gap1 + gap2) break;
            re1 = (int32_t)a[as1 + j].x;
            qe1 = (int32_t)a[as1 + j].y;
            gap1 = gap2;
        }
        if (l > k + 1) {
            int j, end = K[l - 1];
            for (j = K[k]; j < end; ++j) a[as1 + j].y |= SEED_IGNORE;
            a[as1 + end].y |= SEED_LONG_JOIN;
        }
        k = l;
    }
    free(K);
}

static inline void adjust_minier(actx_t *c, const orc128_t *a, int32_t *r, int32_t *q)
{
    if (c->opt->hpc) {
        const uint8_t *qseq = c->qseq0[a->x >> 63];
        int i, cc;
        *q = (int32_t)a->y;
        for (i = *q - 1, cc = qseq[*q]; i > 0; --i)
            if (qseq[i] != cc) break;
        *q = i + 1;
        *r = (int32_t)a->x;
        for (i = *r - 1, cc = c->tseq0[*r]; i > 0; --i)   /* mm_get_hplen_back restated with the same i>0 bound */
            if (c->tseq0[i] != cc) break;
        *r = i + 1;
    } else {
        *r = (int32_t)a->x - (c->opt->k >> 1);
        *q = (int32_t)a->y - (c->opt->k >> 1);
    }
}

static void seq_rev_copy(uint8_t *dst, const uint8_t *src, int len)
{
    int i;
    for (i = 0; i < len; ++i) dst[i] = src[len - 1 - i];
}

static void align1(actx_t *c, reg_t *r, reg_t *r2, int n_a, orc128_t *a)
{
    const orc_opt_t *opt = c->opt;
    orc_ez_t *ez = &c->ez;
    int qlen = c->qlen;
    int32_t rev = (int32_t)(a[r->as].x >> 63), as1, cnt1;
    const uint8_t *qseq;
    uint8_t *tbuf, *qbuf;
    int32_t i, l, bw, bw_long, dropped = 0, rs0, re0, qs0, qe0;
    int32_t rs, re, qs, qe;
    int32_t rs1, qs1, re1, qe1;
    int32_t tl = c->tlen0;

    r2->cnt = 0;
    if (r->cnt == 0) return;
    bw = (int)(opt->bw * 1.5 + 1.);
    bw_long = (int)(opt->bw_long * 1.5 + 1.);
    if (bw_long < bw) bw_long = bw;

    fix_bad_ends(r, a, opt->bw, opt->min_chain_score * 2, &as1, &cnt1);
    filter_bad_seeds(as1, cnt1, a, 10, 40, opt->max_gap >> 1, 10);
    filter_bad_seeds_alt(as1, cnt1, a, 30, opt->max_gap >> 1);
    adjust_minier(c, &a[as1], &rs, &qs);
    adjust_minier(c, &a[as1 + cnt1 - 1], &re, &qe);
    assert(cnt1 > 0);

    rs0 = (int32_t)a[r->as].x + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
    qs0 = (int32_t)a[r->as].y + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
    if (rs0 < 0) rs0 = 0;
    rs1 = qs1 = 0;
    for (i = r->as - 1, l = 0; i >= 0 && a[i].x >> 32 == a[r->as].x >> 32; --i) {
        int32_t x = (int32_t)a[i].x + 1 - (int32_t)(a[i].y >> 32 & 0xff);
        int32_t y = (int32_t)a[i].y + 1 - (int32_t)(a[i].y >> 32 & 0xff);
        if (x < rs0 && y < qs0) {
            if (++l > opt->min_cnt) {
                l = rs0 - x > qs0 - y ? rs0 - x : qs0 - y;
                rs1 = rs0 - l, qs1 = qs0 - l;
                if (rs1 < 0) rs1 = 0;
                break;
            }
        }
    }
    if (qs > 0 && rs > 0) {
        l = qs < opt->max_gap ? qs : opt->max_gap;
        qs1 = qs1 > qs - l ? qs1 : qs - l;
        qs0 = qs0 < qs1 ? qs0 : qs1;
        l += l * opt->a > opt->q ? (l * opt->a - opt->q) / opt->e : 0;
        l = l < opt->max_gap ? l : opt->max_gap;
        l = l < rs ? l : rs;
        rs1 = rs1 > rs - l ? rs1 : rs - l;
        rs0 = rs0 < rs1 ? rs0 : rs1;
        rs0 = rs0 < rs ? rs0 : rs;
    } else rs0 = rs, qs0 = qs;
    re0 = (int32_t)a[r->as + r->cnt - 1].x + 1;
    qe0 = (int32_t)a[r->as + r->cnt - 1].y + 1;
    re1 = tl, qe1 = qlen;
    for (i = r->as + r->cnt, l = 0; i < n_a && a[i].x >> 32 == a[r->as].x >> 32; ++i) {
        int32_t x = (int32_t)a[i].x + 1;
        int32_t y = (int32_t)a[i].y + 1;
        if (x > re0 && y > qe0) {
            if (++l > opt->min_cnt) {
                l = x - re0 > y - qe0 ? x - re0 : y - qe0;
                re1 = re0 + l, qe1 = qe0 + l;
                break;
            }
        }
    }
    if (qe < qlen && re < tl) {
        l = qlen - qe < opt->max_gap ? qlen - qe : opt->max_gap;
        qe1 = qe1 < qe + l ? qe1 : qe + l;
        qe0 = qe0 > qe1 ? qe0 : qe1;
        l += l * opt->a > opt->q ? (l * opt->a - opt->q) / opt->e : 0;
        l = l < opt->max_gap ? l : opt->max_gap;
        l = l < tl - re ? l : tl - re;
        re1 = re1 < re + l ? re1 : re + l;
        re0 = re0 > re1 ? re0 : re1;
    } else re0 = re, qe0 = qe;

    assert(re0 > rs0);
    tbuf = (uint8_t *)malloc((size_t)(re0 - rs0) + 1);
    qbuf = (uint8_t *)malloc((size_t)(qe0 > qs0 ? qe0 - qs0 : 1) + 1);

    if (qs > 0 && rs > 0) {     /* left extension */
        seq_rev_copy(qbuf, &c->qseq0[rev][qs0], qs - qs0);
        seq_rev_copy(tbuf, &c->tseq0[rs0], rs - rs0);
        align_pair(c, qs - qs0, qbuf, rs - rs0, tbuf, bw, opt->end_bonus, r->split_inv ? opt->zdrop_inv : opt->zdrop,
                   ORC_KSW_EXTZ_ONLY | ORC_KSW_RIGHT | ORC_KSW_REV_CIGAR);
        if (ez->n_cigar > 0) {
            append_cigar(c, r, (uint32_t)ez->n_cigar, ez->cigar);
            r->dp_score += ez->max;
        }
        rs1 = rs - (ez->reach_end ? ez->mqe_t + 1 : ez->max_t + 1);
        qs1 = qs - (ez->reach_end ? qs - qs0 : ez->max_q + 1);
    } else rs1 = rs, qs1 = qs;
    re1 = rs, qe1 = qs;
    assert(qs1 >= 0 && rs1 >= 0);

    for (i = 1; i < cnt1; ++i) {    /* gap filling */
        if ((a[as1 + i].y & (SEED_IGNORE | SEED_TANDEM)) && i != cnt1 - 1) continue;
        adjust_minier(c, &a[as1 + i], &re, &qe);
        re1 = re, qe1 = qe;
        if (i == cnt1 - 1 || (a[as1 + i].y & SEED_LONG_JOIN) || (qe - qs >= opt->min_ksw_len && re - rs >= opt->min_ksw_len)) {
            int j, bw1 = bw_long, zdrop_code;
            if (a[as1 + i].y & SEED_LONG_JOIN) bw1 = qe - qs > re - rs ? qe - qs : re - rs;
            qseq = &c->qseq0[rev][qs];
            align_pair(c, qe - qs, qseq, re - rs, &c->tseq0[rs], bw1, -1, opt->zdrop, ORC_KSW_APPROX_MAX);
            if ((zdrop_code = test_zdrop(c, qseq, &c->tseq0[rs], (uint32_t)ez->n_cigar, ez->cigar)) != 0)
                align_pair(c, qe - qs, qseq, re - rs, &c->tseq0[rs], bw1, -1, zdrop_code == 2 ? opt->zdrop_inv : opt->zdrop, 0);
            if (ez->n_cigar > 0) append_cigar(c, r, (uint32_t)ez->n_cigar, ez->cigar);
            if (ez->zdropped) {
                if (!r->has_p) { r->has_p = 1; r->dp_score = r->dp_max = r->dp_max2 = r->n_ambi = 0; r->n_cigar = r->m_cigar = 0; r->cigar = 0; }
                for (j = i - 1; j >= 0; --j)
                    if ((int32_t)a[as1 + j].x <= rs + ez->max_t) break;
                dropped = 1;
                if (j < 0) j = 0;
                r->dp_score += ez->max;
                re1 = rs + (ez->max_t + 1);
                qe1 = qs + (ez->max_q + 1);
                if (cnt1 - (j + 1) >= opt->min_cnt) {
                    split_reg(r, r2, as1 + j + 1 - r->as, qlen, a);
                    if (zdrop_code == 2) r2->split_inv = 1;
                }
                break;
            } else r->dp_score += ez->score;
            rs = re, qs = qe;
        }
    }

    if (!dropped && qe < qe0 && re < re0) {     /* right extension */
        align_pair(c, qe0 - qe, &c->qseq0[rev][qe], re0 - re, &c->tseq0[re], bw, opt->end_bonus, opt->zdrop, ORC_KSW_EXTZ_ONLY);
        if (ez->n_cigar > 0) {
            append_cigar(c, r, (uint32_t)ez->n_cigar, ez->cigar);
            r->dp_score += ez->max;
        }
        re1 = re + (ez->reach_end ? ez->mqe_t + 1 : ez->max_t + 1);
        qe1 = qe + (ez->reach_end ? qe0 - qe : ez->max_q + 1);
    }
    assert(qe1 <= qlen);

    r->rs = rs1, r->re = re1;
    if (rev) r->qs = qlen - qe1, r->qe = qlen - qs1;
    else r->qs = qs1, r->qe = qe1;

    assert(re1 - rs1 <= re0 - rs0);
    if (r->has_p) update_extra(opt, r, &c->qseq0[r->rev][qs1], &c->tseq0[rs1]);
    free(tbuf); free(qbuf);
}

/* [UP] align.c mm_align1_inv */
static int align1_inv(actx_t *c, const reg_t *r1, const reg_t *r2, reg_t *r_inv)
{
    const orc_opt_t *opt = c->opt;
    orc_ez_t *ez = &c->ez;
    int qlen = c->qlen;
    int tl, ql, score, ret = 0, q_off, t_off;
    uint8_t *tseq, *qrev, *trev;
    const uint8_t *qseq;
    memset(r_inv, 0, sizeof(reg_t));
    if (!(r1->split & 1) || !(r2->split & 2)) return 0;
    if (r1->id != r1->parent && r1->parent != PARENT_TMP_PRI) return 0;
    if (r2->id != r2->parent && r2->parent != PARENT_TMP_PRI) return 0;
    if (r1->rid != r2->rid || r1->rev != r2->rev) return 0;
    ql = r1->rev ? r1->qs - r2->qe : r2->qs - r1->qe;
    tl = r2->rs - r1->re;
    if (ql < opt->min_chain_score || ql > opt->max_gap) return 0;
    if (tl < opt->min_chain_score || tl > opt->max_gap) return 0;
    tseq = (uint8_t *)malloc((size_t)tl);
    memcpy(tseq, &c->tseq0[r1->re], (size_t)tl);
    qseq = r1->rev ? &c->qseq0[0][r2->qe] : &c->qseq0[1][qlen - r2->qs];
    qrev = (uint8_t *)malloc((size_t)ql);
    trev = (uint8_t *)malloc((size_t)tl);
    seq_rev_copy(qrev, qseq, ql);
    seq_rev_copy(trev, tseq, tl);
    score = orc_ksw_ll(ql, qrev, tl, trev, opt->a, opt->b, opt->sc_ambi, opt->q, opt->e, &q_off, &t_off);
    c->cells += (int64_t)ql * tl;
    free(qrev); free(trev);
    if (score < opt->min_dp_max) goto end_align1_inv;
    q_off = ql - (q_off + 1), t_off = tl - (t_off + 1);
    align_pair(c, ql - q_off, qseq + q_off, tl - t_off, tseq + t_off, (int)(opt->bw * 1.5), -1, opt->zdrop, ORC_KSW_EXTZ_ONLY);
    if (ez->n_cigar == 0) goto end_align1_inv;
    append_cigar(c, r_inv, (uint32_t)ez->n_cigar, ez->cigar);
    r_inv->dp_score = ez->max;
    r_inv->id = -1;
    r_inv->parent = PARENT_UNSET;
    r_inv->inv = 1;
    r_inv->rev = !r1->rev;
    r_inv->rid = r1->rid;
    if (r_inv->rev == 0) {
        r_inv->qs = r2->qe + q_off;
        r_inv->qe = r_inv->qs + ez->max_q + 1;
    } else {
        r_inv->qe = r2->qs - q_off;
        r_inv->qs = r_inv->qe - (ez->max_q + 1);
    }
    r_inv->rs = r1->re + t_off;
    r_inv->re = r_inv->rs + ez->max_t + 1;
    update_extra(opt, r_inv, &qseq[q_off], &tseq[t_off]);
    ret = 1;
end_align1_inv:
    free(tseq);
    return ret;
}

static reg_t *insert_reg(const reg_t *r, int i, int *n_regs, reg_t *regs)
{
    regs = (reg_t *)realloc(regs, (size_t)(*n_regs + 1) * sizeof(reg_t));
    if (i + 1 != *n_regs) memmove(&regs[i + 2], &regs[i + 1], sizeof(reg_t) * (size_t)(*n_regs - i - 1));
    regs[i + 1] = *r;
    ++*n_regs;
    return regs;
}

static reg_t *align_skeleton(actx_t *c, int *n_regs_, reg_t *regs, orc128_t *a)
{
    const orc_opt_t *opt = c->opt;
    int32_t i, n_regs = *n_regs_, n_a;
    n_a = squeeze_a(n_regs, regs, a);
    for (i = 0; i < n_regs; ++i) {
        reg_t r2;
        memset(&r2, 0, sizeof(r2));
        align1(c, &regs[i], &r2, n_a, a);
        if (r2.cnt > 0) regs = insert_reg(&r2, i, &n_regs, regs);
        if (i > 0 && regs[i].split_inv) {
            if (align1_inv(c, &regs[i - 1], &regs[i], &r2)) {
                regs = insert_reg(&r2, i, &n_regs, regs);
                ++i;
            }
        }
    }
    *n_regs_ = n_regs;
    filter_regs(opt, c->qlen, n_regs_, regs);
    if (c->qlen >= opt->rank_min_len) {
        update_dp_max(c->qlen, *n_regs_, regs, opt->rank_frac, opt->a, opt->b);
        filter_regs(opt, c->qlen, n_regs_, regs);
    }
    hit_sort(n_regs_, regs);
    return regs;
}

/* ------------------------------- mm_map_frag ------------------------------- */
void orc_dbg_free(orc_dbg_t *d)
{
    free(d->mz_x); free(d->mz_y); free(d->a); free(d->u); free(d->ca); free(d->regs0);
    memset(d, 0, sizeof(*d));
}

typedef struct { idx_t mi; const uint8_t *seq; int32_t len; } cidx_t;

static int map_with_index(const orc_opt_t *opt, const idx_t *mi, const uint8_t *contig, int32_t clen,
                          const uint8_t *read, int32_t qlen, uint32_t name_hash,
                          orc_aln_t *aln, int aln_cap, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cigar,
                          int64_t *dp_cells, int64_t *n_dp_tasks, int64_t *n_mz_out, int64_t *n_a_out, orc_dbg_t *dbg)
{
    int64_t n_mz, n_a, cap = qlen + 16;
    uint64_t *mx, *my, *u = 0;
    orc128_t *a;
    uint32_t hash;
    int n_regs0 = 0, i, n_out = 0;
    int32_t rep_len = 0;
    float chn_pen_gap, chn_pen_skip;
    reg_t *regs0;
    int32_t *mapq = 0;
    arena_t A = {0};
    actx_t c;
    uint8_t *qrc;

    if (dbg) memset(dbg, 0, sizeof(*dbg));
    if (qlen <= 0) return 0;
    hash = name_hash;
    hash ^= wang_hash((uint32_t)qlen) + wang_hash((uint32_t)opt->seed);
    hash = wang_hash(hash);

    mx = (uint64_t *)malloc((size_t)cap * 8), my = (uint64_t *)malloc((size_t)cap * 8);
    n_mz = orc_sketch(read, qlen, opt->w, opt->k, opt->hpc, mx, my, cap);
    if (opt->q_occ_frac > 0.0f) n_mz = seed_mz_flt(n_mz, mx, my, mi->mid_occ, opt->q_occ_frac);
    if (n_mz_out) *n_mz_out += n_mz;
    if (dbg) {
        dbg->n_mz = n_mz; dbg->mid_occ = mi->mid_occ;
        dbg->mz_x = (uint64_t *)malloc((size_t)(n_mz + 1) * 8); dbg->mz_y = (uint64_t *)malloc((size_t)(n_mz + 1) * 8);
        memcpy(dbg->mz_x, mx, (size_t)n_mz * 8); memcpy(dbg->mz_y, my, (size_t)n_mz * 8);
    }
    a = collect_seed_hits(opt, mi, qlen, n_mz, mx, my, &n_a, &rep_len);
    free(mx); free(my);
    if (n_a_out) *n_a_out += n_a;
    if (dbg) {
        dbg->n_a = n_a; dbg->a = (orc128_t *)malloc((size_t)(n_a + 1) * 16);
        memcpy(dbg->a, a, (size_t)n_a * 16);
    }
    chn_pen_gap = (float)(opt->chain_gap_scale * 0.01 * opt->k);
    chn_pen_skip = (float)(opt->chain_skip_scale * 0.01 * opt->k);
    a = lchain_dp(opt->max_gap, opt->max_gap, opt->bw, opt->max_chain_skip, opt->max_chain_iter, opt->min_cnt,
                  opt->min_chain_score, chn_pen_gap, chn_pen_skip, n_a, a, &n_regs0, &u);
    if (opt->bw_long > opt->bw && n_regs0 > 1) {
        int32_t st = (int32_t)a[0].y, en = (int32_t)a[(int32_t)u[0] - 1].y;
        if (qlen - (en - st) > opt->rmq_rescue_size || en - st > qlen * opt->rmq_rescue_ratio) {
            for (i = 0, n_a = 0; i < n_regs0; ++i) n_a += (int32_t)u[i];
            free(u); u = 0;
            orc_radix_sort_128x(a, a + n_a);
            a = lchain_rmq(opt->max_gap, opt->rmq_inner_dist, opt->bw_long, opt->max_chain_skip, opt->rmq_size_cap,
                           opt->min_cnt, opt->min_chain_score, chn_pen_gap, chn_pen_skip, n_a, a, &n_regs0, &u);
            if (dbg) dbg->rechained = 1;
        }
    }
    if (dbg) {
        int64_t nca = 0;
        for (i = 0; i < n_regs0; ++i) nca += (int32_t)u[i];
        dbg->n_u = n_regs0; dbg->u = (uint64_t *)malloc((size_t)(n_regs0 + 1) * 8);
        if (n_regs0) memcpy(dbg->u, u, (size_t)n_regs0 * 8);
        dbg->n_ca = nca; dbg->ca = (orc128_t *)malloc((size_t)(nca + 1) * 16);
        if (nca) memcpy(dbg->ca, a, (size_t)nca * 16);
    }
    regs0 = gen_regs(hash, qlen, n_regs0, u, a);
    free(u);
    if (n_regs0 > 0) {
        /* chain_post */
        set_parent(opt->mask_level, opt->mask_len, n_regs0, regs0, opt->a * 2 + opt->b);
        select_sub(opt->pri_ratio, opt->k * 2, opt->best_n, 1, (int)(opt->max_gap * 0.8), &n_regs0, regs0);
    }
    if (dbg) {
        dbg->n_regs0 = n_regs0; dbg->regs0 = (int32_t *)malloc((size_t)(n_regs0 + 1) * 10 * 4);
        for (i = 0; i < n_regs0; ++i) {
            int32_t *o = dbg->regs0 + i * 10;
            reg_t *r = &regs0[i];
            o[0] = r->as, o[1] = r->cnt, o[2] = r->score, o[3] = r->parent, o[4] = r->rs, o[5] = r->re,
            o[6] = r->qs, o[7] = r->qe, o[8] = r->rev, o[9] = (int32_t)r->hash;
        }
    }
    if (n_regs0 > 0) {
        /* align_regs */
        memset(&c, 0, sizeof(c));
        qrc = (uint8_t *)malloc((size_t)qlen);
        for (i = 0; i < qlen; ++i) qrc[qlen - 1 - i] = read[i] < 4 ? 3 - read[i] : 4;
        c.opt = opt; c.tseq0 = contig; c.tlen0 = clen; c.qseq0[0] = read; c.qseq0[1] = qrc; c.qlen = qlen; c.A = &A;
        regs0 = align_skeleton(&c, &n_regs0, regs0, a);
        set_parent(opt->mask_level, opt->mask_len, n_regs0, regs0, opt->a * 2 + opt->b);
        select_sub(opt->pri_ratio, opt->k * 2, opt->best_n, 0, (int)(opt->max_gap * 0.8), &n_regs0, regs0);
        set_sam_pri(n_regs0, regs0);
        mapq = (int32_t *)calloc((size_t)n_regs0 + 1, 4);
        set_mapq(n_regs0, regs0, opt->min_chain_score, opt->a, rep_len, mapq);
        if (dp_cells) *dp_cells += c.cells;
        if (n_dp_tasks) *n_dp_tasks += c.n_tasks;
        free(c.ez.cigar);
        free(qrc);
        for (i = 0; i < n_regs0; ++i) {
            reg_t *r = &regs0[i];
            orc_aln_t *o;
            if (n_out >= aln_cap) { n_out = -1; break; }
            if (*n_cigar + r->n_cigar > cigar_cap) { n_out = -1; break; }
            o = &aln[n_out++];
            o->read = 0, o->strand = 0;
            o->rs = r->rs, o->re = r->re, o->qs = r->qs, o->qe = r->qe, o->rev = r->rev;
            o->flag = (r->rev ? 0x10 : 0) | (r->parent != r->id ? 0x100 : !r->sam_pri ? 0x800 : 0);
            o->dp_max = r->dp_max, o->mlen = r->mlen, o->blen = r->blen;
            o->n_cigar = r->n_cigar, o->cigar_off = *n_cigar;
            o->mapq = mapq[i], o->dp_score = r->dp_score, o->cnt = r->cnt, o->score = r->score, o->subsc = r->subsc;
            o->n_ambi = r->n_ambi, o->inv = r->inv, o->n_sub = r->n_sub;
            if (r->n_cigar) memcpy(cigar + *n_cigar, r->cigar, (size_t)r->n_cigar * 4);
            *n_cigar += r->n_cigar;
        }
    }
    free(mapq);
    free(regs0);
    free(a);
    afree_all(&A);
    return n_out;
}

int orc_map_one(const orc_opt_t *opt, const uint8_t *contig, int32_t clen, const uint8_t *read, int32_t qlen,
                uint32_t name_hash, orc_aln_t *aln, int aln_cap, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cigar,
                int64_t *dp_cells, int64_t *n_dp_tasks, orc_dbg_t *dbg)
{
    idx_t mi;
    int n;
    idx_build(&mi, opt, contig, clen);
    n = map_with_index(opt, &mi, contig, clen, read, qlen, name_hash, aln, aln_cap, cigar, cigar_cap, n_cigar, dp_cells,
                       n_dp_tasks, 0, 0, dbg);
    idx_free(&mi);
    return n;
}

/* exported for orc_af.c */
void *orc_idx_new(const orc_opt_t *opt, const uint8_t *contig, int32_t clen)
{
    idx_t *mi = (idx_t *)malloc(sizeof(idx_t));
    idx_build(mi, opt, contig, clen);
    return mi;
}
void orc_idx_del(void *mi) { idx_free((idx_t *)mi); free(mi); }
int64_t orc_idx_size(void *mi) { return ((idx_t *)mi)->n; }
int orc_map_idx(const orc_opt_t *opt, void *mi, const uint8_t *contig, int32_t clen, const uint8_t *read, int32_t qlen,
                uint32_t name_hash, orc_aln_t *aln, int aln_cap, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cigar,
                int64_t *dp_cells, int64_t *n_dp_tasks, int64_t *n_mz, int64_t *n_a)
{
    return map_with_index(opt, (idx_t *)mi, contig, clen, read, qlen, name_hash, aln, aln_cap, cigar, cigar_cap, n_cigar,
                          dp_cells, n_dp_tasks, n_mz, n_a, 0);
}
