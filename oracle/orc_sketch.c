/*
 * oracle/orc_sketch.c — options, hashes, radix sorts, minimizer sketch.  TEST INFRASTRUCTURE ONLY.
 * PARITY UNPINNED w.r.t. minimap2 2.22 (see orc.h).  Each function names the upstream [UP] routine
 * it restates and the reference call site that reaches it (TELR_te.py:505 `minimap2 -a -x <preset>`).
 */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "orc.h"

/* [UP] options.c mm_idxopt_init / mm_mapopt_init / mm_set_opt for map-ont, map-pb, map-hifi (2.22).
 * Reached from TELR_te.py:595-598 (preset choice) and :505 (-x preset). */
void orc_opt_preset(orc_opt_t *o, int preset)
{
    memset(o, 0, sizeof(*o));
    o->k = 15; o->w = 10; o->hpc = 0;
    o->a = 2; o->b = 4; o->q = 4; o->e = 2; o->q2 = 24; o->e2 = 1; o->sc_ambi = 1;
    o->zdrop = 400; o->zdrop_inv = 200; o->end_bonus = -1;
    o->min_dp_max = 80; o->min_ksw_len = 200;
    o->bw = 500; o->bw_long = 20000; o->max_gap = 5000;
    o->max_chain_skip = 25; o->max_chain_iter = 5000; o->min_cnt = 3; o->min_chain_score = 40;
    o->rmq_inner_dist = 1000; o->rmq_size_cap = 100000; o->rmq_rescue_size = 1000;
    o->rmq_rescue_ratio = 0.1f; o->chain_gap_scale = 0.8f; o->chain_skip_scale = 0.0f;
    o->mask_level = 0.5f; o->mask_len = INT32_MAX; o->pri_ratio = 0.8f; o->best_n = 5;
    o->q_occ_frac = 0.01f; o->mid_occ_frac = 2e-4f; o->min_mid_occ = 10; o->max_mid_occ = 1000000;
    o->max_max_occ = 4095; o->occ_dist = 500;
    o->seed = 11; o->max_sw_mat = 100000000; o->rank_min_len = 500; o->rank_frac = 0.9f;
    o->max_clip_ratio = 1.0f;
    if (preset == 1) {            /* map-pb */
        o->hpc = 1; o->k = 19;
    } else if (preset == 2) {     /* map-hifi */
        o->k = 19; o->w = 19; o->max_gap = 10000;
        o->a = 1; o->b = 4; o->q = 6; o->q2 = 26; o->e = 2; o->e2 = 1;
        o->min_mid_occ = 50; o->max_mid_occ = 500; o->min_dp_max = 200;
    }
}

/* [UP] khash.h __ac_X31_hash_string — feeds the region tie-break hash in mm_map_frag (map.c) */
uint32_t orc_name_hash(const char *s)
{
    uint32_t h = (uint32_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    return h;
}

int orc_pack_seq(const char *s, int32_t len, int64_t off, uint32_t *seq2, uint32_t *nmask)
{
    if (off & 63) return -1;
    int64_t nw = ((int64_t)len + 63) / 64 * 4;
    memset(seq2 + off / 16, 0, (size_t)nw * 4);
    memset(nmask + off / 32, 0, (size_t)(nw / 2) * 4);
    for (int32_t i = 0; i < len; ++i) {
        int c;
        switch (s[i]) {
        case 'A': case 'a': c = 0; break;
        case 'C': case 'c': c = 1; break;
        case 'G': case 'g': c = 2; break;
        case 'T': case 't': case 'U': case 'u': c = 3; break;
        default: c = 4;
        }
        int64_t p = off + i;
        if (c < 4) seq2[p >> 4] |= (uint32_t)c << (2 * (p & 15));
        else nmask[p >> 5] |= 1u << (p & 31);
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * [UP] ksort.h KRADIX_SORT_INIT(128x, mm128_t, .x, 8) and (64, uint64_t, identity, 8):
 * insertion sort up to 64 elements, else in-place MSD radix (8-bit digits, most significant
 * byte first) whose permutation step is NOT stable.  The tie order it leaves is observable
 * (anchors with equal x, chains with equal score), so it is restated operation by operation.
 * ------------------------------------------------------------------------------------------- */
#define RS_MIN_SIZE 64
#define RS_MAX_BITS 8

#define DEF_RADIX(NAME, T, KEY)                                                             \
    typedef struct { T *b, *e; } rsb_##NAME##_t;                                            \
    static void ins_##NAME(T *beg, T *end)                                                  \
    {                                                                                       \
        for (T *i = beg + 1; i < end; ++i)                                                  \
            if (KEY(*i) < KEY(*(i - 1))) {                                                  \
                T *j, tmp = *i;                                                             \
                for (j = i; j > beg && KEY(tmp) < KEY(*(j - 1)); --j) *j = *(j - 1);        \
                *j = tmp;                                                                   \
            }                                                                               \
    }                                                                                       \
    static void rs_##NAME(T *beg, T *end, int n_bits, int s)                                \
    {                                                                                       \
        int size = 1 << n_bits, m = size - 1;                                               \
        rsb_##NAME##_t b[1 << RS_MAX_BITS], *k, *be = b + size;                             \
        for (k = b; k != be; ++k) k->b = k->e = beg;                                        \
        for (T *i = beg; i != end; ++i) ++b[KEY(*i) >> s & m].e;                            \
        for (k = b + 1; k != be; ++k) k->e += (k - 1)->e - beg, k->b = (k - 1)->e;          \
        for (k = b; k != be;) {                                                             \
            if (k->b != k->e) {                                                             \
                rsb_##NAME##_t *l;                                                          \
                if ((l = b + (KEY(*k->b) >> s & m)) != k) {                                 \
                    T tmp = *k->b, swap;                                                    \
                    do {                                                                    \
                        swap = tmp; tmp = *l->b; *l->b++ = swap;                            \
                        l = b + (KEY(tmp) >> s & m);                                        \
                    } while (l != k);                                                       \
                    *k->b++ = tmp;                                                          \
                } else ++k->b;                                                              \
            } else ++k;                                                                     \
        }                                                                                   \
        for (b->b = beg, k = b + 1; k != be; ++k) k->b = (k - 1)->e;                        \
        if (s) {                                                                            \
            s = s > n_bits ? s - n_bits : 0;                                                \
            for (k = b; k != be; ++k)                                                       \
                if (k->e - k->b > RS_MIN_SIZE) rs_##NAME(k->b, k->e, n_bits, s);            \
                else if (k->e - k->b > 1) ins_##NAME(k->b, k->e);                           \
        }                                                                                   \
    }

#define KEY128(a) ((a).x)
#define KEY64(a) (a)
DEF_RADIX(128x, orc128_t, KEY128)
DEF_RADIX(64, uint64_t, KEY64)

void orc_radix_sort_128x(orc128_t *beg, orc128_t *end)
{
    if (end - beg <= RS_MIN_SIZE) ins_128x(beg, end);
    else rs_128x(beg, end, RS_MAX_BITS, (8 - 1) * RS_MAX_BITS);
}
void orc_radix_sort_64(uint64_t *beg, uint64_t *end)
{
    if (end - beg <= RS_MIN_SIZE) ins_64(beg, end);
    else rs_64(beg, end, RS_MAX_BITS, (8 - 1) * RS_MAX_BITS);
}

/* ---------------------------------------------------------------------------------------------
 * [UP] sketch.c hash64 (invertible integer hash, masked) and mm_sketch.
 * ------------------------------------------------------------------------------------------- */
static inline uint64_t hash64m(uint64_t key, uint64_t mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

typedef struct { int front, count; int a[32]; } tiny_queue_t;
static inline void tq_push(tiny_queue_t *q, int x) { q->a[((q->count++) + q->front) & 0x1f] = x; }
static inline int tq_shift(tiny_queue_t *q)
{
    int x;
    if (q->count == 0) return -1;
    x = q->a[q->front++];
    q->front &= 0x1f;
    --q->count;
    return x;
}

#define PUSH(INFO) do { if (n < cap) { ox[n] = (INFO).x; oy[n] = (INFO).y; } ++n; } while (0)

int64_t orc_sketch(const uint8_t *seq, int32_t len, int32_t w, int32_t k, int32_t is_hpc,
                   uint64_t *ox, uint64_t *oy, int64_t cap)
{
    uint64_t shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1, kmer[2] = {0, 0};
    int i, j, l, buf_pos, min_pos, kmer_span = 0;
    orc128_t buf[256], min = {UINT64_MAX, UINT64_MAX};
    tiny_queue_t tq;
    int64_t n = 0;

    if (!(len > 0 && (w > 0 && w < 256) && (k > 0 && k <= 28))) return 0;
    memset(buf, 0xff, (size_t)w * 16);
    memset(&tq, 0, sizeof(tq));

    for (i = l = buf_pos = min_pos = 0; i < len; ++i) {
        int c = seq[i];
        orc128_t info = {UINT64_MAX, UINT64_MAX};
        if (c < 4) {
            int z;
            if (is_hpc) {
                int skip_len = 1;
                if (i + 1 < len && seq[i + 1] == c) {
                    for (skip_len = 2; i + skip_len < len; ++skip_len)
                        if (seq[i + skip_len] != c) break;
                    i += skip_len - 1;
                }
                tq_push(&tq, skip_len);
                kmer_span += skip_len;
                if (tq.count > k) kmer_span -= tq_shift(&tq);
            } else kmer_span = l + 1 < k ? l + 1 : k;
            kmer[0] = (kmer[0] << 2 | (uint64_t)c) & mask;
            kmer[1] = (kmer[1] >> 2) | (3ULL ^ (uint64_t)c) << shift1;
            if (kmer[0] == kmer[1]) continue;
            z = kmer[0] < kmer[1] ? 0 : 1;
            ++l;
            if (l >= k && kmer_span < 256) {
                info.x = hash64m(kmer[z], mask) << 8 | (uint64_t)kmer_span;
                info.y = (uint64_t)0 << 32 | (uint32_t)i << 1 | (uint32_t)z;
            }
        } else l = 0, tq.count = tq.front = 0, kmer_span = 0;
        buf[buf_pos] = info;
        if (l == w + k - 1 && min.x != UINT64_MAX) {
            for (j = buf_pos + 1; j < w; ++j)
                if (min.x == buf[j].x && buf[j].y != min.y) PUSH(buf[j]);
            for (j = 0; j < buf_pos; ++j)
                if (min.x == buf[j].x && buf[j].y != min.y) PUSH(buf[j]);
        }
        if (info.x <= min.x) {
            if (l >= w + k && min.x != UINT64_MAX) PUSH(min);
            min = info, min_pos = buf_pos;
        } else if (buf_pos == min_pos) {
            if (l >= w + k - 1 && min.x != UINT64_MAX) PUSH(min);
            for (j = buf_pos + 1, min.x = UINT64_MAX; j < w; ++j)
                if (min.x >= buf[j].x) min = buf[j], min_pos = j;
            for (j = 0; j <= buf_pos; ++j)
                if (min.x >= buf[j].x) min = buf[j], min_pos = j;
            if (l >= w + k - 1 && min.x != UINT64_MAX) {
                for (j = buf_pos + 1; j < w; ++j)
                    if (min.x == buf[j].x && min.y != buf[j].y) PUSH(buf[j]);
                for (j = 0; j <= buf_pos; ++j)
                    if (min.x == buf[j].x && min.y != buf[j].y) PUSH(buf[j]);
            }
        }
        if (++buf_pos == w) buf_pos = 0;
    }
    if (min.x != UINT64_MAX) PUSH(min);
    return n;
}
