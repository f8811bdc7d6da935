/*
 * oracle/orc_af.c — whole stage-4 body on a packed batch: per locus map every read to both contig
 * strands, pile up depth, take window medians, compute AF.  TEST INFRASTRUCTURE ONLY.
 *
 * [REF] restated here (pinned by tests/test_af_arith_ref.py, which executes the reference's own
 * functions by AST extraction):
 *   get_te_cov            TELR_te.py:841-867       get_flank_cov        TELR_te.py:518-550
 *   get_median_cov        TELR_te.py:870-884       get_te_flank_ratio   TELR_te.py:564-575
 *   rc TE coordinates     TELR_te.py:668-673       AF block             TELR_te.py:810-835
 * [UP] restated here (unpinned): samtools 1.9 `depth -aa -r chr:S-E` region and flag semantics
 *   (bam2depth.c, hts_parse_reg): beg = max(S-1,0), end = E, positions [beg, min(end, L));
 *   records with flag & (UNMAP|SECONDARY|QCFAIL|DUP) skipped; deletions not counted.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "orc.h"

/* command-line style overrides on top of the preset (minimap2 -r NUM[,NUM]); 0 = preset value */
int orc_override_bw = 0, orc_override_bw_long = 0;
void orc_set_bw(int bw, int bw_long) { orc_override_bw = bw; orc_override_bw_long = bw_long; }


void *orc_idx_new(const orc_opt_t *opt, const uint8_t *contig, int32_t clen);
void orc_idx_del(void *mi);
int64_t orc_idx_size(void *mi);
int orc_map_idx(const orc_opt_t *opt, void *mi, const uint8_t *contig, int32_t clen, const uint8_t *read, int32_t qlen,
                uint32_t name_hash, orc_aln_t *aln, int aln_cap, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cigar,
                int64_t *dp_cells, int64_t *n_dp_tasks, int64_t *n_mz, int64_t *n_a);

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* 2*median of depth over samtools region "c:S-E" on a contig of length L; -1 if the region is empty */
int32_t orc_median2x(const int32_t *depth, int32_t L, int32_t S, int32_t E)
{
    int64_t beg = (int64_t)S - 1, end = E;
    int32_t n, r, *tmp;
    if (beg < 0) beg = 0;
    if (end > L) end = L;
    if (beg >= end) return -1;
    n = (int32_t)(end - beg);
    tmp = (int32_t *)malloc((size_t)n * 4);
    memcpy(tmp, depth + beg, (size_t)n * 4);
    qsort(tmp, (size_t)n, 4, cmp_i32);
    r = (n & 1) ? 2 * tmp[n / 2] : tmp[n / 2 - 1] + tmp[n / 2];
    free(tmp);
    return r;
}

static void cov4(const int32_t *depth, int32_t L, int32_t s, int32_t e, int32_t fl, int32_t fo, int32_t ti, int32_t to,
                 int32_t out[4])
{
    int whole = 0;
    out[0] = out[1] = out[2] = out[3] = -1;
    if (ti) {
        if (s + to + ti < e) {
            out[0] = orc_median2x(depth, L, s + to, s + to + ti);
            out[1] = orc_median2x(depth, L, e - ti - to, e - to);
            if (out[0] < 0) out[0] = -3;
            if (out[1] < 0) out[1] = -3;
        } else whole = 1;
    } else whole = 1;
    if (whole) {
        out[0] = orc_median2x(depth, L, s, e);
        if (out[0] < 0) out[0] = -3;
        out[1] = out[0];
    }
    if (s - fl - fo >= 0) {
        out[2] = orc_median2x(depth, L, s - fl - fo, s - fo);
        if (out[2] < 0) out[2] = -3;
    }
    if (e + fl + fo <= L) {
        out[3] = orc_median2x(depth, L, e + fo, e + fl + fo);
        if (out[3] < 0) out[3] = -3;
    }
}

static int ratio(int32_t te2, int32_t fl2, double *r)
{
    if (te2 <= 0 || fl2 <= 0) return 0;        /* None or 0 are falsy */
    *r = ((double)te2 / 2.0) / ((double)fl2 / 2.0);
    if (*r > 1.5) return 0;
    return 1;
}

void orc_cov_af(const int32_t *dfw, const int32_t *drc, int32_t L, int32_t s, int32_t e, int32_t fl, int32_t fo,
                int32_t ti, int32_t to, int32_t cov2x[8], double *af)
{
    double t5 = 0, t3 = 0, f;
    int h5, h3;
    cov4(dfw, L, s, e, fl, fo, ti, to, cov2x);
    cov4(drc, L, L - e, L - s, fl, fo, ti, to, cov2x + 4);
    h5 = ratio(cov2x[0], cov2x[2], &t5);
    h3 = ratio(cov2x[4], cov2x[6], &t3);
    if (h5 && h3) f = fabs(t5 - t3) <= 0.3 ? (t5 + t3) / 2 : NAN;
    else if (h5) f = t5;
    else if (h3) f = t3;
    else f = NAN;
    *af = f;
}

static void unpack(const orc_batch_t *b, int64_t off, int32_t len, uint8_t *dst)
{
    for (int32_t i = 0; i < len; ++i) {
        int64_t p = off + i;
        int c = (b->seq2[p >> 4] >> (2 * (p & 15))) & 3;
        if (b->nmask[p >> 5] >> (p & 31) & 1) c = 4;
        dst[i] = (uint8_t)c;
    }
}

typedef struct {
    orc_aln_t *aln; int n_aln, m_aln;
    uint32_t *cigar; int64_t n_cigar, m_cigar;
} lres_t;

int orc_af_run(const orc_batch_t *b, orc_result_t *res, int n_threads, int first_locus, int n_run)
{
    orc_opt_t opt;
    int32_t l, last;
    int64_t *depth_off = 0;
    lres_t *lr;
    int64_t dp_cells = 0, n_tasks = 0, n_mz = 0, n_anch = 0, n_blocks = 0;
    int err = 0;
    orc_opt_preset(&opt, b->preset);
    if (orc_override_bw > 0) opt.bw = orc_override_bw;
    if (orc_override_bw_long > 0) opt.bw_long = orc_override_bw_long;
    if (n_run <= 0) n_run = b->n_loci - first_locus;
    last = first_locus + n_run;
    if (last > b->n_loci) last = b->n_loci;
    depth_off = (int64_t *)malloc((size_t)(b->n_loci + 1) * 8);
    depth_off[0] = 0;
    for (l = 0; l < b->n_loci; ++l) depth_off[l + 1] = depth_off[l] + 2 * (int64_t)b->contig_len[l];
    lr = (lres_t *)calloc((size_t)b->n_loci + 1, sizeof(lres_t));
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : dp_cells, n_tasks, n_mz, n_anch, n_blocks)
    for (l = first_locus; l < last; ++l) {
        int32_t L = b->contig_len[l], s, i, rb = b->locus_read_begin[l], re_ = b->locus_read_begin[l + 1];
        uint8_t *ctg[2], *rd = 0;
        int32_t *dep[2], rd_cap = 0;
        void *mi[2];
        lres_t *o = &lr[l];
        if (L <= 0) {
            for (i = 0; i < 8; ++i) res->cov2x[(int64_t)l * 8 + i] = -2;
            res->af[l] = NAN;
            continue;
        }
        ctg[0] = (uint8_t *)malloc((size_t)L); ctg[1] = (uint8_t *)malloc((size_t)L);
        unpack(b, b->contig_off[l], L, ctg[0]);
        for (i = 0; i < L; ++i) ctg[1][L - 1 - i] = ctg[0][i] < 4 ? 3 - ctg[0][i] : 4;
        dep[0] = (int32_t *)calloc((size_t)L, 4); dep[1] = (int32_t *)calloc((size_t)L, 4);
        for (s = 0; s < 2; ++s) {
            mi[s] = orc_idx_new(&opt, ctg[s], L);
            n_mz += orc_idx_size(mi[s]);
        }
        for (s = 0; s < 2; ++s) {
            for (i = rb; i < re_; ++i) {
                int32_t qlen = b->read_len[i], n, k;
                if (qlen > rd_cap) { rd_cap = qlen + 1024; rd = (uint8_t *)realloc(rd, (size_t)rd_cap); }
                unpack(b, b->read_off[i], qlen, rd);
                for (;;) {
                    int64_t nc0 = o->n_cigar;
                    if (o->m_aln - o->n_aln < 64) {
                        o->m_aln = o->m_aln * 2 + 128;
                        o->aln = (orc_aln_t *)realloc(o->aln, (size_t)o->m_aln * sizeof(orc_aln_t));
                    }
                    if (o->m_cigar - o->n_cigar < 4 * (int64_t)qlen + 1024) {
                        o->m_cigar = o->m_cigar * 2 + 4 * (int64_t)qlen + 4096;
                        o->cigar = (uint32_t *)realloc(o->cigar, (size_t)o->m_cigar * 4);
                    }
                    {
                        int64_t c0 = 0, t0 = 0, m0 = 0, a0 = 0;
                        n = orc_map_idx(&opt, mi[s], ctg[s], L, rd, qlen, b->read_hash[i], o->aln + o->n_aln,
                                        o->m_aln - o->n_aln, o->cigar, o->m_cigar, &o->n_cigar, &c0, &t0, &m0, &a0);
                        if (n < 0) { o->n_cigar = nc0; o->m_aln *= 2; o->m_cigar *= 2; continue; }
                        dp_cells += c0, n_tasks += t0, n_mz += m0, n_anch += a0;
                    }
                    break;
                }
                for (k = 0; k < n; ++k) {
                    orc_aln_t *al = &o->aln[o->n_aln + k];
                    al->read = i, al->strand = s;
                    if (!(al->flag & 0x100)) {          /* samtools depth skips SECONDARY only (of what minimap2 emits) */
                        int32_t pos = al->rs, c;
                        for (c = 0; c < al->n_cigar; ++c) {
                            uint32_t op = o->cigar[al->cigar_off + c] & 0xf, len = o->cigar[al->cigar_off + c] >> 4;
                            if (op == 0) {
                                uint32_t x;
                                for (x = 0; x < len; ++x)
                                    if (pos + (int32_t)x >= 0 && pos + (int32_t)x < L) ++dep[s][pos + x];
                                pos += len; ++n_blocks;
                            } else if (op == 2) pos += len;
                        }
                    }
                }
                o->n_aln += n;
            }
        }
        if (res->depth) {
            memcpy(res->depth + depth_off[l], dep[0], (size_t)L * 4);
            memcpy(res->depth + depth_off[l] + L, dep[1], (size_t)L * 4);
        }
        if (b->te_start[l] < 0) {
            for (i = 0; i < 8; ++i) res->cov2x[(int64_t)l * 8 + i] = -2;
            res->af[l] = NAN;
        } else
            orc_cov_af(dep[0], dep[1], L, b->te_start[l], b->te_end[l], b->flank_len, b->flank_off, b->te_len, b->te_off,
                       res->cov2x + (int64_t)l * 8, &res->af[l]);
        for (s = 0; s < 2; ++s) { orc_idx_del(mi[s]); free(ctg[s]); free(dep[s]); }
        free(rd);
    }
    /* gather alignment records in locus order */
    res->n_aln = res->n_cigar = 0;
    for (l = first_locus; l < last; ++l) {
        lres_t *o = &lr[l];
        if (res->aln && res->cigar) {
            if (res->n_aln + o->n_aln > res->aln_cap || res->n_cigar + o->n_cigar > res->cigar_cap) err = -5;
            else {
                int k;
                for (k = 0; k < o->n_aln; ++k) {
                    res->aln[res->n_aln + k] = o->aln[k];
                    res->aln[res->n_aln + k].cigar_off += res->n_cigar;
                }
                if (o->n_cigar) memcpy(res->cigar + res->n_cigar, o->cigar, (size_t)o->n_cigar * 4);
                res->n_aln += o->n_aln, res->n_cigar += o->n_cigar;
            }
        }
        free(o->aln); free(o->cigar);
    }
    free(lr); free(depth_off);
    res->dp_cells = dp_cells, res->n_dp_tasks = n_tasks, res->n_minimizers = n_mz, res->n_anchors = n_anch;
    res->n_aln_blocks = n_blocks;
    return err;
}
