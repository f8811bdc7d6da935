/*
 * oracle/orc.h — CPU oracle for the TELR stage-4 AF path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (telr_b200/) never does.
 *
 * PARITY UNPINNED at the minimap2/samtools boundary: the arithmetic of this path lives in
 * minimap2 2.22 and samtools/htslib 1.9 (envs/telr.yml:45,78,24 of the reference), neither
 * of which is vendored in /root/reference or installed in this image, and the reference
 * ships no golden vectors.  Everything marked [UP] below is a restatement of the published
 * upstream algorithm from its call sites in the reference (TELR_te.py:505, :872).
 * The TELR-side arithmetic ([REF], TELR_te.py:518-575, 656-675, 757-884) IS pinned: tests
 * execute the reference's own functions (AST-extracted) against this restatement.
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t x, y; } orc128_t;

typedef struct orc_opt {
    int32_t k, w, hpc;
    int32_t a, b, q, e, q2, e2, sc_ambi;
    int32_t zdrop, zdrop_inv, end_bonus;
    int32_t min_dp_max, min_ksw_len;
    int32_t bw, bw_long, max_gap;
    int32_t max_chain_skip, max_chain_iter, min_cnt, min_chain_score;
    int32_t rmq_inner_dist, rmq_size_cap, rmq_rescue_size;
    float rmq_rescue_ratio, chain_gap_scale, chain_skip_scale;
    float mask_level;
    int32_t mask_len;
    float pri_ratio;
    int32_t best_n;
    float q_occ_frac, mid_occ_frac;
    int32_t min_mid_occ, max_mid_occ;
    int32_t max_max_occ, occ_dist;
    int32_t seed;
    int64_t max_sw_mat;
    int32_t rank_min_len;
    float rank_frac, max_clip_ratio;
} orc_opt_t;

void orc_opt_preset(orc_opt_t *o, int preset);
void orc_set_bw(int bw, int bw_long);   /* minimap2 -r NUM[,NUM] on top of the preset for orc_af_run (0 = preset value) */

/* Reach counters of the stated deviations from upstream: how often an input reaches a spot where this
 * restatement knowingly differs from minimap2 2.22.  0 = query minimizers dropped by the plain `n > mid_occ` filter (upstream:
 * mm_seed_select may rescue some), 1 = RMQ queries whose best priority is tied, 2 = banded DP calls whose traceback path
 * comes within one cell of a band-limited edge (16-lane band rounding), 3 = ksw_ll calls with a tied maximum, 4 = index
 * buckets with more than 64 entries (unstable radix sort order), 5 = banded DP calls, 6 = RMQ queries, 7 = ksw_ll calls. */
extern int64_t orc_dev[8];
void orc_dev_counters(int64_t out[8], int reset);

/* [UP] mm_sketch (sketch.c). seq: nt4 codes (0..3, 4 = ambiguous). Returns #minimizers (may exceed cap: then truncated). */
int64_t orc_sketch(const uint8_t *seq, int32_t len, int32_t w, int32_t k, int32_t hpc,
                   uint64_t *x, uint64_t *y, int64_t cap);

/* [UP] radix_sort_128x / radix_sort_64 (ksort.h): unstable in-place MSD radix sort, restated verbatim in behaviour */
void orc_radix_sort_128x(orc128_t *beg, orc128_t *end);
void orc_radix_sort_64(uint64_t *beg, uint64_t *end);

/* ksw_extd2 restatement (scalar, clean-band semantics) */
#define ORC_KSW_EXTZ_ONLY  0x40
#define ORC_KSW_RIGHT      0x02
#define ORC_KSW_REV_CIGAR  0x80
#define ORC_KSW_APPROX_MAX 0x08
typedef struct orc_ez {
    int32_t max, max_q, max_t, mqe, mqe_t, mte, mte_q, score, zdropped, reach_end;
    int32_t n_cigar, m_cigar;
    uint32_t *cigar;
    int64_t cells;
} orc_ez_t;
void orc_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int a, int b, int sc_ambi, int q, int e, int q2, int e2,
                   int w, int zdrop, int end_bonus, int flag, orc_ez_t *ez);
/* bounded extension (work-saving, exact for every consumed output): on by default; 0 = compute every anti-diagonal like ksw2 */
void orc_set_ext_bound(int on);
/* [UP] ksw_ll_i16: local affine SW, score + end coordinates */
int orc_ksw_ll(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
               int a, int b, int sc_ambi, int gapo, int gape, int *qe, int *te);

/* one alignment record */
typedef struct orc_aln {
    int32_t read, strand;
    int32_t rs, re, qs, qe, rev, flag;
    int32_t dp_max, mlen, blen;
    int32_t n_cigar;
    int64_t cigar_off;
    int32_t mapq, dp_score, cnt, score, subsc, n_ambi, inv, n_sub;   /* SAM MAPQ and the AS / cm / s1 / s2 / nn / tp tags */
} orc_aln_t;

/* batch identical in layout to telr_af_batch (include/telr_af.h) */
typedef struct orc_batch {
    int32_t preset, flank_len, flank_off, te_len, te_off, n_loci, n_reads;
    int64_t n_bases;
    const uint32_t *seq2, *nmask;
    const int64_t *read_off;
    const int32_t *read_len;
    const uint32_t *read_hash;
    const int32_t *locus_read_begin;
    const int64_t *contig_off;
    const int32_t *contig_len, *te_start, *te_end;
} orc_batch_t;

typedef struct orc_result {
    int32_t *cov2x;
    double *af;
    int32_t *depth;
    orc_aln_t *aln;
    int64_t aln_cap;
    uint32_t *cigar;
    int64_t cigar_cap;
    int64_t n_aln, n_cigar;
    int64_t dp_cells, n_minimizers, n_anchors, n_dp_tasks, n_aln_blocks;
    float ms_stage[8];
} orc_result_t;

/* whole path, n_threads OpenMP threads over loci (0 = all). first_locus/n_run select a subrange (bounded samples). */
int orc_af_run(const orc_batch_t *b, orc_result_t *r, int n_threads, int first_locus, int n_run);

/* debugging taps for stage-level parity: map ONE read against ONE contig strand */
typedef struct orc_dbg {
    /* query minimizers after mm_seed_mz_flt */
    int64_t n_mz; uint64_t *mz_x, *mz_y;
    /* anchors after collect_seed_hits (sorted) */
    int64_t n_a; orc128_t *a;
    /* chains after mg_lchain_dp (+rmq re-chain): u[] and compacted anchors */
    int32_t n_u; uint64_t *u; int64_t n_ca; orc128_t *ca;
    int32_t mid_occ, rechained;
    /* regs before alignment (after chain_post): as,cnt,score,parent,rs,re,qs,qe,rev,hash */
    int32_t n_regs0; int32_t *regs0; /* 10 ints per reg */
} orc_dbg_t;
void orc_dbg_free(orc_dbg_t *d);
/* returns #alignment records written to aln (cigars appended to cigar at *n_cigar) */
int orc_map_one(const orc_opt_t *opt, const uint8_t *contig, int32_t clen,
                const uint8_t *read, int32_t qlen, uint32_t name_hash,
                orc_aln_t *aln, int aln_cap, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cigar,
                int64_t *dp_cells, int64_t *n_dp_tasks, orc_dbg_t *dbg);

/* [REF]+[UP] depth/median/AF (TELR_te.py:518-575, 810-835, 841-884; samtools depth -aa region semantics) */
int32_t orc_median2x(const int32_t *depth, int32_t L, int32_t S, int32_t E); /* 2*median over region "c:S-E"; -1 if empty */
void orc_cov_af(const int32_t *depth_fw, const int32_t *depth_rc, int32_t L, int32_t te_s, int32_t te_e,
                int32_t flank_len, int32_t flank_off, int32_t te_len, int32_t te_off,
                int32_t cov2x[8], double *af);

uint32_t orc_name_hash(const char *name);
int orc_pack_seq(const char *ascii, int32_t len, int64_t dst_off, uint32_t *seq2, uint32_t *nmask);

#ifdef __cplusplus
}
#endif
#endif
