/*
 * telr_af.h — C ABI of the B200-native TELR stage-4 allele-frequency path.
 *
 * The reference (bergmanlab/TELR) exposes NO FFI for this path: the boundary
 * is the Python function get_af() (src/telr/TELR_te.py:578-838) which shells
 * out per locus to `minimap2 -a -x <preset>` (TELR_te.py:495-515) and
 * `samtools depth -aa -r` (TELR_te.py:870-884) and then does ~80 lines of
 * Python arithmetic (TELR_te.py:564-575, 757-835).  The entry points below are
 * what a binding for that body binds instead; telr_b200/stage4.py is the
 * ctypes host that keeps get_af()'s signature, return dict and side files.
 *
 * Conventions: extern "C", plain pointers and sizes, no torch types.  Every
 * function returns 0 on success or a negative TELR_E* code and never throws.
 * A ctx belongs to one (device, host thread); calls on one ctx are serialised
 * by the caller.  Functions are stream-ordered on the ctx's own stream.
 */
#ifndef TELR_AF_H
#define TELR_AF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TELR_OK            0
#define TELR_EINVAL       -1   /* bad argument / malformed batch            */
#define TELR_ENOMEM       -2   /* host or device allocation failed          */
#define TELR_ECUDA        -3   /* CUDA runtime error (see telr_af_last_cuda) */
#define TELR_ENODEV       -4   /* no CUDA device / wrong architecture       */
#define TELR_ECAP         -5   /* an internal capacity was exceeded         */
#define TELR_EUNSUPPORTED -6   /* valid input outside the implemented scope */

/* minimap2 presets reachable from TELR (TELR_te.py:595-598) + one extension */
#define TELR_PRESET_MAP_ONT  0
#define TELR_PRESET_MAP_PB   1
#define TELR_PRESET_MAP_HIFI 2  /* extension: TELR itself never selects it */

/* Sequence layout (host and device alike):
 *   seq2  : 2-bit bases A=0 C=1 G=2 T=3, 16 per uint32, base i of a sequence
 *           that starts at base offset `off` is (seq2[(off+i)>>4] >> 2*((off+i)&15)) & 3
 *   nmask : 1 bit per base, 32 per uint32; 1 = ambiguous (seq2 holds 0 there)
 *   every sequence starts at a base offset that is a multiple of 64.
 */
typedef struct telr_af_batch {
    int32_t preset;               /* TELR_PRESET_*                                   */
    int32_t flank_len;            /* --af_flank_interval  (TELR_input.py:217-218)    */
    int32_t flank_off;            /* --af_flank_offset    (TELR_input.py:226-227)    */
    int32_t te_len;               /* --af_te_interval     (TELR_input.py:234-235)    */
    int32_t te_off;               /* --af_te_offset       (TELR_input.py:242-243)    */
    int32_t n_loci;
    int32_t n_reads;
    int64_t n_bases;              /* total bases covered by seq2 (multiple of 64)    */
    const uint32_t *seq2;         /* [n_bases/16]                                    */
    const uint32_t *nmask;        /* [n_bases/32]                                    */
    const int64_t  *read_off;     /* [n_reads]   base offset of read r               */
    const int32_t  *read_len;     /* [n_reads]                                       */
    const uint32_t *read_hash;    /* [n_reads]   X31 hash of the read name           */
    const int32_t  *locus_read_begin; /* [n_loci+1] CSR: reads of locus l            */
    const int64_t  *contig_off;   /* [n_loci]    base offset of the fw contig        */
    const int32_t  *contig_len;   /* [n_loci]    0 = locus has no contig (skipped)   */
    const int32_t  *te_start;     /* [n_loci]    BED start on the fw contig; <0 = no annotation */
    const int32_t  *te_end;       /* [n_loci]                                        */
} telr_af_batch;

/* One alignment record = one minimap2 "reg" that would be written to SAM. */
typedef struct telr_aln {
    int32_t read;      /* global read index                                   */
    int32_t strand;    /* 0: aligned to the fw contig, 1: to the rc contig    */
    int32_t rs, re;    /* 0-based half-open target interval                   */
    int32_t qs, qe;    /* 0-based half-open query interval (read orientation) */
    int32_t rev;       /* read aligned as its reverse complement              */
    int32_t flag;      /* SAM-ish: 0x100 secondary, 0x800 supplementary, 0x10 */
    int32_t dp_max;    /* minimap2 ms:i                                       */
    int32_t mlen, blen;
    int32_t n_cigar;
    int64_t cigar_off; /* into telr_af_result.cigar                           */
    int32_t mapq;      /* SAM MAPQ (mm_set_mapq)                              */
    int32_t dp_score;  /* AS:i                                                */
    int32_t cnt;       /* cm:i  minimizers on the chain                       */
    int32_t score;     /* s1:i  chaining score                                */
    int32_t subsc;     /* s2:i  chaining score of the best secondary          */
    int32_t n_ambi;    /* nn:i  ambiguous bases in the alignment              */
    int32_t inv;       /* tp:A:I  inversion piece                             */
    int32_t n_sub;     /* suboptimal hits counted by mm_set_parent            */
} telr_aln;

typedef struct telr_af_result {
    /* required, caller allocated */
    int32_t *cov2x;     /* [n_loci][8] 2*median: te5p,te3p,flank5p,flank3p (fw), same (rc); -1 = None, -2 = locus skipped */
    double  *af;        /* [n_loci] AF before clamp/round; NaN = None          */
    /* optional (NULL to skip) */
    int32_t *depth;     /* [sum over loci of 2*contig_len]: fw depth then rc depth per locus, loci in order */
    telr_aln *aln;      /* [aln_cap]                                            */
    int64_t  aln_cap;
    uint32_t *cigar;    /* [cigar_cap] BAM-style len<<4|op, op: 0=M 1=I 2=D     */
    int64_t  cigar_cap;
    int64_t  n_aln;     /* out */
    int64_t  n_cigar;   /* out */
    /* statistics (out) */
    int64_t  dp_cells;      /* algorithmic DP cells             */
    int64_t  n_minimizers;
    int64_t  n_anchors;
    int64_t  n_dp_tasks;
    int64_t  n_aln_blocks;  /* M-blocks fed to the depth kernel                 */
    float    ms_stage[8];   /* device ms: sketch, seed+chain, align-plan, dp, traceback, finalize, depth, cov+af */
} telr_af_result;

typedef struct telr_af_ctx telr_af_ctx;

/* replaces: process spawn of minimap2/samtools in realignment() TELR_te.py:495-515 */
int telr_af_create(telr_af_ctx **ctx, int device, size_t workspace_bytes);
int telr_af_destroy(telr_af_ctx *ctx);

/* Whole stage-4 body on HOST buffers (H2D + kernels + D2H inside).
 * replaces: TELR_te.py:640-755 (Pool.map(realignment) x2 + the two serial depth loops) */
int telr_af_run(telr_af_ctx *ctx, const telr_af_batch *host_batch, telr_af_result *host_result);

/* Same, but every pointer in batch and the required outputs of result are DEVICE pointers
 * (inputs already resident in HBM); optional outputs must be NULL. */
int telr_af_run_device(telr_af_ctx *ctx, const telr_af_batch *dev_batch, telr_af_result *dev_result);

/* Stage entry points (host buffers) used by the parity tests. */
/* replaces mm_sketch() behind TELR_te.py:505.  Writes minimizers of every sequence
 * (x = hash<<8|span, y = pos<<1|strand), seq i at [mz_off[i], mz_off[i+1]).  */
int telr_af_sketch(telr_af_ctx *ctx, const uint32_t *seq2, const uint32_t *nmask, int64_t n_bases,
                   int32_t n_seq, const int64_t *seq_off, const int32_t *seq_len,
                   int32_t w, int32_t k, int32_t hpc,
                   uint64_t *mz_x, uint64_t *mz_y, int64_t mz_cap, int64_t *mz_off);

/* replaces `samtools depth -aa` + get_median_cov/get_te_cov/get_flank_cov/AF block
 * (TELR_te.py:841-884, 518-550, 564-575, 810-835) given M-blocks (locus, strand, start, len). */
int telr_af_depth_af(telr_af_ctx *ctx, int32_t n_loci, const int32_t *contig_len,
                     const int32_t *te_start, const int32_t *te_end,
                     int32_t flank_len, int32_t flank_off, int32_t te_len, int32_t te_off,
                     int64_t n_blocks, const int32_t *blk_locus_strand, const int32_t *blk_start,
                     const int32_t *blk_len,
                     int32_t *depth, int32_t *cov2x, double *af);

/* Extension DP alone (ksw_extd2_sse restatement): one task per (query,target) pair, nt4 bytes. */
typedef struct telr_dp_task {
    int64_t q_off, t_off;   /* into the byte arrays */
    int32_t qlen, tlen;
    int32_t w, zdrop, end_bonus, flag;   /* flag: TELR_KSW_* */
} telr_dp_task;
typedef struct telr_dp_out {
    int32_t max, max_q, max_t, mqe, mqe_t, mte, mte_q, score, zdropped, reach_end;
    int32_t n_cigar; int64_t cigar_off; int64_t cells;
} telr_dp_out;
#define TELR_KSW_EXTZ_ONLY  0x40
#define TELR_KSW_RIGHT      0x02
#define TELR_KSW_REV_CIGAR  0x80
#define TELR_KSW_APPROX_MAX 0x08
int telr_af_dp(telr_af_ctx *ctx, int32_t preset, int32_t n_tasks, const telr_dp_task *tasks,
               const uint8_t *qseq, int64_t qbytes, const uint8_t *tseq, int64_t tbytes,
               telr_dp_out *out, uint32_t *cigar, int64_t cigar_cap);

/* Mapping options on top of the preset, as on the minimap2 command line (0 restores the preset's value):
 *   "bw"       -r NUM   chaining / alignment bandwidth (the polishing alignment of TELR_assembly.py:199-212 runs `-r2k`)
 *   "bw_long"  -r ,NUM  long-join bandwidth
 * Returns TELR_EINVAL for an unknown name or a negative value. */
int telr_af_set_option(telr_af_ctx *ctx, const char *name, int32_t value);

const char *telr_af_strerror(int code);
int telr_af_last_cuda(const telr_af_ctx *ctx);      /* last cudaError_t seen by this ctx */
int telr_af_version(void);
long long telr_af_launch_count(const telr_af_ctx *ctx);  /* kernels launched by this ctx so far */
void *telr_af_stream(const telr_af_ctx *ctx);            /* the ctx's cudaStream_t (for event timing by the caller) */

/* Host helper (no device needed): how telr_af_run cuts a batch into chunks of loci for a budget of read bases per chunk
 * (the reference processes one locus per minimap2 process, TELR_te.py:640-647; here a chunk is what one pass of the device
 * pipeline holds).  cuts[0..n] = first locus of every chunk and n_loci; returns n (>= 0) or a negative TELR_E* code. */
int telr_af_plan_chunks(const int32_t *read_len, const int32_t *locus_read_begin, int32_t n_loci, int64_t budget_bases,
                        int32_t *cuts, int32_t cap);

/* Host helper: ASCII -> 2-bit + N mask (one sequence; dst offsets in bases, multiple of 64). */
int telr_pack_seq(const char *ascii, int32_t len, int64_t dst_off, uint32_t *seq2, uint32_t *nmask);
/* Host helper: X31 string hash of a read name (minimap2 __ac_X31_hash_string) */
uint32_t telr_name_hash(const char *name);

#ifdef __cplusplus
}
#endif
#endif /* TELR_AF_H */
