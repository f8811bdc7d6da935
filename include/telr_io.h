/*
 * telr_io.h — C ABI of the host-side I/O that sits either side of the stage-4 device path:
 *
 *   read gather   replaces prep_assembly_inputs(read_type="all") + extract_reads (TELR_assembly.py:384-471): per locus
 *                 pysam.AlignmentFile.fetch over the +-1 kb breakpoint window (BAI query, no whole-file inflate), one
 *                 streaming pass over the raw reads (FASTA/FASTQ, plain or gzip) instead of `seqtk subseq` + SeqIO.index +
 *                 csplit, packed straight into the telr_af_batch layout (include/telr_af.h).
 *   BAM emission  replaces `samtools view -bS | samtools sort | samtools index` of realignment() (TELR_te.py:507-512) and of
 *                 the polishing alignment (TELR_assembly.py:199-212): coordinate-sorted BGZF BAM + BAI from alignment records.
 *
 * Host only (zlib); no CUDA.  Functions return 0 or a negative TELR_IO_E* code; telr_io_strerror names them.
 */
#ifndef TELR_IO_H
#define TELR_IO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TELR_IO_OK        0
#define TELR_IO_EINVAL   -1
#define TELR_IO_ENOMEM   -2
#define TELR_IO_EIO      -3   /* open/read/write failed, truncated or corrupt file */
#define TELR_IO_EFORMAT  -4   /* not BGZF/BAM/BAI/FASTA/FASTQ */
#define TELR_IO_ENOINDEX -5   /* no .bai next to the BAM (pysam.fetch needs one too) */
#define TELR_IO_ECONTIG  -6   /* unknown reference name (pysam raises ValueError) */
#define TELR_IO_EMISSING -7   /* a read named in the BAM is absent from the raw reads (SeqIO.index(...).get_raw raises KeyError) */

const char *telr_io_strerror(int code);

/* ---- BAM + BAI ---- */
typedef struct telr_bam telr_bam;
int  telr_bam_open(const char *bam_path, telr_bam **out);      /* header + <bam_path>.bai (or <stem>.bai) */
void telr_bam_close(telr_bam *b);
int  telr_bam_n_ref(const telr_bam *b);
const char *telr_bam_ref_name(const telr_bam *b, int tid);
int64_t telr_bam_ref_len(const telr_bam *b, int tid);
int  telr_bam_tid(const telr_bam *b, const char *chrom);       /* -1 when unknown */
/* pysam.AlignmentFile.fetch(chrom, beg, end): every record (primary, secondary, supplementary alike) whose reference interval
 * [pos, end) overlaps the half-open window; records without a reference span count as length 1.  Names come back as
 * consecutive NUL-terminated strings in coordinate order; the buffer belongs to `b` and lives until the next call.
 * Returns the number of names or a negative code. */
int64_t telr_bam_fetch(telr_bam *b, int tid, int64_t beg, int64_t end, const char **names, int64_t *names_bytes);
int64_t telr_bam_blocks_inflated(const telr_bam *b);           /* BGZF blocks decompressed so far (tests: no whole-file scan) */
/* `samtools index`: writes the BAI of a coordinate-sorted BAM */
int  telr_bam_index_build(const char *bam_path, const char *bai_path);

/* ---- read gather + pack (rows a2 / f3) ---- */
typedef struct telr_gather_in {
    const char *bam_path;
    const char *raw_reads_path;       /* FASTA or FASTQ, plain or gzip */
    int32_t n_loci;
    const char *const *chrom;         /* [n_loci] */
    const int64_t *win_beg, *win_end; /* [n_loci] half-open window on chrom */
    const char *const *contig_seq;    /* [n_loci] ASCII contig, or NULL: the locus has no assembly (reads still counted and written) */
    const int32_t *contig_len;        /* [n_loci] */
    const char *reads_dir;            /* NULL: do not write <reads_dir>/<locus_name>.reads.fa */
    const char *const *locus_name;    /* [n_loci], used with reads_dir */
    int32_t n_threads;                /* packing / file-writing threads (<= 0: hardware concurrency) */
} telr_gather_in;

typedef struct telr_gather_out {
    /* the packed batch over the LIVE loci (contig present and non-empty), locus by locus: contig, then its reads */
    int32_t n_live, n_reads;
    int64_t n_bases;
    uint32_t *seq2, *nmask;
    int64_t *read_off; int32_t *read_len; uint32_t *read_hash;
    int32_t *locus_read_begin;        /* [n_live + 1] */
    int64_t *contig_off; int32_t *contig_len;   /* [n_live] */
    int32_t *live_index;              /* [n_live] index of the live locus in the input order */
    int32_t *n_names;                 /* [n_loci] unique read names per window (column 15 of <vcf_parsed>.new) */
    /* what it cost */
    double  t_bam_s, t_reads_s, t_pack_s, t_write_s;
    int64_t reads_scanned, bases_scanned, unique_reads, bgzf_blocks;
    char    err[256];                 /* detail of the last error (e.g. the missing read name) */
} telr_gather_out;

int  telr_gather_run(const telr_gather_in *in, telr_gather_out *out);
void telr_gather_free(telr_gather_out *out);

/* ---- BAM emission (row f2; used by f1 as well) ---- */
typedef struct telr_sam_rec {         /* one SAM line of `minimap2 -a` */
    const char *qname;
    int32_t flag, tid, pos, mapq;     /* pos 0-based; tid -1 + flag 4 = unmapped */
    const uint32_t *cigar; int32_t n_cigar;   /* BAM encoding len<<4|op (M I D N S H P = X) */
    const char *seq; int32_t l_seq;   /* ASCII, already in the orientation SAM prints; NULL/0 = '*' */
    const uint8_t *aux; int32_t l_aux;        /* optional fields, BAM-encoded (tag[2] type value ...), appended verbatim */
} telr_sam_rec;
/* Writes `path` (BGZF BAM, records sorted by (tid, pos), unmapped last, stable) and `path`.bai. */
int telr_bam_write_sorted(const char *path, int32_t n_ref, const char *const *ref_name, const int32_t *ref_len,
                          const char *header_text_extra, int64_t n_rec, const telr_sam_rec *recs, int32_t level);

#ifdef __cplusplus
}
#endif
#endif /* TELR_IO_H */
