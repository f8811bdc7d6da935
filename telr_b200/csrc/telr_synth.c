/*
 * telr_synth.c — deterministic synthetic stage-4 batches (SURVEY.md §8d generator), host only.
 *
 * Bench/test tooling, not part of the stage-4 path: it plays the role of stages 1-3 of TELR
 * (reads in the +-1 kb breakpoint window TELR_assembly.py:385-410, polished contig
 * TELR_assembly.py:89-98, TE annotation TELR_te.py:207-235) by drawing them from a model.
 * Every locus has its own RNG stream keyed by (seed, locus index), so any shard of loci can be
 * generated independently and identically (multi-GPU sharding by locus).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct telr_synth_cfg {
    uint64_t seed;
    int32_t n_loci_total;      /* loci in the whole job (locus ids are global)            */
    double depth;              /* sequencing depth                                         */
    double mean_len, sigma_len;/* read length lognormal (mean, sigma of log)               */
    int32_t min_len, max_len;
    int32_t flank_lo, flank_hi;/* contig flank length ~ U[lo, hi]                          */
    int32_t te_min, te_max;    /* clip of TE family lengths                                */
    double te_median, te_sigma;
    double p_sub, p_ins, p_del, hp_mult;   /* read error model                             */
    double p_polish;           /* contig polishing error rate                              */
    double p_n;                /* probability a read base is emitted as N                  */
    int32_t n_families;
} telr_synth_cfg;

typedef struct telr_synth_out {
    int32_t n_loci, n_reads;
    int64_t n_bases;
    uint32_t *seq2, *nmask;
    int64_t *read_off;
    int32_t *read_len;
    uint32_t *read_hash;
    int32_t *locus_read_begin;
    int64_t *contig_off;
    int32_t *contig_len, *te_start, *te_end;
    float *truth_af;
    int32_t *read_truth;       /* 1 = TE-bearing read */
    int32_t *read_origin;      /* [n_reads][4]: start on its haplotype, source length, reverse-complemented, context length (haplotype coordinate of the breakpoint) */
} telr_synth_out;

typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t splitmix(uint64_t *x)
{
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t a, uint64_t b)
{
    uint64_t x = a * 0x9e3779b97f4a7c15ULL ^ (b + 0x632be59bd9b4e019ULL) * 0xd1342543de82ef95ULL;
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix(&x);
}
static inline uint64_t rng_next(rng_t *r)
{
    uint64_t *s = r->s, result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static inline double rng_u(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }
static double rng_normal(rng_t *r)
{
    double u1 = rng_u(r), u2 = rng_u(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static int rng_poisson(rng_t *r, double lam)
{
    if (lam > 50) {
        int v = (int)floor(lam + sqrt(lam) * rng_normal(r) + 0.5);
        return v < 0 ? 0 : v;
    } else {
        double L = exp(-lam), p = 1.0;
        int k = 0;
        do { ++k; p *= rng_u(r); } while (p > L);
        return k - 1;
    }
}
static inline uint8_t rng_base(rng_t *r)
{
    double u = rng_u(r);    /* P(A,C,G,T) = (.29,.21,.21,.29) */
    return u < 0.29 ? 0 : u < 0.50 ? 1 : u < 0.71 ? 2 : 3;
}

typedef struct { uint8_t *seq; int32_t len; } te_t;

static void gen_families(const telr_synth_cfg *c, te_t *fam)
{
    for (int f = 0; f < c->n_families; ++f) {
        rng_t r;
        rng_seed(&r, c->seed ^ 0x7e1eULL, (uint64_t)f);
        double ln = log(c->te_median) + c->te_sigma * rng_normal(&r);
        int32_t len = (int32_t)exp(ln);
        if (len < c->te_min) len = c->te_min;
        if (len > c->te_max) len = c->te_max;
        uint8_t *s = (uint8_t *)malloc((size_t)len);
        for (int i = 0; i < len; ++i) s[i] = rng_base(&r);
        double kind = rng_u(&r);
        if (kind < 0.25 && len >= 1000) {               /* LTR type: identical 350 bp direct terminal repeats */
            memcpy(s + len - 350, s, 350);
        } else if (kind < 0.35 && len >= 1000) {        /* internal 40 x 7-mer tandem array */
            int at = len / 3;
            for (int i = 7; i < 280; ++i) s[at + i] = s[at + i % 7];
        }
        fam[f].seq = s, fam[f].len = len;
    }
}

typedef struct {
    uint8_t *b; int64_t n, m;
} bytes_t;
static inline void bpush(bytes_t *v, uint8_t c)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 4096; v->b = (uint8_t *)realloc(v->b, (size_t)v->m); }
    v->b[v->n++] = c;
}

/* apply the read error model to src[0..n) (already oriented), append to dst */
static void mutate(rng_t *r, const telr_synth_cfg *c, const uint8_t *src, int32_t n, bytes_t *dst)
{
    int run = 0;
    for (int32_t i = 0; i < n; ++i) {
        uint8_t b = src[i];
        run = (i > 0 && src[i - 1] == b) ? run + 1 : 1;
        double m = run >= 4 ? c->hp_mult : 1.0;
        double u = rng_u(r);
        if (u < c->p_del * m) continue;
        u = rng_u(r);
        if (u < c->p_sub) b = (uint8_t)((b + 1 + rng_below(r, 3)) & 3);
        if (c->p_n > 0 && rng_u(r) < c->p_n) b = 4;
        bpush(dst, b);
        u = rng_u(r);
        if (u < c->p_ins * m) {
            do {
                bpush(dst, run >= 4 ? src[i] : rng_base(r));
            } while (rng_u(r) < 0.3);
        }
    }
}

typedef struct {
    bytes_t seq;             /* all reads of the locus concatenated (nt4) then nothing else */
    int32_t *rlen, *rtruth, *rorig; int32_t n_reads, m_reads;
    uint8_t *contig; int32_t clen, te_s, te_e;
    float af;
    uint32_t *p2, *pn; int64_t pbases;   /* the locus packed on its own (contig, then reads; 64-base aligned): keeps the peak footprint at 0.375 B/base */
} locus_t;

static void pack_into(const uint8_t *s, int32_t len, int64_t off, uint32_t *seq2, uint32_t *nmask);

static void gen_locus(const telr_synth_cfg *c, const te_t *fam, int32_t gid, locus_t *o)
{
    rng_t r;
    rng_seed(&r, c->seed, (uint64_t)gid);
    const te_t *te = &fam[rng_below(&r, (uint32_t)c->n_families)];
    int te_rev = rng_below(&r, 2);
    int32_t fl = c->flank_lo + (int32_t)rng_below(&r, (uint32_t)(c->flank_hi - c->flank_lo + 1));
    int32_t fr = c->flank_lo + (int32_t)rng_below(&r, (uint32_t)(c->flank_hi - c->flank_lo + 1));
    static const float afs[4] = {0.25f, 0.5f, 0.75f, 1.0f};
    o->af = afs[rng_below(&r, 4)];
    int32_t ctx = c->max_len + 1000;         /* genome context either side of the breakpoint */
    int32_t T = te->len;
    uint8_t *alt = (uint8_t *)malloc((size_t)(2 * ctx + T));
    uint8_t *ref = (uint8_t *)malloc((size_t)(2 * ctx));
    for (int32_t i = 0; i < ctx; ++i) alt[i] = rng_base(&r);
    for (int32_t i = 0; i < T; ++i) alt[ctx + i] = te_rev ? (uint8_t)(3 - te->seq[T - 1 - i]) : te->seq[i];
    for (int32_t i = 0; i < ctx; ++i) alt[ctx + T + i] = rng_base(&r);
    memcpy(ref, alt, (size_t)ctx);
    memcpy(ref + ctx, alt + ctx + T, (size_t)ctx);
    /* polished contig = flankL + TE + flankR with a few polishing errors */
    {
        bytes_t cb = {0, 0, 0};
        int32_t s0 = ctx - fl, e0 = ctx + T + fr, te_s = -1, te_e = -1;
        for (int32_t i = s0; i < e0; ++i) {
            if (i == ctx) te_s = (int32_t)cb.n;
            if (i == ctx + T) te_e = (int32_t)cb.n;
            double u = rng_u(&r);
            if (u < c->p_polish / 3) continue;                                   /* deletion */
            if (u < 2 * c->p_polish / 3) { bpush(&cb, alt[i]); bpush(&cb, rng_base(&r)); continue; }
            if (u < c->p_polish) { bpush(&cb, (uint8_t)((alt[i] + 1 + rng_below(&r, 3)) & 3)); continue; }
            bpush(&cb, alt[i]);
        }
        if (te_e < 0) te_e = (int32_t)cb.n;
        o->contig = cb.b, o->clen = (int32_t)cb.n, o->te_s = te_s, o->te_e = te_e;
    }
    /* reads */
    double lam = c->depth * (c->mean_len + 2000.0) / c->mean_len;
    int32_t nr = rng_poisson(&r, lam);
    if (nr < 1) nr = 1;
    o->n_reads = 0; o->m_reads = nr;
    o->rlen = (int32_t *)malloc((size_t)nr * 4); o->rtruth = (int32_t *)malloc((size_t)nr * 4); o->rorig = (int32_t *)malloc((size_t)nr * 16);
    memset(&o->seq, 0, sizeof(o->seq));
    double mu_ln = log(c->mean_len) - 0.5 * c->sigma_len * c->sigma_len;
    uint8_t *tmp = (uint8_t *)malloc((size_t)c->max_len + 16);
    for (int32_t k = 0; k < nr; ++k) {
        int32_t len = (int32_t)exp(mu_ln + c->sigma_len * rng_normal(&r));
        if (len < c->min_len) len = c->min_len;
        if (len > c->max_len) len = c->max_len;
        int has_te = rng_u(&r) < o->af;
        const uint8_t *hap = has_te ? alt : ref;
        int32_t hlen = has_te ? 2 * ctx + T : 2 * ctx, start;
        if (!has_te) {
            /* read [s, s+len) must intersect [ctx-1000, ctx+1000) */
            int32_t lo = ctx - 1000 - len + 1, hi = ctx + 1000 - 1;
            start = lo + (int32_t)rng_below(&r, (uint32_t)(hi - lo + 1));
        } else {
            /* must intersect the 1 kb of flank adjacent to either TE end */
            for (;;) {
                int32_t lo = ctx - 1000 - len + 1, hi = ctx + T + 1000 - 1;
                start = lo + (int32_t)rng_below(&r, (uint32_t)(hi - lo + 1));
                if (start < ctx || start + len > ctx + T) break;      /* not entirely inside the TE */
            }
        }
        if (start < 0) start = 0;
        if (start + len > hlen) len = hlen - start;
        int rc = rng_below(&r, 2);
        if (rc) for (int32_t i = 0; i < len; ++i) tmp[i] = (uint8_t)(3 - hap[start + len - 1 - i]);
        else memcpy(tmp, hap + start, (size_t)len);
        int64_t n0 = o->seq.n;
        mutate(&r, c, tmp, len, &o->seq);
        if (o->seq.n == n0) bpush(&o->seq, 0);
        o->rlen[o->n_reads] = (int32_t)(o->seq.n - n0);
        o->rtruth[o->n_reads] = has_te;
        o->rorig[4 * o->n_reads] = start; o->rorig[4 * o->n_reads + 1] = len; o->rorig[4 * o->n_reads + 2] = rc; o->rorig[4 * o->n_reads + 3] = ctx;
        o->n_reads++;
    }
    free(tmp); free(alt); free(ref);
    {   /* pack now and drop the byte arrays */
        int64_t nb = ((int64_t)o->clen + 63) / 64 * 64;
        for (int32_t k = 0; k < o->n_reads; ++k) nb += ((int64_t)o->rlen[k] + 63) / 64 * 64;
        o->pbases = nb;
        o->p2 = (uint32_t *)calloc((size_t)(nb / 16) + 4, 4);
        o->pn = (uint32_t *)calloc((size_t)(nb / 32) + 4, 4);
        int64_t o2 = 0, so = 0;
        pack_into(o->contig, o->clen, o2, o->p2, o->pn);
        o2 += ((int64_t)o->clen + 63) / 64 * 64;
        for (int32_t k = 0; k < o->n_reads; ++k) {
            pack_into(o->seq.b + so, o->rlen[k], o2, o->p2, o->pn);
            so += o->rlen[k];
            o2 += ((int64_t)o->rlen[k] + 63) / 64 * 64;
        }
        free(o->contig); free(o->seq.b);
        o->contig = NULL; memset(&o->seq, 0, sizeof(o->seq));
    }
}

static void pack_into(const uint8_t *s, int32_t len, int64_t off, uint32_t *seq2, uint32_t *nmask)
{
    for (int32_t i = 0; i < len; ++i) {
        int64_t p = off + i;
        if (s[i] < 4) seq2[p >> 4] |= (uint32_t)s[i] << (2 * (p & 15));
        else nmask[p >> 5] |= 1u << (p & 31);
    }
}

static uint32_t x31(const char *s)
{
    uint32_t h = (uint32_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    return h;
}

void telr_synth_default(telr_synth_cfg *c)
{
    memset(c, 0, sizeof(*c));
    c->seed = 20221101; c->n_loci_total = 3000; c->depth = 50; c->mean_len = 10000; c->sigma_len = 0.45;
    c->min_len = 1000; c->max_len = 60000; c->flank_lo = 1500; c->flank_hi = 3500;
    c->te_min = 600; c->te_max = 9000; c->te_median = 4700; c->te_sigma = 0.6;
    c->p_sub = 0.04; c->p_ins = 0.03; c->p_del = 0.04; c->hp_mult = 2.0; c->p_polish = 0.002; c->p_n = 0.0;
    c->n_families = 64;
}

/* read names are "L%06d_R%04d" (global locus id, read index within locus) */
static int synth_generate(const telr_synth_cfg *c, int32_t first_locus, const int32_t *ids, int32_t n_loci, telr_synth_out *out);

int telr_synth_generate(const telr_synth_cfg *c, int32_t first_locus, int32_t n_loci, telr_synth_out *out)
{
    return synth_generate(c, first_locus, NULL, n_loci, out);
}

/* the loci ids[0..n) (global locus ids, any order): what one device of a multi-GPU partition owns */
int telr_synth_generate_list(const telr_synth_cfg *c, const int32_t *ids, int32_t n_loci, telr_synth_out *out)
{
    return synth_generate(c, 0, ids, n_loci, out);
}

static int synth_generate(const telr_synth_cfg *c, int32_t first_locus, const int32_t *ids, int32_t n_loci, telr_synth_out *out)
{
    te_t *fam = (te_t *)calloc((size_t)c->n_families, sizeof(te_t));
    locus_t *L = (locus_t *)calloc((size_t)(n_loci > 0 ? n_loci : 1), sizeof(locus_t));
    gen_families(c, fam);
#define GID(l) (ids ? ids[l] : first_locus + (l))
#pragma omp parallel for schedule(dynamic, 4)
    for (int32_t l = 0; l < n_loci; ++l) gen_locus(c, fam, GID(l), &L[l]);
    memset(out, 0, sizeof(*out));
    out->n_loci = n_loci;
    int64_t nb = 0; int32_t nr = 0;
    for (int32_t l = 0; l < n_loci; ++l) {
        nb += ((int64_t)L[l].clen + 63) / 64 * 64;
        for (int32_t k = 0; k < L[l].n_reads; ++k) nb += ((int64_t)L[l].rlen[k] + 63) / 64 * 64;
        nr += L[l].n_reads;
    }
    out->n_reads = nr; out->n_bases = nb;
    out->seq2 = (uint32_t *)malloc(((size_t)(nb / 16) + 4) * 4);
    out->nmask = (uint32_t *)malloc(((size_t)(nb / 32) + 4) * 4);
    out->read_off = (int64_t *)malloc((size_t)(nr + 1) * 8);
    out->read_len = (int32_t *)malloc((size_t)(nr + 1) * 4);
    out->read_hash = (uint32_t *)malloc((size_t)(nr + 1) * 4);
    out->read_truth = (int32_t *)malloc((size_t)(nr + 1) * 4);
    out->read_origin = (int32_t *)malloc((size_t)(nr + 1) * 16);
    out->locus_read_begin = (int32_t *)malloc((size_t)(n_loci + 1) * 4);
    out->contig_off = (int64_t *)malloc((size_t)(n_loci + 1) * 8);
    out->contig_len = (int32_t *)malloc((size_t)(n_loci + 1) * 4);
    out->te_start = (int32_t *)malloc((size_t)(n_loci + 1) * 4);
    out->te_end = (int32_t *)malloc((size_t)(n_loci + 1) * 4);
    out->truth_af = (float *)malloc((size_t)(n_loci + 1) * 4);
    if (!out->seq2 || !out->nmask) return -2;
    /* offsets first (serial), then pack in parallel */
    int64_t *loff = (int64_t *)malloc((size_t)(n_loci + 1) * 8);
    int64_t off = 0; int32_t ri = 0;
    for (int32_t l = 0; l < n_loci; ++l) {
        loff[l] = off;
        out->locus_read_begin[l] = ri;
        out->contig_off[l] = off; out->contig_len[l] = L[l].clen;
        out->te_start[l] = L[l].te_s; out->te_end[l] = L[l].te_e; out->truth_af[l] = L[l].af;
        off += ((int64_t)L[l].clen + 63) / 64 * 64;
        for (int32_t k = 0; k < L[l].n_reads; ++k) {
            char name[64];
            snprintf(name, sizeof(name), "L%06d_R%04d", GID(l), k);
            out->read_off[ri] = off; out->read_len[ri] = L[l].rlen[k];
            out->read_hash[ri] = x31(name); out->read_truth[ri] = L[l].rtruth[k];
            memcpy(out->read_origin + 4 * (size_t)ri, L[l].rorig + 4 * k, 16);
            off += ((int64_t)L[l].rlen[k] + 63) / 64 * 64;
            ++ri;
        }
    }
    out->locus_read_begin[n_loci] = ri;
#pragma omp parallel for schedule(dynamic, 4)
    for (int32_t l = 0; l < n_loci; ++l) {
        memcpy(out->seq2 + loff[l] / 16, L[l].p2, (size_t)(L[l].pbases / 16) * 4);
        memcpy(out->nmask + loff[l] / 32, L[l].pn, (size_t)(L[l].pbases / 32) * 4);
        free(L[l].p2); free(L[l].pn); free(L[l].rlen); free(L[l].rtruth); free(L[l].rorig);
    }
#undef GID
    free(loff); free(L);
    for (int f = 0; f < c->n_families; ++f) free(fam[f].seq);
    free(fam);
    return 0;
}

void telr_synth_free(telr_synth_out *o)
{
    free(o->seq2); free(o->nmask); free(o->read_off); free(o->read_len); free(o->read_hash); free(o->read_truth); free(o->read_origin);
    free(o->locus_read_begin); free(o->contig_off); free(o->contig_len); free(o->te_start); free(o->te_end);
    free(o->truth_af);
    memset(o, 0, sizeof(*o));
}
