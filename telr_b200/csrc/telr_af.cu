// telr_af.cu — C ABI (include/telr_af.h) and host orchestration of the stage-4 device pipeline.
//
// Stage order per chunk of loci (all on the ctx's stream):
//   k_unpack_contigs -> k_sketch_tiles -> scan -> k_sketch_compact (map-pb: k_sketch<COUNT> -> scan -> k_sketch<WRITE>) -> k_self_count
//   -> k_chain<count> -> scans -> k_chain<fill> -> k_work_flags -> scan -> k_work_scatter -> k_align -> k_depth_af
// There is no CPU execution path: every entry point fails with TELR_ENODEV when no sm_100 device is present.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/telr_af.h"
#include "k_sketch.cuh"
#include "k_misc.cuh"
#include "k_chain.cuh"
#include "k_align.cuh"
#include "k_depth.cuh"

using namespace telr;

#define TELR_VERSION 100

namespace {

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cudaGetLastError(); if (cudaMalloc(&p, bytes) != cudaSuccess) { p = nullptr; return -1; } want = bytes; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() { return (T *)p; }
};

}  // namespace

struct telr_af_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;      // copy_stream: host->device upload of the packed bases, chunk by chunk, under the kernels of earlier chunks
    int last_cuda = 0;
    long long launches = 0;
    size_t ws_limit = 0, total_mem = (size_t)64 << 30;
    int depth_mode = 1;
    int use_fast = 1, use_vec = 1, census = 0;
    int64_t chunk_bases = 0;
    int64_t dir_cap = 8 << 20;
    // device buffers
    DevBuf b_in[10];            // host-variant copies of the batch arrays
    DevBuf b_cov, b_af, b_depth;
    DevBuf b_lrb, b_cboff, b_ctg, b_descs, b_counts, b_mzoff, b_mzx, b_mzy, b_self, b_tabk, b_tabc, b_hpc, b_hpp, b_hpr;
    DevBuf b_pna, b_pread, b_pls, b_paoff, b_prcap, b_proff, b_pnregs, b_pnca, b_anch, b_regs, b_chws, b_alws, b_work;
    DevBuf b_psb, b_psoff, b_pscr, b_pnu, b_pm, b_prep;
    DevBuf b_hpoff, b_hpn;                // compressed step stream: slice per sequence, steps per sequence
    DevBuf b_order, b_wflag, b_woff;      // LPT order of the chunk's problems, work-list filter
    DevBuf b_tfirst, b_tcnt, b_toff, b_tmpx, b_tmpy;     // tile sketch: first tile per sequence, per-tile counts/offsets, per-tile slots
    int sketch_tiles = 1;
    // The alignment stage has two kernels: k_al_fused (one warp owns one problem from start to end) and the role-specialised k_al_queue
    // (SM roles + device-wide task rings; ext_per8 / wide_per8: eighths of the SMs that start in the extension role / the 12-column gap-fill
    // role).  Measured with the offset-form DP loops (profiles/README.md, round 2): k_al_queue is 5.5 % faster on map-ont (gap fills dominate,
    // one loop per SM fits the instruction cache) and 15-25 % slower on map-pb / map-hifi (extensions dominate, the rings add latency), so
    // the default (-1) picks it for map-ont only.  TELR_AL_QUEUE=0/1 forces one of them.
    int al_queue = -1, ext_per8 = 2, wide_per8 = 2;
    int opt_bw = 0, opt_bw_long = 0;     // telr_af_set_option overrides (0 = preset value)
    DevBuf b_qring, b_qstate;
    int al_blocks = AL_BLOCKS_PER_SM;    // resident k_al_fused CTAs per SM this context asks for (fewer leaves room for a second context's kernels)
    DevBuf b_rbytes, b_rboff, b_alwork, b_alctx, b_altask, b_alres, b_alsz, b_aloff, b_cigs, b_pool, b_tlist, b_rc, b_opt, b_idxbig;
    int64_t pool_cap = (int64_t)6144 << 20;
    DevBuf b_blk, b_pblkoff, b_pblkcnt, b_ctr, b_alnout, b_cigout, b_doff, b_big, b_biglock, b_grow, b_lbad;
    int n_big = 8; int64_t big_cap = (int64_t)208 << 20;
    cudaEvent_t ev[10];
};

static void opt_preset(Opt &o, int preset)
{
    memset(&o, 0, sizeof(o));
    o.k = 15; o.w = 10; o.hpc = 0;
    o.a = 2; o.b = 4; o.q = 4; o.e = 2; o.q2 = 24; o.e2 = 1; o.sc_ambi = 1;
    o.zdrop = 400; o.zdrop_inv = 200; o.end_bonus = -1; o.min_dp_max = 80; o.min_ksw_len = 200;
    o.bw = 500; o.bw_long = 20000; o.max_gap = 5000;
    o.max_chain_skip = 25; o.max_chain_iter = 5000; o.min_cnt = 3; o.min_chain_score = 40;
    o.rmq_inner_dist = 1000; o.rmq_size_cap = 100000; o.rmq_rescue_size = 1000; o.rmq_rescue_ratio = 0.1f;
    float gap_scale = 0.8f, skip_scale = 0.0f;
    o.mask_level = 0.5f; o.mask_len = INT32_MAX; o.pri_ratio = 0.8f; o.best_n = 5;
    o.q_occ_frac = 0.01f; o.mid_occ_frac = 2e-4f; o.min_mid_occ = 10; o.max_mid_occ = 1000000;
    o.max_max_occ = 4095; o.occ_dist = 500;
    o.seed_term = wang_hash32(11u); o.max_sw_mat = 100000000LL; o.rank_min_len = 500; o.rank_frac = 0.9f; o.max_clip_ratio = 1.0f;
    if (preset == TELR_PRESET_MAP_PB) { o.hpc = 1; o.k = 19; }
    else if (preset == TELR_PRESET_MAP_HIFI) {
        o.k = 19; o.w = 19; o.max_gap = 10000; o.a = 1; o.b = 4; o.q = 6; o.q2 = 26; o.e = 2; o.e2 = 1;
        o.min_mid_occ = 50; o.max_mid_occ = 500; o.min_dp_max = 200;
    }
    o.chn_pen_gap = (float)(gap_scale * 0.01 * o.k);
    o.chn_pen_skip = (float)(skip_scale * 0.01 * o.k);
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->last_cuda = (int)e_; fprintf(stderr, "[telr_af] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return TELR_ECUDA; } } while (0)
#define ENS(buf, bytes) do { if ((buf).ensure((size_t)(bytes)) != 0) { fprintf(stderr, "[telr_af] device allocation of %zu bytes failed\n", (size_t)(bytes)); return TELR_ENOMEM; } } while (0)

// ------------------------------------------------------------------------------------------------
__global__ void k_build_descs(int n_reads, int n_loci, const int64_t *read_off, const int32_t *read_len,
                              const int64_t *ctg_boff, const int32_t *contig_len, SeqDesc *d)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_reads) { d[i].off = read_off[i]; d[i].len = read_len[i]; d[i].kind = 0; }
    else if (i < n_reads + 2 * n_loci) {
        int j = i - n_reads, s = j >= n_loci, l = s ? j - n_loci : j;
        d[i].off = ctg_boff[l] + (s ? contig_len[l] : 0); d[i].len = contig_len[l]; d[i].kind = 1;
    }
}

__global__ void k_reg_caps(int n, const int32_t *na, int32_t *cap, int32_t *sbytes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int a = na[i];
        cap[i] = a > 0 ? 2 * (a / 3) + 4 : 0;
        sbytes[i] = a > 0 ? (int32_t)(chain_scratch_bytes((size_t)a + 1) + hit_scratch_bytes((size_t)(2 * (a / 3) + 8))) : 0;
    }
}

// Work list of the alignment kernel = the problems that kept a region, longest read first (LPT order for the persistent
// warps).  The order of ALL problems is known before anything runs (read lengths), so the host sorts it once per chunk
// while the sketch kernels execute; the device only filters it: flags -> exclusive scan -> stable scatter.
__global__ void k_work_flags(int n, const int32_t *order, const int32_t *nregs, int32_t *flags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = nregs[order[i]] > 0;
}
__global__ void k_work_scatter(int n, const int32_t *order, const int32_t *flags, const int64_t *offs, int32_t *list, int64_t *count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) list[offs[i]] = order[i];
    if (i == 0) *count = offs[n];
}

__global__ void k_al_sizes(AlignArgs A, int32_t *sizes)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= A.n_work) return;
    const int pidx = A.work_list[w];
    const int qlen = A.read_len[A.prob_read[pidx]], L = A.contig_len[A.prob_ls[pidx] >> 1];
    sizes[w] = (2 * (qlen + L) + 256) + (qlen + L + 16);
}
__global__ void k_al_offsets(AlignArgs A, const int64_t *off)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= A.n_work) return;
    const int pidx = A.work_list[w];
    const int qlen = A.read_len[A.prob_read[pidx]], L = A.contig_len[A.prob_ls[pidx] >> 1];
    A.work[w].cig_off = off[w]; A.work[w].cig_cap = 2 * (qlen + L) + 256;
    A.work[w].ez_off = off[w] + A.work[w].cig_cap; A.work[w].ez_cap = qlen + L + 16;
}

// Kernel (a), uncompressed presets: k_sketch_tiles -> scan of the tile counts -> k_sketch_compact.
// `lens` are the sequence lengths in descriptor order; on return b_mzoff / b_mzx / b_mzy hold the CSR minimizer lists.
// host_overlap() runs on the host while the tile kernel executes (before the one synchronisation of this stage).
template <class F>
static int sketch_tiled(telr_af_ctx *ctx, const SketchArgs &sa, const int32_t *lens, int64_t *n_mz_out, F host_overlap)
{
    cudaStream_t st = ctx->stream;
    const int n_seq = sa.n_seq;
    std::vector<int32_t> tf((size_t)n_seq + 1);
    int64_t nt = 0;
    for (int i = 0; i < n_seq; ++i) { tf[i] = (int32_t)nt; nt += ((int64_t)lens[i] + SKT_TILE - 1) / SKT_TILE; }
    tf[n_seq] = (int32_t)nt;
    if (nt > INT32_MAX / 2) return TELR_ECAP;
    const int n_tiles = (int)nt;
    ENS(ctx->b_tfirst, (size_t)(n_seq + 1) * 4); ENS(ctx->b_tcnt, (size_t)(n_tiles + 1) * 4); ENS(ctx->b_toff, (size_t)(n_tiles + 2) * 8);
    ENS(ctx->b_tmpx, ((size_t)n_tiles * SKT_CAP + 1) * 8); ENS(ctx->b_tmpy, ((size_t)n_tiles * SKT_CAP + 1) * 4);
    CK(cudaMemcpyAsync(ctx->b_tfirst.p, tf.data(), (size_t)(n_seq + 1) * 4, cudaMemcpyHostToDevice, st));
    SketchTileArgs ta; memset(&ta, 0, sizeof(ta));
    ta.seq2 = sa.seq2; ta.nmask = sa.nmask; ta.bytes = sa.bytes; ta.seqs = sa.seqs; ta.n_seq = n_seq; ta.w = sa.w; ta.k = sa.k; ta.n_tiles = n_tiles;
    ta.tile_first = ctx->b_tfirst.as<int32_t>(); ta.tmp_x = ctx->b_tmpx.as<uint64_t>(); ta.tmp_y = ctx->b_tmpy.as<uint32_t>();
    ta.tile_cnt = ctx->b_tcnt.as<int32_t>(); ta.tile_off = ctx->b_toff.as<int64_t>(); ta.mz_off = ctx->b_mzoff.as<int64_t>();
    const int grid = std::max(1, std::min((n_tiles + SKT_WARPS - 1) / SKT_WARPS, ctx->sm_count * 8));
    std::vector<int64_t> hoff;
    if (sa.hpc && n_seq > 0) {      // map-pb: the step stream of every sequence (runs), at most its length
        hoff.resize((size_t)n_seq + 1);
        int64_t acc = 0;
        for (int i = 0; i < n_seq; ++i) { hoff[i] = acc; acc += lens[i]; }
        hoff[n_seq] = acc;
        ENS(ctx->b_hpc, acc + 64); ENS(ctx->b_hpp, (acc + 16) * 4); ENS(ctx->b_hpoff, (size_t)(n_seq + 1) * 8); ENS(ctx->b_hpn, (size_t)(n_seq + 1) * 4);
        CK(cudaMemcpyAsync(ctx->b_hpoff.p, hoff.data(), (size_t)(n_seq + 1) * 8, cudaMemcpyHostToDevice, st));
        ta.hp_code = ctx->b_hpc.as<uint8_t>(); ta.hp_pos = ctx->b_hpp.as<int32_t>(); ta.hp_off = ctx->b_hpoff.as<int64_t>(); ta.hp_n = ctx->b_hpn.as<int32_t>();
        { ++ctx->launches; k_hpc_compress<<<std::max(1, std::min((n_seq + 7) / 8, ctx->sm_count * 8)), 256, 0, st>>>(ta); }
    }
    if (n_tiles > 0) {
        const bool k32 = 2 * sa.k <= 30;
        if (sa.hpc) {
            if (sa.w == 10) { ++ctx->launches; k_sketch_tiles<false, 10, true><<<grid, SKT_WARPS * 32, 0, st>>>(ta); }         // map-pb
            else { ++ctx->launches; k_sketch_tiles<false, 0, true><<<grid, SKT_WARPS * 32, 0, st>>>(ta); }
        }
        else if (k32 && sa.w == 10) { ++ctx->launches; k_sketch_tiles<true, 10, false><<<grid, SKT_WARPS * 32, 0, st>>>(ta); }          // map-ont
        else if (!k32 && sa.w == 19) { ++ctx->launches; k_sketch_tiles<false, 19, false><<<grid, SKT_WARPS * 32, 0, st>>>(ta); }   // map-hifi
        else if (k32) { ++ctx->launches; k_sketch_tiles<true, 0, false><<<grid, SKT_WARPS * 32, 0, st>>>(ta); }
        else { ++ctx->launches; k_sketch_tiles<false, 0, false><<<grid, SKT_WARPS * 32, 0, st>>>(ta); }
    }
    { ++ctx->launches; k_excl_scan<int32_t><<<1, 1024, 0, st>>>(ctx->b_tcnt.as<int32_t>(), ctx->b_toff.as<int64_t>(), n_tiles, nullptr); }
    int64_t n_mz = 0;
    CK(cudaMemcpyAsync(&n_mz, ctx->b_toff.as<int64_t>() + n_tiles, 8, cudaMemcpyDeviceToHost, st));
    host_overlap();
    CK(cudaStreamSynchronize(st));     // tf[] must outlive its upload; n_mz sizes the CSR arrays
    ENS(ctx->b_mzx, (n_mz + 1) * 8); ENS(ctx->b_mzy, (n_mz + 1) * 4);
    ta.mz_x = ctx->b_mzx.as<uint64_t>(); ta.mz_y = ctx->b_mzy.as<uint32_t>();
    { ++ctx->launches; k_sketch_compact<<<std::max(1, std::min((n_tiles + 7) / 8, ctx->sm_count * 8)), 256, 0, st>>>(ta); }
    *n_mz_out = n_mz;
    return TELR_OK;
}

// Cuts a batch into chunks of loci of about equal read bases: n = ceil(total / budget) chunks, chunk i ends at the first locus
// that brings the running total to (i + 1) / n of the batch.  A locus is never split, so a single locus above the budget forms
// its own chunk; trailing loci without reads join the last chunk.  cuts = [0, ..., n_loci].
static void plan_chunks(const int32_t *read_len, const int32_t *lrb, int n_loci, int64_t budget, std::vector<int32_t> &cuts)
{
    cuts.clear(); cuts.push_back(0);
    if (n_loci <= 0) return;
    if (budget < 1) budget = 1;
    int64_t total = 0;
    for (int r = lrb[0]; r < lrb[n_loci]; ++r) total += read_len[r];
    const int64_t n_chunks = std::max<int64_t>(1, (total + budget - 1) / budget);
    int l0 = 0; int64_t cum = 0;
    for (int64_t ci = 0; l0 < n_loci; ++ci) {
        const bool last = ci + 1 >= n_chunks;
        const int64_t target = last ? total : (total * (ci + 1) + n_chunks - 1) / n_chunks;
        int l1 = l0;
        while (l1 < n_loci && (l1 == l0 || cum < target)) {
            for (int r = lrb[l1]; r < lrb[l1 + 1]; ++r) cum += read_len[r];
            ++l1;
        }
        if (last && cum >= total) l1 = n_loci;
        cuts.push_back(l1);
        l0 = l1;
    }
}

// Launches k_depth_af: the depth row of a contig strand lives in shared memory when it fits (216 KB = 55 296 positions),
// otherwise in a per-CTA row of global memory, so that any contig length is served (the reference runs samtools on any contig).
static int launch_depth(telr_af_ctx *ctx, DepthArgs &da, int n_loci, int max_len)
{
    const int cap_ints = 54 * 1024;
    const int want = max_len + 8;
    da.smem_ints = want < cap_ints ? want : cap_ints;
    int grid = std::max(1, std::min(n_loci, ctx->sm_count * 4));
    da.grow = nullptr; da.grow_stride = 0;
    if (want > cap_ints) {
        grid = std::max(1, std::min(n_loci, ctx->sm_count));
        da.grow_stride = ((int64_t)want + 63) & ~63LL;
        ENS(ctx->b_grow, (size_t)da.grow_stride * grid * 4);
        da.grow = ctx->b_grow.as<int32_t>();
    }
    const size_t dp_smem = (size_t)da.smem_ints * 4;
    CK(cudaFuncSetAttribute(k_depth_af, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp_smem));
    { ++ctx->launches; k_depth_af<<<grid, DP_THREADS, dp_smem, ctx->stream>>>(da); }
    return TELR_OK;
}

struct HostMeta {       // host copies of the small per-read / per-locus arrays
    std::vector<int32_t> read_len, lrb, contig_len;
};

// counters layout in b_ctr (int64 slots)
enum { C_WORK_CHAIN = 0, C_WORK_ALIGN = 1, C_ERR = 2, C_NWORK = 3, C_ANCH = 4, C_CELLS = 5, C_TASKS = 6, C_NBLK = 7, C_NALN = 8, C_NCIG = 9, C_MAXNA = 10, C_WORK_DP = 11, C_WORK_RMQ = 12, C_SLOTS = 16 };

static int run_chunk(telr_af_ctx *ctx, const Opt &o, const telr_af_batch *db /* device pointers */, const HostMeta &hm,
                     int l0, int l1, telr_af_result *dres /* device pointers for cov2x/af/depth */, const int64_t *h_depth_off,
                     telr_af_result *stats, int32_t *d_aln_out, int64_t aln_cap, uint32_t *d_cig_out, int64_t cig_cap)
{
    cudaStream_t st = ctx->stream;
    const int n_loci = l1 - l0;
    const int r0 = hm.lrb[l0], r1 = hm.lrb[l1], n_reads = r1 - r0;
    const int n_seq = n_reads + 2 * n_loci, n_prob = 2 * n_reads;
    const int sm = ctx->sm_count;
    if (n_loci <= 0) return TELR_OK;
    // ---- chunk-local index arrays ----
    std::vector<int32_t> lrb(n_loci + 1);
    std::vector<int64_t> cboff(n_loci + 1);
    int64_t ctg_total = 0; int max_tlen = 0, max_qlen = 0;
    for (int l = 0; l < n_loci; ++l) {
        lrb[l] = hm.lrb[l0 + l] - r0;
        cboff[l] = ctg_total;
        int L = hm.contig_len[l0 + l];
        if (L < 0) return TELR_EINVAL;
        ctg_total += 2 * (int64_t)L;
        max_tlen = std::max(max_tlen, L);
    }
    lrb[n_loci] = n_reads; cboff[n_loci] = ctg_total;
    for (int r = r0; r < r1; ++r) max_qlen = std::max(max_qlen, hm.read_len[r]);
    ENS(ctx->b_lrb, (n_loci + 1) * 4); ENS(ctx->b_cboff, (n_loci + 1) * 8); ENS(ctx->b_ctg, ctg_total + 64);
    ENS(ctx->b_ctr, C_SLOTS * 8);
    CK(cudaMemcpyAsync(ctx->b_lrb.p, lrb.data(), (n_loci + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_cboff.p, cboff.data(), (n_loci + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->b_ctr.p, 0, C_SLOTS * 8, st));
    int64_t *ctr = ctx->b_ctr.as<int64_t>();
    const int64_t *read_off = db->read_off + r0; const int32_t *read_len = db->read_len + r0; const uint32_t *read_hash = db->read_hash + r0;
    const int64_t *contig_off = db->contig_off + l0; const int32_t *contig_len = db->contig_len + l0;

    // ---- (e)+(f) depth, medians, AF (also the whole pipeline of a chunk without reads: zero coverage everywhere) ----
    auto run_depth = [&]() -> int {
        DepthArgs da; memset(&da, 0, sizeof(da));
        da.n_loci = n_loci; da.mode = ctx->depth_mode; da.contig_len = contig_len; da.te_start = db->te_start + l0; da.te_end = db->te_end + l0;
        da.locus_read_begin = ctx->b_lrb.as<int32_t>(); da.prob_blk_off = ctx->b_pblkoff.as<int64_t>(); da.prob_blk_cnt = ctx->b_pblkcnt.as<int32_t>();
        da.blocks = ctx->b_blk.as<int2>(); da.flank_len = db->flank_len; da.flank_off = db->flank_off; da.te_len = db->te_len; da.te_off = db->te_off;
        da.cov2x = dres->cov2x + (int64_t)l0 * 8; da.af = dres->af + l0; da.max_len = max_tlen;
        if (dres->depth) {
            std::vector<int64_t> doff(n_loci + 1);
            for (int l = 0; l < n_loci; ++l) doff[l] = h_depth_off[l0 + l];
            ENS(ctx->b_doff, (n_loci + 1) * 8);
            CK(cudaMemcpyAsync(ctx->b_doff.p, doff.data(), (size_t)n_loci * 8, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
            da.depth = dres->depth; da.depth_off = ctx->b_doff.as<int64_t>();
        }
        da.locus_bad = ctx->b_lbad.as<uint8_t>();
        return launch_depth(ctx, da, n_loci, max_tlen);
    };
    ENS(ctx->b_lbad, (size_t)n_loci + 64);
    CK(cudaMemsetAsync(ctx->b_lbad.p, 0, (size_t)n_loci + 64, st));
    if (n_reads == 0) {     // nothing to align: every window has depth 0 (the reference writes 0 coverages and freq None)
        ENS(ctx->b_pblkoff, 64); ENS(ctx->b_pblkcnt, 64); ENS(ctx->b_blk, 64);
        int rc = run_depth();
        if (rc != TELR_OK) return rc;
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        return TELR_OK;
    }

    CK(cudaEventRecord(ctx->ev[0], st));
    { ++ctx->launches; k_unpack_contigs<<<std::min(n_loci, sm * 8), 256, 0, st>>>(db->seq2, db->nmask, n_loci, contig_off, contig_len, ctx->b_cboff.as<int64_t>(), ctx->b_ctg.as<uint8_t>()); }
    ENS(ctx->b_descs, (size_t)n_seq * sizeof(SeqDesc)); ENS(ctx->b_counts, (size_t)(n_seq + 1) * 4); ENS(ctx->b_mzoff, (size_t)(n_seq + 2) * 8);
    { ++ctx->launches; k_build_descs<<<(n_seq + 255) / 256, 256, 0, st>>>(n_reads, n_loci, read_off, read_len, ctx->b_cboff.as<int64_t>(), contig_len, ctx->b_descs.as<SeqDesc>()); }
    // ---- (a) sketch ----
    SketchArgs sa; memset(&sa, 0, sizeof(sa));
    sa.seq2 = db->seq2; sa.nmask = db->nmask; sa.bytes = ctx->b_ctg.as<uint8_t>(); sa.seqs = ctx->b_descs.as<SeqDesc>();
    sa.n_seq = n_seq; sa.w = o.w; sa.k = o.k; sa.hpc = o.hpc;
    const int sk_grid = std::min(n_seq, sm * 8);
    if (o.hpc) {
        int64_t stride = ((int64_t)std::max(max_qlen, max_tlen) + 64) & ~63LL;
        ENS(ctx->b_hpc, stride * sk_grid); ENS(ctx->b_hpp, stride * sk_grid * 4); ENS(ctx->b_hpr, stride * sk_grid * 2);
        sa.hp_code = ctx->b_hpc.as<uint8_t>(); sa.hp_pos = ctx->b_hpp.as<int32_t>(); sa.hp_rl = ctx->b_hpr.as<uint16_t>(); sa.hp_stride = stride;
    }
    // LPT order of the chunk's problems (problem index 2*rb + strand*nr + r of k_chain -> read rb + r; longest read first,
    // ties by problem index): sorted on the host while the sketch kernels run
    std::vector<int32_t> order((size_t)std::max(n_prob, 1));
    auto make_order = [&]() {
        std::vector<int32_t> plen((size_t)std::max(n_prob, 1));
        for (int l = 0; l < n_loci; ++l) {
            const int rb = lrb[l], nr = lrb[l + 1] - rb;
            for (int sr = 0; sr < 2; ++sr)
                for (int r = 0; r < nr; ++r) plen[2 * rb + sr * nr + r] = hm.read_len[r0 + rb + r];
        }
        std::vector<int32_t> start((size_t)max_qlen + 2, 0);      // counting sort: descending length, ascending problem index
        for (int i = 0; i < n_prob; ++i) ++start[max_qlen - plen[i] + 1];
        for (int v = 0; v <= max_qlen; ++v) start[v + 1] += start[v];
        for (int i = 0; i < n_prob; ++i) order[start[max_qlen - plen[i]]++] = i;
    };
    sa.counts = ctx->b_counts.as<int32_t>();
    int64_t n_mz = 0;
    if (ctx->sketch_tiles) {
        std::vector<int32_t> lens((size_t)n_seq);
        for (int r = 0; r < n_reads; ++r) lens[r] = hm.read_len[r0 + r];
        for (int l = 0; l < n_loci; ++l) lens[n_reads + l] = lens[n_reads + n_loci + l] = hm.contig_len[l0 + l];
        int rc = sketch_tiled(ctx, sa, lens.data(), &n_mz, make_order);
        if (rc != TELR_OK) return rc;
        ENS(ctx->b_self, (n_mz + 1) * 2);
    } else {
        { ++ctx->launches; k_sketch<false><<<sk_grid, SK_THREADS, 0, st>>>(sa); }
        { ++ctx->launches; k_excl_scan<int32_t><<<1, 1024, 0, st>>>(ctx->b_counts.as<int32_t>(), ctx->b_mzoff.as<int64_t>(), n_seq, nullptr); }
        CK(cudaMemcpyAsync(&n_mz, ctx->b_mzoff.as<int64_t>() + n_seq, 8, cudaMemcpyDeviceToHost, st));
        make_order();
        CK(cudaStreamSynchronize(st));
        ENS(ctx->b_mzx, (n_mz + 1) * 8); ENS(ctx->b_mzy, (n_mz + 1) * 4); ENS(ctx->b_self, (n_mz + 1) * 2);
        sa.offs = ctx->b_mzoff.as<int64_t>(); sa.mz_x = ctx->b_mzx.as<uint64_t>(); sa.mz_y = ctx->b_mzy.as<uint32_t>();
        { ++ctx->launches; k_sketch<true><<<sk_grid, SK_THREADS, 0, st>>>(sa); }
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    ENS(ctx->b_order, (size_t)(n_prob + 1) * 4);
    CK(cudaMemcpyAsync(ctx->b_order.p, order.data(), (size_t)n_prob * 4, cudaMemcpyHostToDevice, st));
    {
        int64_t max_nmz = (int64_t)max_qlen + 16, tab = 64;
        while (tab < 2 * max_nmz) tab <<= 1;
        int grid = std::min(n_reads, sm * 4);
        if (grid < 1) grid = 1;
        ENS(ctx->b_tabk, tab * grid * 8); ENS(ctx->b_tabc, tab * grid * 4);
        { ++ctx->launches; k_self_count<<<grid, 256, 0, st>>>(n_reads, ctx->b_mzoff.as<int64_t>(), ctx->b_mzx.as<uint64_t>(), ctx->b_self.as<uint16_t>(),
                                           ctx->b_tabk.as<uint64_t>(), ctx->b_tabc.as<uint32_t>(), tab); }
    }
    // ---- (b)+(c) index, seeds, chains ----
    ENS(ctx->b_pna, (size_t)(n_prob + 1) * 4); ENS(ctx->b_pread, (size_t)(n_prob + 1) * 4); ENS(ctx->b_pls, (size_t)(n_prob + 1) * 4);
    ENS(ctx->b_paoff, (size_t)(n_prob + 2) * 8); ENS(ctx->b_prcap, (size_t)(n_prob + 1) * 4); ENS(ctx->b_proff, (size_t)(n_prob + 2) * 8);
    ENS(ctx->b_pnregs, (size_t)(n_prob + 1) * 4); ENS(ctx->b_pnca, (size_t)(n_prob + 1) * 4);
    ChainArgs ca; memset(&ca, 0, sizeof(ca));
    ca.o = o; ca.n_loci = n_loci; ca.n_reads = n_reads; ca.mode = 0;
    ca.locus_read_begin = ctx->b_lrb.as<int32_t>();
    ca.mz_off = ctx->b_mzoff.as<int64_t>(); ca.mz_x = ctx->b_mzx.as<uint64_t>(); ca.mz_y = ctx->b_mzy.as<uint32_t>(); ca.selfcnt = ctx->b_self.as<uint16_t>();
    ca.read_len = read_len; ca.read_hash = read_hash;
    ca.prob_na = ctx->b_pna.as<int32_t>(); ca.prob_read = ctx->b_pread.as<int32_t>(); ca.prob_ls = ctx->b_pls.as<int32_t>();
    ca.prob_nregs = ctx->b_pnregs.as<int32_t>(); ca.prob_nca = ctx->b_pnca.as<int32_t>();
    ca.work_counter = (int32_t *)(ctr + C_WORK_CHAIN); ca.err = (int32_t *)(ctr + C_ERR);
    ca.dp_counter = (int32_t *)(ctr + C_WORK_DP); ca.rmq_counter = (int32_t *)(ctr + C_WORK_RMQ); ca.stat_anchors = (unsigned long long *)(ctr + C_ANCH);
    const int ch_grid = std::min(2 * n_loci, sm);
    ENS(ctx->b_idxbig, sizeof(IdxBig) * (size_t)ch_grid);
    ca.idx_big = ctx->b_idxbig.as<IdxBig>();
    ca.locus_bad = ctx->b_lbad.as<uint8_t>();
    const size_t ch_smem = sizeof(IdxSmem);
    // per-warp seeding scratch: three int32 lists of up to one entry per read base (kept minimizers, occurrences, work list)
    const size_t ch_stride = (((size_t)max_qlen + 64) * 12 + 255) & ~(size_t)255;
    ENS(ctx->b_chws, ch_stride * (size_t)ch_grid * CH_WARPS);
    ENS(ctx->b_prep, (size_t)(n_prob + 1) * 4);
    ca.warp_scratch = ctx->b_chws.as<uint8_t>(); ca.warp_scratch_stride = ch_stride; ca.prob_replen = ctx->b_prep.as<int32_t>();
    CK(cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch_smem));
    { ++ctx->launches; k_chain<<<ch_grid, CH_THREADS, ch_smem, st>>>(ca); }
    { ++ctx->launches; k_excl_scan<int32_t><<<1, 1024, 0, st>>>(ctx->b_pna.as<int32_t>(), ctx->b_paoff.as<int64_t>(), n_prob, ctr + C_MAXNA); }
    ENS(ctx->b_psb, (size_t)(n_prob + 1) * 4); ENS(ctx->b_psoff, (size_t)(n_prob + 2) * 8); ENS(ctx->b_pnu, (size_t)(n_prob + 1) * 4); ENS(ctx->b_pm, (size_t)(n_prob + 1) * 4);
    { ++ctx->launches; k_reg_caps<<<(n_prob + 255) / 256, 256, 0, st>>>(n_prob, ctx->b_pna.as<int32_t>(), ctx->b_prcap.as<int32_t>(), ctx->b_psb.as<int32_t>()); }
    { ++ctx->launches; k_excl_scan<int32_t><<<1, 1024, 0, st>>>(ctx->b_prcap.as<int32_t>(), ctx->b_proff.as<int64_t>(), n_prob, nullptr); }
    { ++ctx->launches; k_excl_scan<int32_t><<<1, 1024, 0, st>>>(ctx->b_psb.as<int32_t>(), ctx->b_psoff.as<int64_t>(), n_prob, nullptr); }
    int64_t tot_na = 0, tot_rcap = 0, max_na = 0, tot_scr = 0;
    CK(cudaMemcpyAsync(&tot_na, ctx->b_paoff.as<int64_t>() + n_prob, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&tot_rcap, ctx->b_proff.as<int64_t>() + n_prob, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&tot_scr, ctx->b_psoff.as<int64_t>() + n_prob, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&max_na, ctr + C_MAXNA, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ctx->census) {
        std::vector<int32_t> na((size_t)n_prob);
        CK(cudaMemcpy(na.data(), ctx->b_pna.p, (size_t)n_prob * 4, cudaMemcpyDeviceToHost));
        std::sort(na.begin(), na.end());
        long long s2 = 0; for (int v : na) s2 += (long long)v * v;
        fprintf(stderr, "[census] anchors per problem: n %d total %lld max %lld p50 %d p90 %d p99 %d p99.9 %d sum-of-squares %.3g\n", n_prob, (long long)tot_na, (long long)max_na,
                na[n_prob / 2], na[(size_t)n_prob * 9 / 10], na[(size_t)n_prob * 99 / 100], na[(size_t)n_prob * 999 / 1000], (double)s2);
    }
    ENS(ctx->b_anch, (tot_na + 1) * sizeof(Anchor)); ENS(ctx->b_regs, (tot_rcap + 1) * sizeof(Reg)); ENS(ctx->b_pscr, tot_scr + 256);
    CK(cudaMemsetAsync(ctr + C_WORK_CHAIN, 0, 8, st));
    ca.mode = 1; ca.prob_aoff = ctx->b_paoff.as<int64_t>(); ca.prob_roff = ctx->b_proff.as<int64_t>();
    ca.anchors = ctx->b_anch.as<Anchor>(); ca.regs = ctx->b_regs.as<Reg>();
    ca.max_na = (int)max_na;
    ca.prob_scratch = ctx->b_pscr.as<uint8_t>(); ca.prob_soff = ctx->b_psoff.as<int64_t>();
    ca.prob_nu = ctx->b_pnu.as<int32_t>(); ca.prob_m = ctx->b_pm.as<int32_t>(); ca.n_prob = n_prob;
    { ++ctx->launches; k_chain<<<ch_grid, CH_THREADS, ch_smem, st>>>(ca); }
    {
        const int tpb = tp_blocks<TP_CHAIN>(n_prob, 128), wg = std::max(1, std::min((n_prob + 7) / 8, sm * 8));
        { ++ctx->launches; k_chain_sort<<<tpb, 128, 0, st>>>(ca); }
        { ++ctx->launches; k_chain_dp<<<wg, 256, 0, st>>>(ca); }
        { ++ctx->launches; k_chain_bt<<<tpb, 128, 0, st>>>(ca); }
        { ++ctx->launches; k_chain_rmq<<<wg, 256, 0, st>>>(ca); }
        { ++ctx->launches; k_chain_regs<<<tpb, 128, 0, st>>>(ca); }
    }
    CK(cudaEventRecord(ctx->ev[2], st));
    // ---- (d) alignment ----
    ENS(ctx->b_work, (size_t)(n_prob + 1) * 4);
    ENS(ctx->b_wflag, (size_t)(n_prob + 1) * 4); ENS(ctx->b_woff, (size_t)(n_prob + 2) * 8);
    { ++ctx->launches; k_work_flags<<<(n_prob + 255) / 256, 256, 0, st>>>(n_prob, ctx->b_order.as<int32_t>(), ctx->b_pnregs.as<int32_t>(), ctx->b_wflag.as<int32_t>()); }
    { ++ctx->launches; k_excl_scan<int32_t><<<1, 1024, 0, st>>>(ctx->b_wflag.as<int32_t>(), ctx->b_woff.as<int64_t>(), n_prob, nullptr); }
    { ++ctx->launches; k_work_scatter<<<(n_prob + 255) / 256, 256, 0, st>>>(n_prob, ctx->b_order.as<int32_t>(), ctx->b_wflag.as<int32_t>(), ctx->b_woff.as<int64_t>(),
                                                         ctx->b_work.as<int32_t>(), ctr + C_NWORK); }
    int64_t n_work64 = 0;
    CK(cudaMemcpyAsync(&n_work64, ctr + C_NWORK, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int n_work = (int)(n_work64 & 0xffffffff);
    int64_t chunk_bases = 0;
    for (int r = r0; r < r1; ++r) chunk_bases += hm.read_len[r];
    int64_t blocks_cap = chunk_bases / 2 + (int64_t)n_prob * 8 + 1024;
    ENS(ctx->b_blk, blocks_cap * 8); ENS(ctx->b_pblkoff, (size_t)(n_prob + 1) * 8); ENS(ctx->b_pblkcnt, (size_t)(n_prob + 1) * 4);
    CK(cudaMemsetAsync(ctx->b_pblkcnt.p, 0, (size_t)(n_prob + 1) * 4, st));
    CK(cudaMemsetAsync(ctx->b_pblkoff.p, 0, (size_t)(n_prob + 1) * 8, st));
    AlignArgs aa; memset(&aa, 0, sizeof(aa));
    ENS(ctx->b_opt, sizeof(Opt));
    CK(cudaMemcpyAsync(ctx->b_opt.p, &o, sizeof(Opt), cudaMemcpyHostToDevice, st));
    aa.o = o; aa.d_opt = ctx->b_opt.as<Opt>(); aa.n_prob = n_prob; aa.read_base = r0; aa.n_work = n_work; aa.read_len = read_len;
    aa.contig_len = contig_len; aa.ctg_boff = ctx->b_cboff.as<int64_t>(); aa.ctg_bytes = ctx->b_ctg.as<uint8_t>();
    aa.prob_read = ctx->b_pread.as<int32_t>(); aa.prob_ls = ctx->b_pls.as<int32_t>(); aa.prob_nca = ctx->b_pnca.as<int32_t>();
    aa.prob_nregs = ctx->b_pnregs.as<int32_t>(); aa.prob_aoff = ctx->b_paoff.as<int64_t>(); aa.prob_roff = ctx->b_proff.as<int64_t>();
    aa.anchors = ctx->b_anch.as<Anchor>(); aa.regs = ctx->b_regs.as<Reg>();
    aa.prob_scratch = ctx->b_pscr.as<uint8_t>(); aa.prob_soff = ctx->b_psoff.as<int64_t>(); aa.prob_replen = ctx->b_prep.as<int32_t>();
    aa.work_list = ctx->b_work.as<int32_t>();
    std::vector<int64_t> rbo(n_reads + 1);     // outlives its upload: the function synchronises again before it returns
    {   // nt4 bytes of every read of the chunk (forward + reverse complement)
        int64_t acc = 0;
        for (int r = 0; r < n_reads; ++r) { rbo[r] = acc; acc += 2 * (((int64_t)hm.read_len[r0 + r] + 15) & ~15LL); }
        rbo[n_reads] = acc;
        ENS(ctx->b_rboff, (size_t)(n_reads + 1) * 8); ENS(ctx->b_rbytes, acc + 64);
        CK(cudaMemcpyAsync(ctx->b_rboff.p, rbo.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
        { ++ctx->launches; k_unpack_reads<<<std::max(1, std::min(n_reads, sm * 16)), 128, 0, st>>>(n_reads, db->seq2, db->nmask, read_off, read_len, ctx->b_rboff.as<int64_t>(), ctx->b_rbytes.as<uint8_t>()); }
        aa.read_bytes = ctx->b_rbytes.as<uint8_t>(); aa.rbyte_off = ctx->b_rboff.as<int64_t>();
    }
    const int nwk = std::max(n_work, 1);
    ENS(ctx->b_alwork, (size_t)nwk * sizeof(AlWork)); ENS(ctx->b_alctx, (size_t)nwk * sizeof(AlnCtx)); ENS(ctx->b_altask, (size_t)nwk * sizeof(DpTask));
    ENS(ctx->b_alres, (size_t)nwk * sizeof(DpRes)); ENS(ctx->b_alsz, (size_t)(nwk + 1) * 4); ENS(ctx->b_aloff, (size_t)(nwk + 2) * 8);
    ENS(ctx->b_rc, 1024);
    const int al_grid = std::max(1, std::min((n_work + AL_WARPS - 1) / AL_WARPS, sm * ctx->al_blocks));
    {
        size_t maxT = ((size_t)max_tlen + 64) & ~(size_t)15;
        aa.max_tlen = max_tlen; aa.max_qlen = max_qlen; aa.dir_cap = ctx->dir_cap; aa.use_fast = ctx->use_fast; aa.use_vec = ctx->use_vec; aa.census = ctx->census;
        aa.warp_scratch_stride = (maxT * (6 + 4 + 24) + (((size_t)max_qlen + 64) & ~(size_t)15) * 6 + 512 + (size_t)ctx->dir_cap + 255) & ~(size_t)255;
        ENS(ctx->b_alws, aa.warp_scratch_stride * (size_t)al_grid * AL_WARPS);
        aa.warp_scratch = ctx->b_alws.as<uint8_t>();
        ENS(ctx->b_big, (size_t)ctx->n_big * ctx->big_cap); ENS(ctx->b_biglock, 256);
        CK(cudaMemsetAsync(ctx->b_biglock.p, 0, 256, st));
        aa.big = ctx->b_big.as<uint8_t>(); aa.big_cap = ctx->big_cap; aa.n_big = ctx->n_big; aa.big_lock = ctx->b_biglock.as<int32_t>();
    }
    aa.work = ctx->b_alwork.as<AlWork>(); aa.actx = ctx->b_alctx.as<AlnCtx>(); aa.tasks = ctx->b_altask.as<DpTask>(); aa.res = ctx->b_alres.as<DpRes>();
    aa.rc = ctx->b_rc.as<unsigned long long>();
    aa.err = (int32_t *)(ctr + C_ERR); aa.stat_cells = (unsigned long long *)(ctr + C_CELLS); aa.stat_tasks = (unsigned long long *)(ctr + C_TASKS);
    aa.blocks = ctx->b_blk.as<int2>(); aa.n_blocks = (unsigned long long *)(ctr + C_NBLK); aa.blocks_cap = blocks_cap;
    aa.prob_blk_off = ctx->b_pblkoff.as<int64_t>(); aa.prob_blk_cnt = ctx->b_pblkcnt.as<int32_t>();
    aa.aln_out = d_aln_out; aa.n_aln = (unsigned long long *)(ctr + C_NALN); aa.aln_cap = aln_cap;
    aa.cig_out = d_cig_out; aa.n_cig = (unsigned long long *)(ctr + C_NCIG); aa.cig_out_cap = cig_cap;
    if (d_aln_out) {    // continue numbering across chunks
        int64_t init[2] = {stats->n_aln, stats->n_cigar};
        CK(cudaMemcpyAsync(ctr + C_NALN, init, 16, cudaMemcpyHostToDevice, st));
    }
    CK(cudaEventRecord(ctx->ev[3], st));
    if (n_work > 0) {
        const int tb = (n_work + 127) / 128;
        { ++ctx->launches; k_al_sizes<<<tb, 128, 0, st>>>(aa, ctx->b_alsz.as<int32_t>()); }
        { ++ctx->launches; k_excl_scan<int32_t><<<1, 1024, 0, st>>>(ctx->b_alsz.as<int32_t>(), ctx->b_aloff.as<int64_t>(), n_work, nullptr); }
        int64_t cig_total = 0;
        CK(cudaMemcpyAsync(&cig_total, ctx->b_aloff.as<int64_t>() + n_work, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ENS(ctx->b_cigs, (size_t)(cig_total + 16) * 4);
        aa.cigs = ctx->b_cigs.as<uint32_t>();
        CK(cudaMemsetAsync(ctx->b_rc.p, 0, 1024, st));
        { ++ctx->launches; k_al_offsets<<<tb, 128, 0, st>>>(aa, ctx->b_aloff.as<int64_t>()); }
        { ++ctx->launches; k_al_init<<<tb, 128, 0, st>>>(aa); }
        if (ctx->al_queue < 0 ? db->preset == TELR_PRESET_MAP_ONT : ctx->al_queue > 0) {
            aa.q_cap = n_work + 1; aa.ext_per8 = ctx->ext_per8; aa.wide_per8 = ctx->wide_per8;
            ENS(ctx->b_qring, (size_t)AQ_ROLES * aa.q_cap * 4); ENS(ctx->b_qstate, (AQ_SLOTS + AQ_MAX_SM) * 4);
            CK(cudaMemsetAsync(ctx->b_qring.p, 0, (size_t)AQ_ROLES * aa.q_cap * 4, st));
            CK(cudaMemsetAsync(ctx->b_qstate.p, 0, (AQ_SLOTS + AQ_MAX_SM) * 4, st));
            aa.q_ring[0] = ctx->b_qring.as<int32_t>(); aa.q_ring[1] = aa.q_ring[0] + aa.q_cap; aa.q_ring[2] = aa.q_ring[1] + aa.q_cap; aa.q_state = ctx->b_qstate.as<int32_t>();
            CK(cudaFuncSetAttribute(k_al_queue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(AL_WARPS * sizeof(VecSmem))));
            { ++ctx->launches; k_al_queue<<<al_grid, AL_THREADS, AL_WARPS * sizeof(VecSmem), st>>>(aa); }
        } else {
            CK(cudaFuncSetAttribute(k_al_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(AL_WARPS * sizeof(VecSmem))));
            { ++ctx->launches; k_al_fused<<<al_grid, AL_THREADS, AL_WARPS * sizeof(VecSmem), st>>>(aa); }
        }
        { ++ctx->launches; k_al_regfin<<<std::max(1, std::min((n_work + 7) / 8, sm * 8)), 256, 0, st>>>(aa); }
        { ++ctx->launches; k_al_finish<<<tp_blocks<TP_FINISH>(n_work, 128), 128, 0, st>>>(aa); }
    }
    CK(cudaEventRecord(ctx->ev[4], st));
    { int rc = run_depth(); if (rc != TELR_OK) return rc; }
    CK(cudaEventRecord(ctx->ev[5], st));
    int64_t hc[C_SLOTS];
    CK(cudaMemcpyAsync(hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const int err = (int)(hc[C_ERR] & 0xffffffff);
    if (err) {
        fprintf(stderr, "[telr_af] device pipeline flagged error mask 0x%x (1 regcap 2 cigcap 4 kcap 8 dircap 16 idxcap 32 blkcap 64 alncap)\n", err);
        return (err & 16) ? TELR_EUNSUPPORTED : TELR_ECAP;
    }
    if (ctx->census && n_work > 0) {     // where the alignment kernel's warp cycles go (diagnostic, TELR_CENSUS=1)
        unsigned long long rc[128];
        CK(cudaMemcpy(rc, ctx->b_rc.p, sizeof(rc), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[census] cycles(M): coroutine %.1f | fill dp %.1f tb %.1f | vec dp %.1f tb %.1f | scalar dp %.1f tb %.1f | ll %.1f\n",
                rc[4] / 1e6, rc[5] / 1e6, rc[8] / 1e6, rc[6] / 1e6, rc[9] / 1e6, rc[7] / 1e6, rc[10] / 1e6, rc[11] / 1e6);
        fprintf(stderr, "[census] tasks: fill %llu vec %llu scalar %llu ll %llu | q*t (M): fill %.1f vec %.1f scalar %.1f\n",
                rc[12], rc[13], rc[14], rc[15], rc[16] / 1e6, rc[17] / 1e6, rc[18] / 1e6);
        static const char *fb[6] = {"<=256", "<=288", "<=320", "<=384", "<=512", ">512"};
        for (int b = 0; b < 6; ++b)
            fprintf(stderr, "[census] fill tlen %s: tasks %llu cycles(M) %.1f q*t(M) %.1f\n", fb[b], rc[64 + b], rc[70 + b] / 1e6, rc[76 + b] / 1e6);
        for (int b = 0; b < 6; ++b)
            fprintf(stderr, "[census] vec min(q,t) bucket %d: tasks %llu zdropped %llu cycles(M) %.1f mean max-diag %.0f mean q+t %.0f\n", b, rc[38 + b], rc[44 + b],
                    rc[32 + b] / 1e6, rc[38 + b] ? (double)rc[50 + b] / rc[38 + b] : 0.0, rc[38 + b] ? (double)rc[56 + b] / rc[38 + b] : 0.0);
    }
    stats->dp_cells += hc[C_CELLS]; stats->n_dp_tasks += hc[C_TASKS]; stats->n_anchors += hc[C_ANCH]; stats->n_minimizers += n_mz;
    stats->n_aln_blocks += hc[C_NBLK];
    if (d_aln_out) { stats->n_aln = hc[C_NALN]; stats->n_cigar = hc[C_NCIG]; }
    float ms;
    static const int pairs[5][3] = {{0, 1, 0}, {1, 2, 1}, {2, 3, 2}, {3, 4, 3}, {4, 5, 6}};
    for (auto &p : pairs) { cudaEventElapsedTime(&ms, ctx->ev[p[0]], ctx->ev[p[1]]); stats->ms_stage[p[2]] += ms; }
    return TELR_OK;
}

// Chunk loci by read bases.  Large chunks amortise the tails of the latency-bound chaining kernels and of the persistent
// alignment kernel (config 2 on one B200: 384 Mbase chunks 1924 loci/s, 768 Mbase 1973, one 1.8 Gbase chunk 2007), so the
// default takes what the device holds: about 80 B of workspace per read base (CIGAR arenas sized for the worst case ~45,
// sketch slots 12, alignment blocks 4, minimizers, anchors, chaining scratch, read bytes) on top of ~28 GB of per-warp
// alignment scratch, capped at 2 Gbase; equal-sized chunks.  On a 180 GB B200 that is 1.4 Gbase (measured peak 84 GB).
static int64_t chunk_budget(const telr_af_ctx *ctx)
{
    int64_t budget = ctx->chunk_bases;
    if (budget <= 0) {
        double usable = 0.8 * (double)ctx->total_mem;
        if (ctx->ws_limit > 0 && (double)ctx->ws_limit < usable) usable = (double)ctx->ws_limit;
        budget = (int64_t)((usable - 28.0 * (1 << 30)) / 80.0);
        budget = std::max<int64_t>((int64_t)128 << 20, std::min<int64_t>(budget, (int64_t)2048 << 20));
    }
    return budget;
}

// Host -> device upload of the packed bases behind the kernels: a host thread copies, chunk after chunk, the base range the
// chunk's loci cover (seq2 + nmask) on the ctx's copy stream and records one event per chunk; the compute stream waits for
// the event of the chunk it is about to process.  Pageable caller memory is fine (the driver stages it; that blocks only the
// uploader thread).  Ranges must advance monotonically with the chunks (they do for every batch packed locus by locus);
// otherwise the whole batch goes up with chunk 0.
struct Uploader {
    telr_af_ctx *ctx = nullptr;
    const telr_af_batch *hb = nullptr;
    uint32_t *d_seq2 = nullptr, *d_nmask = nullptr;
    std::vector<int64_t> lo, hi;                 // base range uploaded for chunk i
    std::vector<cudaEvent_t> ev;
    std::atomic<int> recorded{0};                // events recorded so far
    std::atomic<int> failed{0};
    std::thread th;

    int start(telr_af_ctx *c, const telr_af_batch *b, const HostMeta &hm, const std::vector<int32_t> &cuts, uint32_t *ds, uint32_t *dn)
    {
        ctx = c; hb = b; d_seq2 = ds; d_nmask = dn;
        const int nc = (int)cuts.size() - 1;
        lo.assign(nc, 0); hi.assign(nc, 0);
        bool mono = true; int64_t prev_hi = 0, prev_lo = 0;
        for (int ci = 0; ci < nc; ++ci) {
            int64_t a = INT64_MAX, z = 0;
            for (int l = cuts[ci]; l < cuts[ci + 1]; ++l) {
                if (hm.contig_len[l] > 0) { a = std::min<int64_t>(a, b->contig_off[l]); z = std::max<int64_t>(z, b->contig_off[l] + (((int64_t)hm.contig_len[l] + 63) & ~63LL)); }
            }
            for (int r = hm.lrb[cuts[ci]]; r < hm.lrb[cuts[ci + 1]]; ++r) { a = std::min<int64_t>(a, b->read_off[r]); z = std::max<int64_t>(z, b->read_off[r] + (((int64_t)hm.read_len[r] + 63) & ~63LL)); }
            if (a == INT64_MAX) a = z = prev_hi;
            a &= ~63LL; z = std::min<int64_t>((z + 63) & ~63LL, b->n_bases);
            if (a < prev_lo || z < prev_hi) mono = false;
            lo[ci] = std::max(a, prev_hi); hi[ci] = std::max(z, lo[ci]);
            prev_lo = a; prev_hi = std::max(prev_hi, z);
        }
        if (!mono && nc > 0) { for (int ci = 0; ci < nc; ++ci) lo[ci] = hi[ci] = b->n_bases; lo[0] = 0; }
        ev.resize(nc);
        for (auto &e : ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return TELR_ECUDA;
        th = std::thread([this]() { this->run(); });
        return TELR_OK;
    }
    void run()
    {
        if (cudaSetDevice(ctx->device) != cudaSuccess) { failed = 1; recorded = (int)ev.size(); return; }
        for (size_t ci = 0; ci < ev.size(); ++ci) {
            const int64_t a = lo[ci], z = hi[ci];
            if (z > a && !failed) {
                if (cudaMemcpyAsync(d_seq2 + a / 16, hb->seq2 + a / 16, (size_t)(z - a) / 4, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess ||
                    cudaMemcpyAsync(d_nmask + a / 32, hb->nmask + a / 32, (size_t)(z - a) / 8, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess) failed = 1;
            }
            if (cudaEventRecord(ev[ci], ctx->copy_stream) != cudaSuccess) failed = 1;
            recorded.store((int)ci + 1, std::memory_order_release);
        }
    }
    // make the compute stream wait for chunk ci's bases
    int wait(int ci)
    {
        while (recorded.load(std::memory_order_acquire) <= ci) std::this_thread::yield();
        if (failed) return TELR_ECUDA;
        return cudaStreamWaitEvent(ctx->stream, ev[ci], 0) == cudaSuccess ? TELR_OK : TELR_ECUDA;
    }
    void finish()
    {
        if (th.joinable()) th.join();
        for (auto &e : ev) cudaEventDestroy(e);
        ev.clear();
    }
    ~Uploader() { finish(); }
};

static int run_device(telr_af_ctx *ctx, const telr_af_batch *db, const HostMeta &hm, telr_af_result *dres, telr_af_result *stats,
                      int32_t *d_aln_out, int64_t aln_cap, uint32_t *d_cig_out, int64_t cig_cap,
                      const std::vector<int32_t> *cuts_in = nullptr, Uploader *up = nullptr)
{
    Opt o;
    if (db->preset < 0 || db->preset > 2) return TELR_EINVAL;
    opt_preset(o, db->preset);
    if (ctx->opt_bw > 0) o.bw = ctx->opt_bw;
    if (ctx->opt_bw_long > 0) o.bw_long = ctx->opt_bw_long;
    const int n_loci = db->n_loci;
    std::vector<int64_t> depth_off(n_loci + 1, 0);
    for (int l = 0; l < n_loci; ++l) depth_off[l + 1] = depth_off[l] + 2 * (int64_t)hm.contig_len[l];
    stats->dp_cells = stats->n_dp_tasks = stats->n_anchors = stats->n_minimizers = stats->n_aln_blocks = 0;
    stats->n_aln = stats->n_cigar = 0;
    for (int i = 0; i < 8; ++i) stats->ms_stage[i] = 0.f;
    std::vector<int32_t> cuts;
    if (cuts_in) cuts = *cuts_in;
    else plan_chunks(hm.read_len.data(), hm.lrb.data(), n_loci, chunk_budget(ctx), cuts);
    for (size_t ci = 0; ci + 1 < cuts.size(); ++ci) {
        const int l0 = cuts[ci], l1 = cuts[ci + 1];
        if (up) { int rc = up->wait((int)ci); if (rc != TELR_OK) return rc; }
        int rc = run_chunk(ctx, o, db, hm, l0, l1, dres, depth_off.data(), stats, d_aln_out, aln_cap, d_cig_out, cig_cap);
        if (rc != TELR_OK) return rc;
    }
    return TELR_OK;
}

static int validate_host_meta(const telr_af_batch *b, const HostMeta &hm)
{
    if (b->n_loci < 0 || b->n_reads < 0) return TELR_EINVAL;
    if (hm.lrb[0] != 0 || hm.lrb[b->n_loci] != b->n_reads) return TELR_EINVAL;
    for (int l = 0; l < b->n_loci; ++l) if (hm.lrb[l + 1] < hm.lrb[l]) return TELR_EINVAL;
    for (int r = 0; r < b->n_reads; ++r) if (hm.read_len[r] <= 0) return TELR_EINVAL;
    return TELR_OK;
}

extern "C" {

int telr_af_version(void) { return TELR_VERSION; }

const char *telr_af_strerror(int code)
{
    switch (code) {
    case TELR_OK: return "ok";
    case TELR_EINVAL: return "invalid argument or malformed batch";
    case TELR_ENOMEM: return "out of host or device memory";
    case TELR_ECUDA: return "CUDA runtime error";
    case TELR_ENODEV: return "no sm_100 CUDA device available";
    case TELR_ECAP: return "an internal capacity was exceeded";
    case TELR_EUNSUPPORTED: return "input outside the implemented scope";
    default: return "unknown error";
    }
}

int telr_af_last_cuda(const telr_af_ctx *ctx) { return ctx ? ctx->last_cuda : 0; }
long long telr_af_launch_count(const telr_af_ctx *ctx) { return ctx ? ctx->launches : 0; }
void *telr_af_stream(const telr_af_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int telr_af_set_option(telr_af_ctx *ctx, const char *name, int32_t value)
{
    if (!ctx || !name || value < 0) return TELR_EINVAL;
    if (!strcmp(name, "bw")) ctx->opt_bw = value;
    else if (!strcmp(name, "bw_long")) ctx->opt_bw_long = value;
    else return TELR_EINVAL;
    return TELR_OK;
}

int telr_af_create(telr_af_ctx **out, int device, size_t workspace_bytes)
{
    if (!out) return TELR_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) { cudaGetLastError(); return TELR_ENODEV; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TELR_ENODEV;
    if (prop.major != 10) { fprintf(stderr, "[telr_af] device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major, prop.minor); return TELR_ENODEV; }
    telr_af_ctx *ctx = new telr_af_ctx();
    ctx->device = device; ctx->sm_count = prop.multiProcessorCount; ctx->ws_limit = workspace_bytes; ctx->total_mem = prop.totalGlobalMem;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return TELR_ECUDA; }
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    const char *dm = getenv("TELR_DEPTH_MODE");
    if (dm) ctx->depth_mode = atoi(dm) ? 1 : 0;
    const char *cb = getenv("TELR_CHUNK_MBASES");
    if (cb) ctx->chunk_bases = (int64_t)atoll(cb) << 20;
    const char *uf = getenv("TELR_FAST_FILL");
    if (uf) ctx->use_fast = atoi(uf) ? 1 : 0;
    const char *uv = getenv("TELR_VEC_EXT");
    if (uv) ctx->use_vec = atoi(uv) ? 1 : 0;
    const char *alb = getenv("TELR_AL_CTAS");
    if (alb && atoi(alb) >= 1 && atoi(alb) <= AL_BLOCKS_PER_SM) ctx->al_blocks = atoi(alb);
    const char *skt = getenv("TELR_SKETCH_TILES");
    if (skt) ctx->sketch_tiles = atoi(skt) != 0;
    const char *cs = getenv("TELR_CENSUS");
    if (cs) ctx->census = atoi(cs) ? 1 : 0;
    const char *pm = getenv("TELR_POOL_MB");
    if (pm) ctx->pool_cap = (int64_t)atoll(pm) << 20;
    const char *dc = getenv("TELR_DIR_MB");
    if (dc) ctx->dir_cap = (int64_t)atoll(dc) << 20;
    const char *aq = getenv("TELR_AL_QUEUE");
    if (aq) ctx->al_queue = atoi(aq) ? 1 : 0;
    const char *e8 = getenv("TELR_AL_EXT8");
    if (e8 && atoi(e8) >= 0 && atoi(e8) <= 8) ctx->ext_per8 = atoi(e8);
    const char *w8 = getenv("TELR_AL_WIDE8");
    if (w8 && atoi(w8) >= 0 && atoi(w8) + ctx->ext_per8 <= 8) ctx->wide_per8 = atoi(w8);
    *out = ctx;
    return TELR_OK;
}

int telr_af_destroy(telr_af_ctx *ctx)
{
    if (!ctx) return TELR_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *all[] = {&ctx->b_cov, &ctx->b_af, &ctx->b_depth, &ctx->b_lrb, &ctx->b_cboff, &ctx->b_ctg, &ctx->b_descs, &ctx->b_counts, &ctx->b_mzoff,
                     &ctx->b_mzx, &ctx->b_mzy, &ctx->b_self, &ctx->b_tabk, &ctx->b_tabc, &ctx->b_hpc, &ctx->b_hpp, &ctx->b_hpr, &ctx->b_pna, &ctx->b_pread,
                     &ctx->b_pls, &ctx->b_paoff, &ctx->b_prcap, &ctx->b_proff, &ctx->b_pnregs, &ctx->b_pnca, &ctx->b_anch, &ctx->b_regs, &ctx->b_chws,
                     &ctx->b_alws, &ctx->b_work, &ctx->b_blk, &ctx->b_pblkoff, &ctx->b_pblkcnt, &ctx->b_ctr, &ctx->b_alnout, &ctx->b_cigout, &ctx->b_doff, &ctx->b_grow, &ctx->b_lbad,
                     &ctx->b_big, &ctx->b_biglock, &ctx->b_rbytes, &ctx->b_rboff, &ctx->b_alwork, &ctx->b_alctx, &ctx->b_altask, &ctx->b_alres, &ctx->b_alsz, &ctx->b_aloff,
                     &ctx->b_cigs, &ctx->b_pool, &ctx->b_tlist, &ctx->b_rc, &ctx->b_opt, &ctx->b_idxbig, &ctx->b_psb, &ctx->b_psoff, &ctx->b_pscr, &ctx->b_pnu, &ctx->b_pm,
                     &ctx->b_tfirst, &ctx->b_tcnt, &ctx->b_toff, &ctx->b_tmpx, &ctx->b_tmpy, &ctx->b_order, &ctx->b_wflag, &ctx->b_woff, &ctx->b_hpoff, &ctx->b_hpn, &ctx->b_qring, &ctx->b_qstate, &ctx->b_prep};
    for (auto *b : all) b->release();
    for (auto &b : ctx->b_in) b.release();
    for (auto &e : ctx->ev) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
    return TELR_OK;
}

int telr_af_run(telr_af_ctx *ctx, const telr_af_batch *hb, telr_af_result *hr)
{
    if (!ctx || !hb || !hr || !hr->cov2x || !hr->af) return TELR_EINVAL;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return TELR_ECUDA;
    cudaStream_t st = ctx->stream;
    HostMeta hm;
    hm.read_len.assign(hb->read_len, hb->read_len + hb->n_reads);
    hm.lrb.assign(hb->locus_read_begin, hb->locus_read_begin + hb->n_loci + 1);
    hm.contig_len.assign(hb->contig_len, hb->contig_len + hb->n_loci);
    int rc = validate_host_meta(hb, hm);
    if (rc) return rc;
    if (hb->n_bases % 64) return TELR_EINVAL;
    // ---- H2D: the small index arrays now, the packed bases chunk by chunk behind the kernels (Uploader) ----
    telr_af_batch db = *hb;
    const void *src[10] = {hb->seq2, hb->nmask, hb->read_off, hb->read_len, hb->read_hash, hb->locus_read_begin, hb->contig_off, hb->contig_len, hb->te_start, hb->te_end};
    const size_t bytes[10] = {(size_t)hb->n_bases / 4, (size_t)hb->n_bases / 8, (size_t)hb->n_reads * 8, (size_t)hb->n_reads * 4, (size_t)hb->n_reads * 4,
                              (size_t)(hb->n_loci + 1) * 4, (size_t)hb->n_loci * 8, (size_t)hb->n_loci * 4, (size_t)hb->n_loci * 4, (size_t)hb->n_loci * 4};
    for (int i = 0; i < 10; ++i) {
        ENS(ctx->b_in[i], bytes[i] + 64);
        if (i >= 2 && bytes[i]) CK(cudaMemcpyAsync(ctx->b_in[i].p, src[i], bytes[i], cudaMemcpyHostToDevice, st));
    }
    std::vector<int32_t> cuts;
    plan_chunks(hm.read_len.data(), hm.lrb.data(), hb->n_loci, chunk_budget(ctx), cuts);
    Uploader up;
    if (up.start(ctx, hb, hm, cuts, ctx->b_in[0].as<uint32_t>(), ctx->b_in[1].as<uint32_t>()) != TELR_OK) return TELR_ECUDA;
    db.seq2 = ctx->b_in[0].as<uint32_t>(); db.nmask = ctx->b_in[1].as<uint32_t>(); db.read_off = ctx->b_in[2].as<int64_t>();
    db.read_len = ctx->b_in[3].as<int32_t>(); db.read_hash = ctx->b_in[4].as<uint32_t>(); db.locus_read_begin = ctx->b_in[5].as<int32_t>();
    db.contig_off = ctx->b_in[6].as<int64_t>(); db.contig_len = ctx->b_in[7].as<int32_t>(); db.te_start = ctx->b_in[8].as<int32_t>();
    db.te_end = ctx->b_in[9].as<int32_t>();
    int64_t depth_n = 0;
    for (int l = 0; l < hb->n_loci; ++l) depth_n += 2 * (int64_t)hm.contig_len[l];
    telr_af_result dres; memset(&dres, 0, sizeof(dres));
    ENS(ctx->b_cov, (size_t)hb->n_loci * 32 + 64); ENS(ctx->b_af, (size_t)hb->n_loci * 8 + 64);
    dres.cov2x = ctx->b_cov.as<int32_t>(); dres.af = ctx->b_af.as<double>();
    if (hr->depth) { ENS(ctx->b_depth, (size_t)depth_n * 4 + 64); dres.depth = ctx->b_depth.as<int32_t>(); }
    int32_t *d_aln = nullptr; uint32_t *d_cig = nullptr;
    if (hr->aln && hr->cigar) {
        ENS(ctx->b_alnout, (size_t)hr->aln_cap * ALN_REC_INTS * 4 + 64); ENS(ctx->b_cigout, (size_t)hr->cigar_cap * 4 + 64);
        d_aln = ctx->b_alnout.as<int32_t>(); d_cig = ctx->b_cigout.as<uint32_t>();
    }
    rc = run_device(ctx, &db, hm, &dres, hr, d_aln, hr->aln_cap, d_cig, hr->cigar_cap, &cuts, &up);
    up.finish();
    if (rc) return rc;
    // ---- D2H ----
    CK(cudaMemcpyAsync(hr->cov2x, dres.cov2x, (size_t)hb->n_loci * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hr->af, dres.af, (size_t)hb->n_loci * 8, cudaMemcpyDeviceToHost, st));
    if (hr->depth) CK(cudaMemcpyAsync(hr->depth, dres.depth, (size_t)depth_n * 4, cudaMemcpyDeviceToHost, st));
    std::vector<int32_t> raw;
    if (d_aln) {
        raw.resize((size_t)hr->n_aln * ALN_REC_INTS + 16);
        if (hr->n_aln) CK(cudaMemcpyAsync(raw.data(), d_aln, (size_t)hr->n_aln * ALN_REC_INTS * 4, cudaMemcpyDeviceToHost, st));
        if (hr->n_cigar) CK(cudaMemcpyAsync(hr->cigar, d_cig, (size_t)hr->n_cigar * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    if (d_aln) {
        // records arrive in completion order; present them in (locus, strand, read, rank) order
        std::vector<int64_t> ord(hr->n_aln);
        for (int64_t i = 0; i < hr->n_aln; ++i) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) {
            const int32_t *x = &raw[a * ALN_REC_INTS], *y = &raw[b * ALN_REC_INTS];       // problem index orders by (locus, strand, read)
            if (x[14] != y[14]) return x[14] < y[14];
            return x[15] < y[15];
        });
        for (int64_t i = 0; i < hr->n_aln; ++i) {
            const int32_t *x = &raw[ord[i] * ALN_REC_INTS];
            telr_aln &a = hr->aln[i];
            a.read = x[0]; a.strand = x[1]; a.rs = x[2]; a.re = x[3]; a.qs = x[4]; a.qe = x[5]; a.rev = x[6]; a.flag = x[7];
            a.dp_max = x[8]; a.mlen = x[9]; a.blen = x[10]; a.n_cigar = x[11];
            a.cigar_off = (int64_t)(uint32_t)x[12] | (int64_t)x[13] << 32;
            a.mapq = x[16]; a.dp_score = x[17]; a.cnt = x[18]; a.score = x[19]; a.subsc = x[20]; a.n_ambi = x[21]; a.inv = x[22]; a.n_sub = x[23];
        }
    }
    return TELR_OK;
}

int telr_af_run_device(telr_af_ctx *ctx, const telr_af_batch *db, telr_af_result *dr)
{
    if (!ctx || !db || !dr || !dr->cov2x || !dr->af) return TELR_EINVAL;
    if (dr->aln || dr->cigar) return TELR_EINVAL;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return TELR_ECUDA;
    HostMeta hm;
    hm.read_len.resize(db->n_reads); hm.lrb.resize(db->n_loci + 1); hm.contig_len.resize(db->n_loci);
    CK(cudaMemcpyAsync(hm.read_len.data(), db->read_len, (size_t)db->n_reads * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hm.lrb.data(), db->locus_read_begin, (size_t)(db->n_loci + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hm.contig_len.data(), db->contig_len, (size_t)db->n_loci * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int rc = validate_host_meta(db, hm);
    if (rc) return rc;
    telr_af_result dres = *dr;
    rc = run_device(ctx, db, hm, &dres, dr, nullptr, 0, nullptr, 0);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return TELR_OK;
}

int telr_af_sketch(telr_af_ctx *ctx, const uint32_t *seq2, const uint32_t *nmask, int64_t n_bases, int32_t n_seq,
                   const int64_t *seq_off, const int32_t *seq_len, int32_t w, int32_t k, int32_t hpc,
                   uint64_t *mz_x, uint64_t *mz_y, int64_t mz_cap, int64_t *mz_off)
{
    if (!ctx || n_seq < 0 || (k & 1) == 0 || k > 28 || w <= 0 || w > SK_XH || w + k > SK_CH) return TELR_EINVAL;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return TELR_ECUDA;
    cudaStream_t st = ctx->stream;
    ENS(ctx->b_in[0], n_bases / 4 + 64); ENS(ctx->b_in[1], n_bases / 8 + 64);
    CK(cudaMemcpyAsync(ctx->b_in[0].p, seq2, n_bases / 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_in[1].p, nmask, n_bases / 8, cudaMemcpyHostToDevice, st));
    std::vector<SeqDesc> d(n_seq);
    int max_len = 0;
    for (int i = 0; i < n_seq; ++i) { d[i].off = seq_off[i]; d[i].len = seq_len[i]; d[i].kind = 0; max_len = std::max(max_len, seq_len[i]); }
    ENS(ctx->b_descs, (size_t)n_seq * sizeof(SeqDesc) + 64); ENS(ctx->b_counts, (size_t)(n_seq + 1) * 4); ENS(ctx->b_mzoff, (size_t)(n_seq + 2) * 8);
    CK(cudaMemcpyAsync(ctx->b_descs.p, d.data(), (size_t)n_seq * sizeof(SeqDesc), cudaMemcpyHostToDevice, st));
    SketchArgs sa; memset(&sa, 0, sizeof(sa));
    sa.seq2 = ctx->b_in[0].as<uint32_t>(); sa.nmask = ctx->b_in[1].as<uint32_t>(); sa.seqs = ctx->b_descs.as<SeqDesc>();
    sa.n_seq = n_seq; sa.w = w; sa.k = k; sa.hpc = hpc;
    const int grid = std::max(1, std::min(n_seq, ctx->sm_count * 8));
    if (hpc) {
        int64_t stride = ((int64_t)max_len + 64) & ~63LL;
        ENS(ctx->b_hpc, stride * grid); ENS(ctx->b_hpp, stride * grid * 4); ENS(ctx->b_hpr, stride * grid * 2);
        sa.hp_code = ctx->b_hpc.as<uint8_t>(); sa.hp_pos = ctx->b_hpp.as<int32_t>(); sa.hp_rl = ctx->b_hpr.as<uint16_t>(); sa.hp_stride = stride;
    }
    sa.counts = ctx->b_counts.as<int32_t>();
    int64_t n_mz = 0;
    if (ctx->sketch_tiles) {
        int rc = sketch_tiled(ctx, sa, seq_len, &n_mz, []() {});
        if (rc != TELR_OK) return rc;
        CK(cudaMemcpyAsync(mz_off, ctx->b_mzoff.p, (size_t)(n_seq + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (n_mz > mz_cap) return TELR_ECAP;
    } else {
        { ++ctx->launches; k_sketch<false><<<grid, SK_THREADS, 0, st>>>(sa); }
        { ++ctx->launches; k_excl_scan<int32_t><<<1, 1024, 0, st>>>(ctx->b_counts.as<int32_t>(), ctx->b_mzoff.as<int64_t>(), n_seq, nullptr); }
        CK(cudaMemcpyAsync(mz_off, ctx->b_mzoff.p, (size_t)(n_seq + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        n_mz = mz_off[n_seq];
        if (n_mz > mz_cap) return TELR_ECAP;
        ENS(ctx->b_mzx, (n_mz + 1) * 8); ENS(ctx->b_mzy, (n_mz + 1) * 4);
        sa.offs = ctx->b_mzoff.as<int64_t>(); sa.mz_x = ctx->b_mzx.as<uint64_t>(); sa.mz_y = ctx->b_mzy.as<uint32_t>();
        { ++ctx->launches; k_sketch<true><<<grid, SK_THREADS, 0, st>>>(sa); }
    }
    std::vector<uint32_t> y32(n_mz + 1);
    CK(cudaMemcpyAsync(mz_x, ctx->b_mzx.p, (size_t)n_mz * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(y32.data(), ctx->b_mzy.p, (size_t)n_mz * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (int64_t i = 0; i < n_mz; ++i) mz_y[i] = y32[i];
    return TELR_OK;
}

int telr_af_depth_af(telr_af_ctx *ctx, int32_t n_loci, const int32_t *contig_len, const int32_t *te_start, const int32_t *te_end,
                     int32_t flank_len, int32_t flank_off, int32_t te_len, int32_t te_off, int64_t n_blocks,
                     const int32_t *blk_ls, const int32_t *blk_start, const int32_t *blk_len, int32_t *depth, int32_t *cov2x, double *af)
{
    if (!ctx || n_loci <= 0 || !cov2x || !af) return TELR_EINVAL;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return TELR_ECUDA;
    cudaStream_t st = ctx->stream;
    // one pseudo-read per (locus, strand): bucket the blocks on the host
    std::vector<int32_t> lrb(n_loci + 1), cnt(2 * n_loci, 0);
    std::vector<int64_t> off(2 * n_loci + 1, 0), doff(n_loci + 1, 0);
    for (int l = 0; l <= n_loci; ++l) lrb[l] = l;
    int max_len = 0;
    for (int l = 0; l < n_loci; ++l) { doff[l + 1] = doff[l] + 2 * (int64_t)contig_len[l]; max_len = std::max(max_len, contig_len[l]); }
    for (int64_t i = 0; i < n_blocks; ++i) { if (blk_ls[i] < 0 || blk_ls[i] >= 2 * n_loci) return TELR_EINVAL; ++cnt[blk_ls[i]]; }
    for (int i = 0; i < 2 * n_loci; ++i) off[i + 1] = off[i] + cnt[i];
    std::vector<int2> blk(n_blocks + 1);
    std::vector<int64_t> fill(off.begin(), off.end() - 1);
    for (int64_t i = 0; i < n_blocks; ++i) blk[fill[blk_ls[i]]++] = make_int2(blk_start[i], blk_len[i]);
    // problem index of (locus l, strand s) with one read per locus is 2*l + s
    ENS(ctx->b_lrb, (n_loci + 1) * 4); ENS(ctx->b_pblkoff, (size_t)(2 * n_loci + 1) * 8); ENS(ctx->b_pblkcnt, (size_t)(2 * n_loci + 1) * 4);
    ENS(ctx->b_blk, (size_t)(n_blocks + 1) * 8); ENS(ctx->b_in[7], (size_t)n_loci * 4 + 64); ENS(ctx->b_in[8], (size_t)n_loci * 4 + 64); ENS(ctx->b_in[9], (size_t)n_loci * 4 + 64);
    ENS(ctx->b_cov, (size_t)n_loci * 32 + 64); ENS(ctx->b_af, (size_t)n_loci * 8 + 64); ENS(ctx->b_doff, (size_t)(n_loci + 1) * 8);
    ENS(ctx->b_depth, (size_t)doff[n_loci] * 4 + 64);
    CK(cudaMemcpyAsync(ctx->b_lrb.p, lrb.data(), (size_t)(n_loci + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_pblkoff.p, off.data(), (size_t)(2 * n_loci) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_pblkcnt.p, cnt.data(), (size_t)(2 * n_loci) * 4, cudaMemcpyHostToDevice, st));
    if (n_blocks) CK(cudaMemcpyAsync(ctx->b_blk.p, blk.data(), (size_t)n_blocks * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_in[7].p, contig_len, (size_t)n_loci * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_in[8].p, te_start, (size_t)n_loci * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_in[9].p, te_end, (size_t)n_loci * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_doff.p, doff.data(), (size_t)(n_loci + 1) * 8, cudaMemcpyHostToDevice, st));
    DepthArgs da; memset(&da, 0, sizeof(da));
    da.n_loci = n_loci; da.mode = ctx->depth_mode; da.contig_len = ctx->b_in[7].as<int32_t>(); da.te_start = ctx->b_in[8].as<int32_t>();
    da.te_end = ctx->b_in[9].as<int32_t>(); da.locus_read_begin = ctx->b_lrb.as<int32_t>();
    da.prob_blk_off = ctx->b_pblkoff.as<int64_t>(); da.prob_blk_cnt = ctx->b_pblkcnt.as<int32_t>(); da.blocks = ctx->b_blk.as<int2>();
    da.flank_len = flank_len; da.flank_off = flank_off; da.te_len = te_len; da.te_off = te_off;
    da.depth = ctx->b_depth.as<int32_t>(); da.depth_off = ctx->b_doff.as<int64_t>();
    da.cov2x = ctx->b_cov.as<int32_t>(); da.af = ctx->b_af.as<double>(); da.max_len = max_len;
    { int rc = launch_depth(ctx, da, n_loci, max_len); if (rc != TELR_OK) return rc; }
    CK(cudaMemcpyAsync(cov2x, ctx->b_cov.p, (size_t)n_loci * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(af, ctx->b_af.p, (size_t)n_loci * 8, cudaMemcpyDeviceToHost, st));
    if (depth) CK(cudaMemcpyAsync(depth, ctx->b_depth.p, (size_t)doff[n_loci] * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return TELR_OK;
}

}  // extern "C"

// ---- DP stage entry point ---------------------------------------------------------------------
struct DpStageArgs {
    Opt o; int32_t n_tasks; const telr_dp_task *tasks; const uint8_t *q, *t; telr_dp_out *out;
    uint32_t *cig; unsigned long long *n_cig; int64_t cig_cap;
    uint8_t *warp_scratch; size_t stride; int32_t maxQ, maxT, use_fast, use_vec; int64_t dir_cap;
    int32_t *work_counter, *err; unsigned long long *cells;
};

__global__ void __launch_bounds__(AL_THREADS, AL_BLOCKS_PER_SM) k_dp_stage(const __grid_constant__ DpStageArgs A)
{
    __shared__ DpRes RS[AL_WARPS];
    __shared__ DpTask TS[AL_WARPS];
    __shared__ unsigned long long CS[AL_WARPS];
    __shared__ uint2 stab[256];
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    VecSmem *DS = reinterpret_cast<VecSmem *>(dyn_smem);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t *base = A.warp_scratch + (size_t)(blockIdx.x * AL_WARPS + wid) * A.stride;
    const size_t maxT = (size_t)A.maxT, maxQ = (size_t)A.maxQ;
    DpScratch S;
    S.u = (int8_t *)base; base += maxT; S.v = (int8_t *)base; base += maxT; S.x = (int8_t *)base; base += maxT;
    S.y = (int8_t *)base; base += maxT; S.x2 = (int8_t *)base; base += maxT; S.y2 = (int8_t *)base; base += maxT;
    S.H = (int32_t *)base; base += maxT * 4;
    S.ll = (int32_t *)base; base += maxT * 24;
    S.ezcap = (int32_t)(maxQ + maxT); S.ezcig = (uint32_t *)base; base += (size_t)S.ezcap * 4;
    S.bnd = (uint32_t *)base; base += maxQ * 6;
    base = (uint8_t *)(((uintptr_t)base + 255) & ~(uintptr_t)255);
    S.dir = base; S.dir_cap = A.dir_cap;
    S.s_state = nullptr; S.s_H = nullptr; S.vsm = A.use_vec ? &DS[wid] : nullptr; S.stab = stab;
    telr::vec_fill_stab(stab, A.o);
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(A.work_counter, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= A.n_tasks) break;
        if (lane == 0) {
            const telr_dp_task &t = A.tasks[i];
            DpTask &T = TS[wid];
            T.kind = 0; T.q = A.q + t.q_off; T.t = A.t + t.t_off; T.qstep = T.tstep = 1; T.qcomp = 0;
            T.qlen = t.qlen; T.tlen = t.tlen; T.w = t.w; T.zdrop = t.zdrop; T.end_bonus = t.end_bonus; T.flag = t.flag;
            CS[wid] = 0;
        }
        __syncwarp();
        bool done_fast = false;
        if (A.use_fast && fill_fast_ok(TS[wid]) && (int64_t)TS[wid].qlen * fill_stride(TS[wid].tlen) <= S.dir_cap)
            done_fast = warp_fill_fast(A.o, TS[wid], RS[wid], S.dir, S.bnd, &CS[wid]);
        int vec = 0;
        if (!done_fast) vec = warp_extd2(A.o, TS[wid], RS[wid], S, &CS[wid], A.err);
        __syncwarp();
        if (done_fast) fill_traceback(TS[wid], RS[wid], S.dir, S.ezcig, S.ezcap, A.err, *reinterpret_cast<TbSmem *>(DS[wid].H));     // the DP state window is idle during traceback
        else if (lane == 0) { if (vec) extd2_traceback_vec(TS[wid], RS[wid], S.dir, S.ezcig, S.ezcap, A.err); else extd2_traceback(TS[wid], RS[wid], S.dir, S.ezcig, S.ezcap, A.err); }
        __syncwarp();
        if (lane == 0) {
            const DpRes &R = RS[wid];
            telr_dp_out &o = A.out[i];
            o.max = R.max; o.max_q = R.max_q; o.max_t = R.max_t; o.mqe = R.mqe; o.mqe_t = R.mqe_t; o.mte = R.mte; o.mte_q = R.mte_q;
            o.score = R.score; o.zdropped = R.zdropped; o.reach_end = R.reach_end; o.n_cigar = R.n_cigar; o.cells = (int64_t)CS[wid];
            long long co = (long long)atomicAdd(A.n_cig, (unsigned long long)R.n_cigar);
            o.cigar_off = co;
            if (co + R.n_cigar <= A.cig_cap) for (int k = 0; k < R.n_cigar; ++k) A.cig[co + k] = R.cigar[k];
            else atomicOr(A.err, 64);
            atomicAdd(A.cells, CS[wid]);
        }
        __syncwarp();
    }
}

extern "C" int telr_af_dp(telr_af_ctx *ctx, int32_t preset, int32_t n_tasks, const telr_dp_task *tasks, const uint8_t *qseq, int64_t qbytes,
                          const uint8_t *tseq, int64_t tbytes, telr_dp_out *out, uint32_t *cigar, int64_t cigar_cap)
{
    if (!ctx || n_tasks < 0 || preset < 0 || preset > 2) return TELR_EINVAL;
    if (n_tasks == 0) return TELR_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return TELR_ECUDA;
    cudaStream_t st = ctx->stream;
    DpStageArgs A; memset(&A, 0, sizeof(A));
    opt_preset(A.o, preset);
    int maxQ = 0, maxT = 0;
    for (int i = 0; i < n_tasks; ++i) { maxQ = std::max(maxQ, tasks[i].qlen); maxT = std::max(maxT, tasks[i].tlen); }
    A.maxQ = (maxQ + 64) & ~15; A.maxT = (maxT + 64) & ~15; A.dir_cap = ctx->dir_cap;
    A.use_fast = ctx->use_fast ? (ctx->al_queue > 0 && ctx->wide_per8 > 0 ? 2 : 1) : 0; A.use_vec = ctx->use_vec;
    A.stride = ((size_t)A.maxT * (6 + 4 + 24) + (size_t)(A.maxQ + A.maxT) * 4 + (size_t)A.maxQ * 6 + 512 + (size_t)A.dir_cap + 255) & ~(size_t)255;
    const int grid = std::max(1, std::min((n_tasks + AL_WARPS - 1) / AL_WARPS, ctx->sm_count * AL_BLOCKS_PER_SM));
    ENS(ctx->b_alws, A.stride * (size_t)grid * AL_WARPS);
    ENS(ctx->b_in[0], qbytes + 64); ENS(ctx->b_in[1], tbytes + 64); ENS(ctx->b_in[2], (size_t)n_tasks * sizeof(telr_dp_task));
    ENS(ctx->b_in[3], (size_t)n_tasks * sizeof(telr_dp_out)); ENS(ctx->b_cigout, (size_t)cigar_cap * 4 + 64); ENS(ctx->b_ctr, C_SLOTS * 8);
    CK(cudaMemcpyAsync(ctx->b_in[0].p, qseq, qbytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_in[1].p, tseq, tbytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->b_in[2].p, tasks, (size_t)n_tasks * sizeof(telr_dp_task), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->b_ctr.p, 0, C_SLOTS * 8, st));
    int64_t *ctr = ctx->b_ctr.as<int64_t>();
    A.n_tasks = n_tasks; A.tasks = ctx->b_in[2].as<telr_dp_task>(); A.q = ctx->b_in[0].as<uint8_t>(); A.t = ctx->b_in[1].as<uint8_t>();
    A.out = ctx->b_in[3].as<telr_dp_out>(); A.cig = ctx->b_cigout.as<uint32_t>(); A.n_cig = (unsigned long long *)(ctr + C_NCIG); A.cig_cap = cigar_cap;
    A.warp_scratch = ctx->b_alws.as<uint8_t>(); A.work_counter = (int32_t *)(ctr + C_WORK_ALIGN); A.err = (int32_t *)(ctr + C_ERR);
    A.cells = (unsigned long long *)(ctr + C_CELLS);
    CK(cudaFuncSetAttribute(k_dp_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(AL_WARPS * sizeof(VecSmem))));
    CK(cudaEventRecord(ctx->ev[0], st));
    { ++ctx->launches; k_dp_stage<<<grid, AL_THREADS, AL_WARPS * sizeof(VecSmem), st>>>(A); }
    CK(cudaEventRecord(ctx->ev[1], st));
    int64_t hc[C_SLOTS];
    CK(cudaMemcpyAsync(hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out, ctx->b_in[3].p, (size_t)n_tasks * sizeof(telr_dp_out), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (ctx->census) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); fprintf(stderr, "[census] k_dp_stage %.3f ms, %lld cells\n", ms, (long long)hc[C_CELLS]); }
    if (hc[C_ERR] & 0xffffffff) return TELR_ECAP;
    if (hc[C_NCIG]) CK(cudaMemcpy(cigar, ctx->b_cigout.p, (size_t)hc[C_NCIG] * 4, cudaMemcpyDeviceToHost));
    return TELR_OK;
}

// ---- host helpers ------------------------------------------------------------------------------
extern "C" int telr_af_plan_chunks(const int32_t *read_len, const int32_t *locus_read_begin, int32_t n_loci, int64_t budget_bases, int32_t *cuts, int32_t cap)
{
    if (!read_len || !locus_read_begin || n_loci < 0 || !cuts) return TELR_EINVAL;
    std::vector<int32_t> c;
    plan_chunks(read_len, locus_read_begin, n_loci, budget_bases, c);
    if ((int)c.size() > cap) return TELR_ECAP;
    for (size_t i = 0; i < c.size(); ++i) cuts[i] = c[i];
    return (int)c.size() - 1;
}

extern "C" int telr_pack_seq(const char *s, int32_t len, int64_t off, uint32_t *seq2, uint32_t *nmask)
{
    if ((off & 63) || len < 0) return TELR_EINVAL;
    int64_t nw = ((int64_t)len + 63) / 64 * 4;
    memset(seq2 + off / 16, 0, (size_t)nw * 4);
    memset(nmask + off / 32, 0, (size_t)(nw / 2) * 4);
    static const int8_t lut[256] = {
#define N4 4, 4, 4, 4
#define N16 N4, N4, N4, N4
        N16, N16, N16, N16,
        4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
        4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
        N16, N16, N16, N16, N16, N16, N16, N16};
    for (int32_t i = 0; i < len; ++i) {
        int c = lut[(uint8_t)s[i]];
        int64_t p = off + i;
        if (c < 4) seq2[p >> 4] |= (uint32_t)c << (2 * (p & 15));
        else nmask[p >> 5] |= 1u << (p & 31);
    }
    return TELR_OK;
}

extern "C" uint32_t telr_name_hash(const char *s)
{
    uint32_t h = (uint32_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    return h;
}
