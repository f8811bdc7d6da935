// k_sketch.cuh — kernel (a): batched (w,k)-minimizer sketching of reads and contig strands.
//
// Replaces mm_sketch() (minimap2 sketch.c) reached from the reference at TELR_te.py:505.  Every step evaluates, from the
// w+1 hashes around it, exactly the pushes the sequential ring-buffer algorithm performs at that step (first-window
// special case, "new minimum", "minimum left the window" with its identical-hash flush: sk_step_g), so the emitted list
// is byte-identical to mm_sketch, including around ambiguous bases.
//
// Two implementations live here:
//   * k_sketch_tiles (+ k_hpc_compress for map-pb, + k_sketch_compact): the product path — one warp per tile of 2048
//     steps, single pass, see the block comment above it;
//   * k_sketch<COUNT/WRITE>: the first-generation kernel (one CTA per sequence, tiles of 2048 steps behind CTA barriers,
//     a count pass and a write pass into exact CSR offsets, homopolymer compression as a per-sequence pre-pass), kept as
//     a second implementation behind TELR_SKETCH_TILES=0.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include "mm_types.cuh"

namespace telr {

constexpr int SK_THREADS = 256;
constexpr int SK_TILE = 2048;
constexpr int SK_CH = 64;     // code halo
constexpr int SK_XH = 32;     // hash halo (>= w)

struct SeqDesc {
    int64_t off;     // base offset into the packed arrays (kind 0) or byte offset into bytes (kind 1)
    int32_t len;
    int32_t kind;
};

struct SketchArgs {
    const uint32_t *seq2, *nmask;
    const uint8_t *bytes;
    const SeqDesc *seqs;
    int32_t n_seq, w, k, hpc;
    // HPC scratch, one slice per CTA
    uint8_t *hp_code; int32_t *hp_pos; uint16_t *hp_rl; int64_t hp_stride;
    // COUNT pass output / WRITE pass input
    int32_t *counts;          // [n_seq]
    const int64_t *offs;      // [n_seq+1]
    uint64_t *mz_x; uint32_t *mz_y;
};

// exclusive scan over the NT threads of a CTA (NT a multiple of 32, <= 1024); smem_ws >= 33 ints
template <int NT> __device__ __forceinline__ int block_excl_scan_t(int v, int *total, int *smem_ws)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) smem_ws[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int s = lane < (NT / 32) ? smem_ws[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += y;
        }
        if (lane < NT / 32) smem_ws[lane] = s;
    }
    __syncthreads();
    int base = wid ? smem_ws[wid - 1] : 0;
    *total = smem_ws[NT / 32 - 1];
    __syncthreads();
    return base + x - v;
}
__device__ __forceinline__ int block_excl_scan(int v, int *total, int *smem_ws) { return block_excl_scan_t<SK_THREADS>(v, total, smem_ws); }

struct SkSmem {
    uint64_t X[SK_XH + SK_TILE];
    int32_t pos[SK_XH + SK_TILE];
    uint16_t rl[SK_CH + SK_TILE + 8];
    uint8_t code[SK_CH + SK_TILE + 8];
    uint8_t z[SK_XH + SK_TILE];
    uint8_t lcap[SK_TILE];
    int ws[40];
};

// pushes of the sequential algorithm at one step.  Xd(d) = hash of the step d steps back (d = 0: this step), l = the
// capped count of unambiguous bases ending here; emit(j - d) names the step whose minimizer is pushed.
// One scan over steps -(w-1)..-1 yields the window minimum m1 (rightmost on ties), its distance d1 and how many
// entries equal it; the minimum over -w..-1 (mp), the one over -(w-1)..0 (mn) and the number of identical-hash
// duplicates each case would flush follow from m1, X[-w] and X[0] without further loops.  The duplicate-flush loops
// themselves only run when such duplicates exist (tandem repeats).
// T = key type (uint64_t: hash<<8|span; uint32_t: the bare hash when 2k <= 30 and the span is the constant k — same order),
// WT = compile-time window length (0: use the run-time w).  The scan is branch-free.
template <class T, int WT, class XF, class Emit> __device__ __forceinline__ void sk_step_g(XF Xd, int j, int l, int w_rt, int k, Emit &emit)
{
    const int w = WT ? WT : w_rt;
    const T MAXV = ~(T)0;
    const T info = Xd(0), xw = Xd(w);
    T m1 = MAXV; int d1 = w - 1, c1 = 0;
#pragma unroll
    for (int d = w - 1; d >= 1; --d) {
        const T v = Xd(d);
        const bool lt = v < m1, eq = v == m1;
        m1 = lt ? v : m1;
        d1 = (lt || eq) ? d : d1;
        c1 = lt ? 1 : c1 + (eq ? 1 : 0);
    }
    if (w < 2) c1 = 0;
    // previous minimum over steps -w..-1 (rightmost on ties)
    const bool from_w = xw < m1;
    const T mp = from_w ? xw : m1;
    const int mpd = (from_w || w < 2) ? w : d1;
    if (l == w + k - 1 && mp != MAXV) {
        const int dup = from_w ? 0 : c1 - 1;      // other entries of -(w-1)..-1 equal to mp
        if (dup > 0) {
#pragma unroll 1
            for (int d = w - 1; d >= 1; --d)
                if (Xd(d) == mp && d != mpd) emit(j - d);
        }
    }
    if (info <= mp) {
        if (l >= w + k && mp != MAXV) emit(j - mpd);
    } else if (mpd == w) {
        if (l >= w + k - 1 && mp != MAXV) emit(j - mpd);
        // new minimum over steps -(w-1)..0 (rightmost on ties)
        const bool from_0 = info <= m1;
        const T mn = from_0 ? info : m1;
        const int mnd = from_0 ? 0 : d1;
        if (l >= w + k - 1 && mn != MAXV) {
            const int dup = info == m1 ? c1 : from_0 ? 0 : c1 - 1;
            if (dup > 0) {
#pragma unroll 1
                for (int d = w - 1; d >= 0; --d)
                    if (Xd(d) == mn && d != mnd) emit(j - d);
            }
        }
    }
}
template <class Emit> __device__ __forceinline__ void sk_step(const SkSmem &S, int j, int w, int k, Emit &emit)
{
    const uint64_t *X = S.X + SK_XH + j;        // X[0] = this step, X[-d] = d steps back
    sk_step_g<uint64_t, 0>([X](int d) { return X[-d]; }, j, (int)S.lcap[j], w, k, emit);
}

struct SkCount { int n; __device__ __forceinline__ void operator()(int) { ++n; } };
struct SkWrite {
    const SkSmem *S; uint64_t *ox; uint32_t *oy; int64_t at;
    __device__ __forceinline__ void operator()(int jj)
    {
        ox[at] = S->X[SK_XH + jj];
        oy[at] = (uint32_t)S->pos[SK_XH + jj] << 1 | S->z[SK_XH + jj];
        ++at;
    }
};

__device__ __forceinline__ int sk_load_code(const SketchArgs &A, const SeqDesc &sd, int i)
{
    if (i < 0 || i >= sd.len) return 4;
    if (sd.kind) return A.bytes[sd.off + i];
    int64_t p = sd.off + i;
    if ((A.nmask[p >> 5] >> (p & 31)) & 1) return 4;
    return (A.seq2[p >> 4] >> (2 * (p & 15))) & 3;
}

template <bool WRITE> __global__ void __launch_bounds__(SK_THREADS) k_sketch(SketchArgs A)
{
    __shared__ SkSmem S;
    const int tid = threadIdx.x;
    const int w = A.w, k = A.k;
    const uint64_t mask = (1ULL << 2 * k) - 1;
    const int shift1 = 2 * (k - 1);
    for (int seq = blockIdx.x; seq < A.n_seq; seq += gridDim.x) {
        const SeqDesc sd = A.seqs[seq];
        int n_steps = sd.len;
        uint8_t *hpc = nullptr; int32_t *hpp = nullptr; uint16_t *hpr = nullptr;
        if (A.hpc) {
            // ---- homopolymer compression: one step per run of identical bases, one per ambiguous base ----
            hpc = A.hp_code + (int64_t)blockIdx.x * A.hp_stride;
            hpp = A.hp_pos + (int64_t)blockIdx.x * A.hp_stride;
            hpr = A.hp_rl + (int64_t)blockIdx.x * A.hp_stride;
            int done = 0, prev_end = -1;
            for (int T0 = 0; T0 < sd.len; T0 += SK_TILE) {
                for (int i = tid; i < SK_TILE + 1; i += SK_THREADS) S.code[i] = (uint8_t)sk_load_code(A, sd, T0 + i);
                __syncthreads();
                // flags for 8 consecutive positions per thread
                int nf = 0; unsigned fm = 0;
                for (int c = 0; c < 8; ++c) {
                    int i = tid * 8 + c;
                    if (T0 + i < sd.len) {
                        int cc = S.code[i];
                        bool e = cc == 4 || T0 + i == sd.len - 1 || S.code[i + 1] != cc;
                        if (e) fm |= 1u << c, ++nf;
                    }
                }
                int tot, base = block_excl_scan(nf, &tot, S.ws);
                // previous run end for the first flagged position of this thread: search backwards
                // (steps are written with pos; run length = pos - previous step pos)
                for (int c = 0; c < 8; ++c)
                    if (fm >> c & 1) {
                        int i = tid * 8 + c;
                        hpc[done + base] = S.code[i];
                        hpp[done + base] = T0 + i;
                        ++base;
                    }
                done += tot;
                __syncthreads();
            }
            (void)prev_end;
            n_steps = done;
            __syncthreads();
            for (int m = tid; m < n_steps; m += SK_THREADS) {
                int prev = m ? hpp[m - 1] : -1;
                int r = hpp[m] - prev;
                hpr[m] = (uint16_t)(r > 65535 ? 65535 : r);
            }
            __syncthreads();
        }
        // ---- halo init ----
        for (int i = tid; i < SK_CH; i += SK_THREADS) S.code[i] = 4, S.rl[i] = 0;
        for (int i = tid; i < SK_XH; i += SK_THREADS) S.X[i] = ~0ULL, S.z[i] = 0, S.pos[i] = 0;
        int64_t out_run = 0;
        const int64_t out_base = WRITE ? A.offs[seq] : 0;
        __syncthreads();
        for (int T0 = 0; T0 < n_steps; T0 += SK_TILE) {
            const int tn = min(SK_TILE, n_steps - T0);
            // ---- (a) stage step codes ----
            if (A.hpc) {
                for (int i = tid; i < SK_TILE; i += SK_THREADS) {
                    bool in = i < tn;
                    S.code[SK_CH + i] = in ? hpc[T0 + i] : 4;
                    S.rl[SK_CH + i] = in ? hpr[T0 + i] : 0;
                    S.pos[SK_XH + i] = in ? hpp[T0 + i] : 0;
                }
            } else if (sd.kind == 0) {
                if (tid < SK_TILE / 64) {        // 64 bases per lane: one uint4 of bases + one uint2 of N bits
                    int64_t p = sd.off + T0 + (int64_t)tid * 64;
                    uint4 b4 = make_uint4(0, 0, 0, 0); uint2 n2 = make_uint2(0, 0);
                    if (T0 + tid * 64 < sd.len) {
                        b4 = *reinterpret_cast<const uint4 *>(A.seq2 + (p >> 4));
                        n2 = *reinterpret_cast<const uint2 *>(A.nmask + (p >> 5));
                    }
                    uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
                    uint64_t nm = (uint64_t)n2.y << 32 | n2.x;
                    uint32_t *dst = reinterpret_cast<uint32_t *>(&S.code[SK_CH + tid * 64]);
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        uint32_t wv = bw[q >> 2] >> (8 * (q & 3));
                        uint32_t out = 0;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            int bi = q * 4 + e;
                            uint32_t c = (wv >> (2 * e)) & 3;
                            if ((nm >> bi & 1) || T0 + tid * 64 + bi >= sd.len) c = 4;
                            out |= c << (8 * e);
                        }
                        dst[q] = out;
                    }
                }
                for (int i = tid; i < SK_TILE; i += SK_THREADS) S.pos[SK_XH + i] = T0 + i;
            } else {
                for (int i = tid; i < SK_TILE; i += SK_THREADS) {
                    S.code[SK_CH + i] = i < tn ? A.bytes[sd.off + T0 + i] : 4;
                    S.pos[SK_XH + i] = T0 + i;
                }
            }
            __syncthreads();
            // ---- (b) hashes: 8 consecutive steps per thread, rolling k-mers ----
            {
                const uint8_t *cd = S.code + SK_CH + tid * 8;
                const uint16_t *rl = S.rl + SK_CH + tid * 8;
                const int cap = w + k;
                int l = 0;
                for (int b = 1; b <= cap && b <= SK_CH; ++b) { if (cd[-b] == 4) break; l = b; }
                uint64_t k0 = 0, k1 = 0;
                int span = 0;
                {
                    int nb = l < k - 1 ? l : k - 1;
                    for (int b = nb; b >= 1; --b) {
                        uint64_t c = cd[-b];
                        k0 = (k0 << 2 | c) & mask;
                        k1 = (k1 >> 2) | (3ULL ^ c) << shift1;
                    }
                    if (A.hpc) { int ns = l < k ? l : k; for (int b = 1; b <= ns; ++b) span += rl[-b]; }
                }
                for (int c8 = 0; c8 < 8; ++c8) {
                    int j = tid * 8 + c8;
                    uint64_t X = ~0ULL; int z = 0;
                    int c = cd[c8];
                    if (c < 4) {
                        k0 = (k0 << 2 | (uint64_t)c) & mask;
                        k1 = (k1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
                        int sp;
                        if (A.hpc) {
                            span += rl[c8];
                            if (l >= k) span -= rl[c8 - k];      // queue already held k runs
                            sp = span;
                        } else sp = l + 1 < k ? l + 1 : k;
                        if (l < cap) ++l;
                        if (l >= k && sp < 256) {
                            z = k0 < k1 ? 0 : 1;
                            X = mix64_masked(z ? k1 : k0, mask) << 8 | (uint64_t)sp;
                        }
                    } else l = 0, span = 0;
                    if (j < tn) { S.X[SK_XH + j] = X; S.z[SK_XH + j] = (uint8_t)z; S.lcap[j] = (uint8_t)l; }
                    else { S.X[SK_XH + j] = ~0ULL; S.z[SK_XH + j] = 0; S.lcap[j] = 0; }
                }
            }
            __syncthreads();
            // ---- (c) emission, 256 steps at a time in order ----
            for (int sub = 0; sub < SK_TILE; sub += SK_THREADS) {
                if (sub >= tn) break;
                int j = sub + tid;
                SkCount cnt; cnt.n = 0;
                if (j < tn) sk_step(S, j, w, k, cnt);
                int tot, pre = block_excl_scan(cnt.n, &tot, S.ws);
                if (WRITE && cnt.n) {
                    SkWrite wr; wr.S = &S; wr.ox = A.mz_x; wr.oy = A.mz_y; wr.at = out_base + out_run + pre;
                    sk_step(S, j, w, k, wr);
                }
                out_run += tot;
            }
            __syncthreads();
            // ---- (d) carry halos ----
            if (T0 + SK_TILE < n_steps) {
                uint8_t c0 = 0; uint16_t r0 = 0; uint64_t x0 = 0; uint8_t z0 = 0; int32_t p0 = 0;
                if (tid < SK_CH) c0 = S.code[SK_TILE + tid], r0 = S.rl[SK_TILE + tid];
                if (tid < SK_XH) x0 = S.X[SK_TILE + tid], z0 = S.z[SK_TILE + tid], p0 = S.pos[SK_TILE + tid];
                __syncthreads();
                if (tid < SK_CH) S.code[tid] = c0, S.rl[tid] = r0;
                if (tid < SK_XH) S.X[tid] = x0, S.z[tid] = z0, S.pos[tid] = p0;
                __syncthreads();
            }
        }
        // ---- final flush: the pending minimum of the last window ----
        if (tid == 0 && n_steps > 0) {
            int last = (n_steps - 1) % SK_TILE;     // local index of the last step in the last tile
            const uint64_t *X = S.X + SK_XH + last;
            uint64_t mn = ~0ULL; int mnd = 0;
            for (int d = w - 1; d >= 0; --d) { uint64_t v = X[-d]; if (v <= mn) mn = v, mnd = d; }
            if (mn != ~0ULL) {
                if (WRITE) {
                    int jj = last - mnd;
                    A.mz_x[out_base + out_run] = mn;
                    A.mz_y[out_base + out_run] = (uint32_t)S.pos[SK_XH + jj] << 1 | S.z[SK_XH + jj];
                }
                ++out_run;
            }
            if (!WRITE) A.counts[seq] = (int32_t)out_run;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tile form of kernel (a) for the uncompressed presets (map-ont, map-hifi): one WARP per tile of 2048 steps, no CTA
// barrier, no recomputation.
//   * The warp stages the tile's 2-bit codes and N bits (plus 128 bases of lead-in) as packed words in shared
//     memory: 17 coalesced 128-byte rows for a read tile, or a byte->2-bit pack for a contig strand.
//   * Every iteration handles 32 consecutive steps, one per lane.  The k-mer of a step is cut out of the packed
//     stream with two funnel shifts (no rolling state): the forward k-mer is its 2-bit-group reversal (BREV), the
//     reverse-complement k-mer its complement.  The count of unambiguous bases behind the step is a CLZ of the
//     64 N bits that end at it.
//   * Hashes live in a two-block (previous 32 steps, current 32 steps) shared array; sk_step_g reads its w+1 neighbours from
//     it at fixed offsets (conflict-free: consecutive lanes, consecutive words) and records at most two pushes per step in
//     registers; one ballot orders them.  With 2k <= 30 (map-ont) the array holds the bare 32-bit hash and the k-mer is one
//     funnel shift; the window length is a template parameter for the two presets, so the scan is straight-line code.
//   * Minimizers go to a per-tile slot of a temporary array; k_sketch_compact packs the tiles into the CSR layout
//     after an exclusive scan of the tile counts.  Every base is read once and hashed once.
constexpr int SKT_TILE = 2048, SKT_LEAD = 128, SKT_WARPS = 8, SKT_CAP = SKT_TILE + 32;
constexpr int SKT_CW = (SKT_TILE + SKT_LEAD) / 16 + 2, SKT_NW = (SKT_TILE + SKT_LEAD) / 32 + 2;

struct SketchTileArgs {
    const uint32_t *seq2, *nmask;
    const uint8_t *bytes;
    const SeqDesc *seqs;
    int32_t n_seq, w, k, n_tiles;
    const int32_t *tile_first;        // [n_seq + 1] first tile of every sequence
    uint64_t *tmp_x; uint32_t *tmp_y; // [n_tiles * SKT_CAP]
    int32_t *tile_cnt;                // [n_tiles]
    // homopolymer-compressed presets: the step stream made by k_hpc_compress (one step per run / per ambiguous base)
    uint8_t *hp_code; int32_t *hp_pos;  // [sum of lengths]: code of the step, position of the run's last base
    const int64_t *hp_off;            // [n_seq + 1] slice of every sequence in hp_code / hp_pos
    int32_t *hp_n;                    // [n_seq] steps per sequence
    const int64_t *tile_off;          // [n_tiles + 1] (compact pass)
    int64_t *mz_off;                  // [n_seq + 1]   (compact pass)
    uint64_t *mz_x; uint32_t *mz_y;
};

template <class T> struct SktWarp { T X[64]; uint32_t cw[SKT_CW]; uint32_t nw[SKT_NW]; };   // X[0..31]: previous 32 steps, X[32..63]: current

struct SkRec {
    int n, t0, t1;
    __device__ __forceinline__ void operator()(int jj) { if (n == 0) t0 = jj; else if (n == 1) t1 = jj; ++n; }
};
// writes the minimizer of local step jj (jj in [c*32 - w, c*32 + 31]) into the tile's slot
template <class T> struct SkTileWrite {
    const T *X; uint64_t *ox; uint32_t *oy; int at, cap, c, T0, k; uint32_t zcur, zprev;
    const int32_t *hpos;       // compressed presets: position of every step's last base (nullptr: step = base)
    __device__ __forceinline__ void operator()(int jj)
    {
        if (at < cap) {
            const T x = X[32 + jj - c * 32];
            const int pos = hpos ? hpos[T0 + jj] : T0 + jj;
            ox[at] = sizeof(T) == 4 ? ((uint64_t)x << 8 | (uint64_t)k) : (uint64_t)x;
            oy[at] = (uint32_t)pos << 1 | ((((jj >> 5) == c ? zcur : zprev) >> (jj & 31)) & 1u);
        }
        ++at;
    }
};

// Homopolymer compression (map-pb): one warp per sequence walks it 32 bases at a time; a base ends a run when it is
// ambiguous, the last one, or differs from its successor; a ballot ranks the run ends and the warp appends (code, position).
__global__ void __launch_bounds__(256) k_hpc_compress(const __grid_constant__ SketchTileArgs A)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int nw = gridDim.x * (blockDim.x >> 5);
    SketchArgs L; L.seq2 = A.seq2; L.nmask = A.nmask; L.bytes = A.bytes;
    for (int seq = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); seq < A.n_seq; seq += nw) {
        const SeqDesc sd = A.seqs[seq];
        uint8_t *oc = A.hp_code + A.hp_off[seq]; int32_t *op = A.hp_pos + A.hp_off[seq];
        int done = 0;
        int cur = lane < sd.len ? sk_load_code(L, sd, lane) : 4;
        for (int i0 = 0; i0 < sd.len; i0 += 32) {
            const int i = i0 + lane;
            const int nxt_blk = sk_load_code(L, sd, i0 + 32 + lane);          // 4 beyond the end
            int nxt = __shfl_down_sync(FULL, cur, 1);
            const int first_next = __shfl_sync(FULL, nxt_blk, 0);
            if (lane == 31) nxt = first_next;
            const bool e = i < sd.len && (cur == 4 || i == sd.len - 1 || nxt != cur);
            const unsigned m = __ballot_sync(FULL, e);
            if (e) { const int at = done + __popc(m & ((1u << lane) - 1)); oc[at] = (uint8_t)cur; op[at] = i; }
            done += __popc(m);
            cur = nxt_blk;
        }
        if (lane == 0) A.hp_n[seq] = done;
    }
}

// K32: 2k <= 30, the ring holds the bare 30-bit hash (the span is the constant k, so the order is the same).
// WT: compile-time window length of the instance (0 = run-time w).
// HPC: the steps are the runs written by k_hpc_compress (tiles are laid out for the uncompressed length, the surplus ones
// are empty); the span of a k-mer is the distance between run ends, and a k-mer spanning 256 bases or more is dropped.
template <bool K32, int WT, bool HPC>
__global__ void __launch_bounds__(SKT_WARPS * 32) k_sketch_tiles(const __grid_constant__ SketchTileArgs A)
{
    static_assert(!(K32 && HPC), "compressed k-mers carry their span in the key");
    typedef typename std::conditional<K32, uint32_t, uint64_t>::type T;
    __shared__ SktWarp<T> SW[SKT_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    SktWarp<T> &W = SW[wid];
    const int w = WT ? WT : A.w, k = A.k, cap = w + k;
    const T MAXV = ~(T)0;
    const uint64_t mask = (1ULL << 2 * k) - 1;
    for (int tile = blockIdx.x * SKT_WARPS + wid; tile < A.n_tiles; tile += gridDim.x * SKT_WARPS) {
        // sequence of this tile: the last one whose first tile is <= tile
        int lo = 0, hi = A.n_seq;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (A.tile_first[mid] <= tile) lo = mid; else hi = mid; }
        const int seq = lo;
        SeqDesc sd = A.seqs[seq];
        const int32_t *hpos = nullptr;
        if (HPC) { sd.kind = 1; sd.off = A.hp_off[seq]; sd.len = A.hp_n[seq]; hpos = A.hp_pos + sd.off; }
        const uint8_t *bytes = HPC ? A.hp_code : A.bytes;
        const int T0 = (tile - A.tile_first[seq]) * SKT_TILE;
        const int tn = min(SKT_TILE, sd.len - T0);
        if (HPC && tn <= 0) { if (lane == 0) A.tile_cnt[tile] = 0; continue; }
        // ---- stage packed codes: stream index q <-> sequence position T0 - SKT_LEAD + q ----
        if (sd.kind == 0) {
            for (int m = lane; m < SKT_CW; m += 32) {
                const int i0 = T0 - SKT_LEAD + 16 * m;
                W.cw[m] = (m < SKT_CW - 2 && i0 >= 0 && i0 < sd.len) ? A.seq2[(sd.off + i0) >> 4] : 0u;
            }
            for (int m = lane; m < SKT_NW; m += 32) {
                const int i0 = T0 - SKT_LEAD + 32 * m;
                uint32_t v = 0xffffffffu;
                if (m < SKT_NW - 2 && i0 >= 0 && i0 < sd.len) {
                    v = A.nmask[(sd.off + i0) >> 5];
                    if (i0 + 32 > sd.len) v |= 0xffffffffu << (sd.len - i0);
                }
                W.nw[m] = v;
            }
        } else {
            for (int m0 = 0; m0 < SKT_CW + 31; m0 += 32) {        // every lane of the warp takes part in the shuffle
                const int m = m0 + lane, i0 = T0 - SKT_LEAD + 16 * m;
                uint32_t pk = 0, nb = 0;
                if (m < SKT_CW - 2) {
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const int i = i0 + t;
                        const uint32_t c = (i >= 0 && i < sd.len) ? bytes[sd.off + i] : 4u;
                        pk |= (c & 3u) << (2 * t);
                        nb |= (c > 3u ? 1u : 0u) << t;
                    }
                } else nb = 0xffffu;
                const uint32_t nb_hi = __shfl_xor_sync(FULL, nb, 1);
                if (m < SKT_CW) W.cw[m] = pk;
                if (!(lane & 1) && (m >> 1) < SKT_NW) W.nw[m >> 1] = nb | nb_hi << 16;
            }
        }
        __syncwarp();
        uint32_t zprev = 0, zcur = 0;
        int out_run = 0;
        const int nblk = (tn + 31) >> 5;
        uint64_t *ox = A.tmp_x + (int64_t)tile * SKT_CAP; uint32_t *oy = A.tmp_y + (int64_t)tile * SKT_CAP;
        const T *Xme = W.X + 32 + lane;            // X of this lane's step; Xme[-d] = d steps back
        T Xprev = MAXV;
        for (int c = -1; c < nblk; ++c) {
            const int j = c * 32 + lane, s = j + SKT_LEAD;
            // unambiguous bases ending at s: CLZ of the 64 N bits [s-63, s]
            int l;
            {
                const int e = s - 63, wi = e >> 5, sh = e & 31;
                const uint32_t n0 = W.nw[wi], n1 = W.nw[wi + 1], n2 = W.nw[wi + 2];
                const uint32_t wl = __funnelshift_r(n0, n1, sh), wh = __funnelshift_r(n1, n2, sh);
                const int run = wh ? __clz((int)wh) : 32 + __clz((int)wl);
                l = run < cap ? run : cap;
            }
            T X = MAXV; int z = 0;
            if (l >= k) {
                const int sb = s - k + 1, wi = sb >> 4, sh = 2 * (sb & 15);
                const uint32_t a = W.cw[wi], b = W.cw[wi + 1], cc = W.cw[wi + 2];
                if (K32) {
                    const uint32_t m32 = (uint32_t)mask;
                    const uint32_t val = __funnelshift_r(a, b, sh) & m32;                 // 2k <= 30 bits: one funnel shift
                    uint32_t r = __brev(val);
                    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
                    const uint32_t k0 = r >> (32 - 2 * k), k1 = ~val & m32;
                    z = k0 < k1 ? 0 : 1;
                    uint32_t key = z ? k1 : k0;
                    key = (~key + (key << 21)) & m32;
                    key = key ^ key >> 24;
                    key = (key * 265u) & m32;
                    key = key ^ key >> 14;
                    key = (key * 21u) & m32;
                    key = key ^ key >> 28;
                    key = (key + (key << 31)) & m32;
                    X = (T)key;
                } else {
                    const uint64_t val = ((uint64_t)__funnelshift_r(b, cc, sh) << 32 | __funnelshift_r(a, b, sh)) & mask;
                    uint64_t r = __brevll(val);
                    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
                    const uint64_t k0 = r >> (64 - 2 * k), k1 = ~val & mask;
                    z = k0 < k1 ? 0 : 1;
                    int sp = k;
                    if (HPC) { const int jj = T0 + j; sp = hpos[jj] - (jj - k >= 0 ? hpos[jj - k] : -1); }
                    if (!HPC || sp < 256) X = (T)(mix64_masked(z ? k1 : k0, mask) << 8 | (uint64_t)sp);
                }
            }
            zprev = zcur; zcur = __ballot_sync(FULL, z);
            W.X[lane] = Xprev; W.X[32 + lane] = X; Xprev = X;
            __syncwarp();
            if (c >= 0) {
                SkRec rec; rec.n = 0; rec.t0 = 0; rec.t1 = 0;
                if (j < tn) sk_step_g<T, WT>([Xme](int d) { return Xme[-d]; }, j, l, w, k, rec);
                const unsigned m1 = __ballot_sync(FULL, rec.n > 0);
                int pre, tot;
                if (!__any_sync(FULL, rec.n > 1)) { pre = __popc(m1 & ((1u << lane) - 1)); tot = __popc(m1); }
                else {
                    int x = rec.n;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(FULL, x, d); if (lane >= d) x += y; }
                    tot = __shfl_sync(FULL, x, 31); pre = x - rec.n;
                }
                if (rec.n) {
                    SkTileWrite<T> wr; wr.X = W.X; wr.ox = ox; wr.oy = oy; wr.at = out_run + pre; wr.cap = SKT_CAP; wr.c = c; wr.T0 = T0; wr.k = k; wr.zcur = zcur; wr.zprev = zprev; wr.hpos = hpos;
                    if (rec.n <= 2) { wr(rec.t0); if (rec.n == 2) wr(rec.t1); }
                    else sk_step_g<T, WT>([Xme](int d) { return Xme[-d]; }, j, l, w, k, wr);
                }
                out_run += tot;
            }
            __syncwarp();
        }
        // ---- the pending minimum of the sequence's last window ----
        if (lane == 0) {
            if (tn > 0 && T0 + tn == sd.len) {
                const int last = tn - 1, c = nblk - 1;
                const T *Xl = W.X + 32 + (last - c * 32);
                T mn = MAXV; int mnd = 0;
                for (int d = w - 1; d >= 0; --d) { const T v = Xl[-d]; if (v <= mn) mn = v, mnd = d; }
                if (mn != MAXV) {
                    SkTileWrite<T> wr; wr.X = W.X; wr.ox = ox; wr.oy = oy; wr.at = out_run; wr.cap = SKT_CAP; wr.c = c; wr.T0 = T0; wr.k = k; wr.zcur = zcur; wr.zprev = zprev; wr.hpos = hpos;
                    wr(last - mnd);
                    ++out_run;
                }
            }
            A.tile_cnt[tile] = out_run < SKT_CAP ? out_run : SKT_CAP;
        }
        __syncwarp();
    }
}

// pack the per-tile slots into the CSR minimizer arrays; one warp per tile.  Also publishes the per-sequence offsets.
__global__ void __launch_bounds__(256) k_sketch_compact(const __grid_constant__ SketchTileArgs A)
{
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= A.n_seq; i += gridDim.x * blockDim.x)
        A.mz_off[i] = A.tile_off[A.tile_first[i]];
    for (int tile = gw; tile < A.n_tiles; tile += nw) {
        const int64_t base = A.tile_off[tile];
        const int cnt = (int)(A.tile_off[tile + 1] - base);
        const uint64_t *sx = A.tmp_x + (int64_t)tile * SKT_CAP; const uint32_t *sy = A.tmp_y + (int64_t)tile * SKT_CAP;
        for (int i = lane; i < cnt; i += 32) { A.mz_x[base + i] = sx[i]; A.mz_y[base + i] = sy[i]; }
    }
}

}  // namespace telr
