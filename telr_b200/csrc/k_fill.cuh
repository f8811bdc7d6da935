// k_fill.cuh — fast path of kernel (d) for approximate-max GLOBAL gap fills (82 % of all DP cells on config 2):
// `ksw_extd2_sse(..., w >= max(qlen,tlen), flag = KSW_EZ_APPROX_MAX)` as issued by mm_align1's gap-filling loop
// (minimap2 align.c; reference call site TELR_te.py:505).
//
// Mapping.  The DP state never leaves registers.  Lane l owns FC = 8 consecutive target columns of a 256-column
// pass; it walks the query two rows at a time (row pair m at step m + l, a systolic skew of one pair per lane).
// Inside a lane-step the two rows are swept left to right with the second row one column behind the first, so the
// two cells handled together are independent and sit in the two 16-bit halves of every register:
//     lo half: cell (column k,   row j)      hi half: cell (column k-1, row j+1)
// Direction information is 8 bits per cell (which of s,a,b,a2 is below the maximum; which gap states do not continue);
// one 8-byte store per row per lane writes them row-major.
// Left-to-right values (v,x,x2) travel between lanes by shuffle, between 256-column passes through a small
// per-warp boundary array.  No tensor cores: this is not a dense contraction.
//
// Number representation (round 2, "offset form").  The measured issue budget of an SM sub-partition is 0.5 instructions per clock on
// the ALU pipe (PRMT, LOP3, packed min/max, IADD3) and 0.5 on the FMA pipe (IMAD, IMAD.IADD), ~0.95 together
// (profiles/ubench/pipes2.cu); sm_100a has no packed 16x2 subtract (`__vsub2` = LOP3 + 2 x VIADD.16x2) and the recurrence subtracts
// seven times per cell pair.  Every stored half-word therefore carries an offset that keeps it NON-NEGATIVE, so that plain 32-bit
// adds and subtracts act on the two halves independently — no borrow ever crosses bit 16:
//     u, v         + FB                      (FB = 60; the differences are bounded by the gap costs)
//     s, z, a, b.. + 2 FB                    (the score table holds s + 2 FB as a positive int8)
//     x, y         + (q + e - 1) + 0x8000,   x2, y2 + (q2 + e2 - 1) + 0x8000
// With that offset a gap state equals 0x7fff exactly when it was reset to -(q + e) and has bit 15 set exactly when it continues:
// the continuation flag is the stored value's top bit (no compare), and "t - z + 0x8000" has bit 15 set exactly when t is the
// maximum.  The flag bytes are gathered by one top-bit-replicating PRMT per flag and a tree of multiply-adds (see the cell loop).
// Per cell pair: 4 IADD3 + 2 VIMNMX3.S16x2 + 4 VIADDMNMX.U16x2 + 8 PRMT on the ALU pipe, 12 IMAD / IMAD.IADD on the FMA pipe
// (21 ALU-pipe operations + 18 VIADD.16x2 in round 1).  tests/test_offset_form.py replays this arithmetic on the host.
#pragma once
#include <cuda_runtime.h>
#include "mm_align.cuh"

namespace telr {

#ifndef TELR_FILL_FC12
#define TELR_FILL_FC12 1
#endif
constexpr int FC_MAX = 12;            // columns per lane: 8 (256-column passes) or 12 (384-column passes)
#ifndef TELR_FILL_CADD
#define TELR_FILL_CADD 1              // 1: two-input adds left to the compiler's pipe choice (mostly IMAD.IADD); 0: forced IMADs (1 % slower)
#endif
constexpr int FB = 60;                // offset of the stored differences (see "Number representation" above)

// prmt.b32 in default mode: selector nibble bit 3 replicates the sign of the selected byte (__byte_perm drops that bit)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// a * m + c as an integer multiply-add: IMAD issues on the FMA pipe, which the recurrence's permutes, min/max ops and IADD3s
// (all on the ALU pipe) leave idle.  The multiplier is a run-time register so that ptxas keeps the IMAD
// (profiles/ubench/pipes2.cu: IMAD + PRMT 0.92 warp-instructions per clock and sub-partition, IADD3 + PRMT 0.50).
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t m, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
    return d;
}
// 1 that the compiler cannot see through (gridDim.z of every launch in this library)
__device__ __forceinline__ uint32_t opaque_one() { uint32_t d; asm volatile("mov.u32 %0, %%nctaid.z;" : "=r"(d)); return d; }
// sum_f r[f] 2^f as a depth-3 tree of IMADs (a Horner chain is 7 dependent operations)
__device__ __forceinline__ uint32_t flag_sum(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t r4, uint32_t r5, uint32_t r6, uint32_t r7,
                                             uint32_t TWO, uint32_t FOUR, uint32_t SIXTEEN)
{
    const uint32_t a = imad(r1, TWO, r0), b = imad(r3, TWO, r2), c = imad(r5, TWO, r4), d = imad(r7, TWO, r6);
    return imad(imad(d, FOUR, c), SIXTEEN, imad(b, FOUR, a));
}
#if TELR_FILL_CADD
#define FADD(a, b) ((a) + (b))
#define FSUB(a, b) ((a) - (b))
#else
#define FADD(a, b) imad((a), ONE, (b))
#define FSUB(a, b) imad((b), MONE, (a))
#endif
__device__ __forceinline__ uint32_t pk2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ uint32_t pk1(int v) { return pk2(v, v); }
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(int16_t)(v >> 16); }

// sum of the first n boundary differences f(0..n-1) of ksw_extd2 (first row / first column)
__device__ __forceinline__ int bnd_sum(int n, int qe, int e, int e2, int LT, int LD)
{
    if (n <= 0) return 0;
    int s = -qe;                                           // r = 0
    int n1 = (n - 1 < LT - 1 ? n - 1 : LT - 1);            // r = 1 .. LT-1
    if (n1 > 0) s -= e * n1;
    if (n - 1 >= LT && LT >= 1) s += LD;                    // r = LT
    int n3 = n - 1 - (LT > 0 ? LT : 0);                     // r = LT+1 .. n-1
    if (n3 > 0) s -= e2 * n3;
    return s;
}

// direction rows are padded to a multiple of 24 bytes so that both lane widths store whole lanes
__host__ __device__ __forceinline__ int fill_stride(int tlen) { return (tlen + 23) / 24 * 24; }

// true when the request can take the fast path (shape / flags); ambiguous bases are checked inside
__device__ __forceinline__ bool fill_fast_ok(const DpTask &T)
{
    if (T.kind != 0 || T.flag != KSW_APPROX_MAX || T.qstep != 1 || T.tstep != 1 || T.qcomp) return false;
    if (T.qlen < 1 || T.tlen < 1) return false;
    int mx = T.qlen > T.tlen ? T.qlen : T.tlen;
    return T.w < 0 || T.w >= mx;
}

// 12 columns per lane when that saves a pass (e.g. 257..384 target bases: one pass instead of two)
__host__ __device__ __forceinline__ int fill_width(int tlen)
{
    const int p8 = (tlen + 255) / 256, p12 = (tlen + 383) / 384;
    return p12 * 13 < p8 * 9 ? 12 : 8;
}

// forward pass for one lane width; the caller has excluded ambiguous bases
template <int FC>
__device__ void warp_fill_fwd(const Opt &o, const DpTask &T, DpRes &R, uint8_t *dir, uint32_t *bnd, unsigned long long *cells_acc)
{
    constexpr int FW = 32 * FC;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int qlen = T.qlen, tlen = T.tlen;
    const uint8_t *__restrict__ Q = T.q, *__restrict__ Tg = T.t;
    int q = o.q, e = o.e, q2 = o.q2, e2 = o.e2;
    if (q2 + e2 < q + e) { int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }
    const int qe = q + e, qe2 = q2 + e2;
    int LT = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
    if (q2 + e2 + LT * e2 > q + e + LT * e) ++LT;
    const int LD = LT * (e - e2) - (q2 - q) - e2;
#define BNDF(r) (FB + ((r) == 0 ? -qe : (r) < LT ? -e : (r) == LT ? LD : -e2))
    const int BX = qe - 1 + 0x8000, BX2 = qe2 - 1 + 0x8000;                     // offsets of x, y and of x2, y2
    const uint32_t CA = (uint32_t)(FB - BX) * 65537u, CA2 = (uint32_t)(FB - BX2) * 65537u;      // a = x + v + CA: exact on both halves
    const uint32_t QM1 = pk1(q - 1), Q2M1 = pk1(q2 - 1), XFLOOR = 0x7fff7fffu, DK = 0x80008000u;
    const uint32_t X0 = pk1(BX - qe), X20 = pk1(BX2 - qe2);                     // a gap state that has just been reset (= XFLOOR)
    // score table of one target base: byte c = score against query base c, + 2 FB (a positive int8)
    const uint32_t TS_MIS = 0x01010101u * (uint32_t)(2 * FB - o.b), TS_FLIP = (uint32_t)(2 * FB + o.a) ^ (uint32_t)(2 * FB - o.b);
    const uint32_t ONE = opaque_one(), MONE = 0u - ONE, TWO = ONE + ONE, FOUR = TWO + TWO, SIXTEEN = FOUR * FOUR, K01 = 0x01010101u * ONE;
    const int stride = fill_stride(tlen);
    const int npairs = (qlen + 1) >> 1, npass = (tlen + FW - 1) / FW;
    int usum = 0;
    if (lane == 0) { res_reset(R); }
    for (int pass = 0; pass < npass; ++pass) {
        const int t0 = pass * FW + lane * FC;
        const bool live = t0 < tlen;
        uint32_t TS[FC], Uu[FC], Uy[FC], Uy2[FC];
#pragma unroll
        for (int k = 0; k < FC; ++k) {
            const int tc = t0 + k < tlen ? Tg[t0 + k] : 0;
            TS[k] = TS_MIS ^ (TS_FLIP << (8 * tc));
        }
#pragma unroll
        for (int k = 0; k < FC; ++k) {              // top boundary sits in the hi halves (row "-1" of pair 0)
            const int r = t0 + k;
            Uu[k] = pk2(FB, BNDF(r)); Uy[k] = X0; Uy2[k] = X20;
        }
        uint32_t inV = 0, inX = 0, inX2 = 0, Ufirst = 0;
        // query bases of the lane's next row pair (q[j] | q[j+1] << 8), loaded one step ahead of their use
        const uint32_t BND_TAIL = pk1(FB - e2);      // boundary differences beyond row LT
        const bool last_pass = pass == npass - 1;
        const int nlive = tlen - pass * FW >= FW ? 32 : (tlen - pass * FW + FC - 1) / FC;      // lanes that own columns in this pass
        for (int s = 0; s < npairs + nlive - 1; ++s) {
            const int m = s - lane;
            const bool active = live && m >= 0 && m < npairs;
            if (lane == 0 && s < npairs) {
                const int j = 2 * s;
                if (pass == 0) {
                    inV = j > LT ? BND_TAIL : pk2(BNDF(j), BNDF(j + 1)); inX = X0; inX2 = X20;
                } else { inV = bnd[3 * s]; inX = bnd[3 * s + 1]; inX2 = bnd[3 * s + 2]; }
            }
            uint32_t outV = 0, outX = 0, outX2 = 0;
            if (active) {
                const int j = 2 * m;
                // the two query bases of this row pair become the PRMT selector that looks them up in the score tables:
                // lo half <- sign-extended byte q[j] of the first table, hi half <- byte q[j+1] of the second
                const uint32_t q0 = Q[j], q1 = j + 1 < qlen ? Q[j + 1] : 0;
                const uint32_t inQ = q0 | (q0 | 8u) << 4 | (q1 + 4u) << 8 | (q1 + 12u) << 12;
                uint32_t Lv = inV, Lx = inX, Lx2 = inX2, pu = pk1(FB), py = X0, py2 = X20, kV = 0, kX = 0, kX2 = 0;
                uint32_t W[FC / 2 + 1];             // W[p]: flag bytes of cells (2p, j), (2p-1, j+1), (2p+1, j), (2p, j+1)
                uint32_t eDS = 0, eDA = 0, eDB = 0, eDA2 = 0, eNX = 0, eNY = 0, eNX2 = 0, eNY2 = 0;   // even iteration, waiting for its partner
#pragma unroll
                for (int k = 0; k <= FC; ++k) {
                    const int kk = k < FC ? k : FC - 1, kh = k > 0 ? k - 1 : 0;
                    // up inputs: lo <- previous row pair's second row (hi half of the saved register), hi <- cell above in this pair
                    const uint32_t up_u = __byte_perm(Uu[kk], pu, 0x5432), up_y = __byte_perm(Uy[kk], py, 0x5432), up_y2 = __byte_perm(Uy2[kk], py2, 0x5432);
                    if (k == 1) {                   // second row enters: its left neighbour is the previous lane's last column
                        Lv = __byte_perm(Lv, inV, 0x7610); Lx = __byte_perm(Lx, inX, 0x7610); Lx2 = __byte_perm(Lx2, inX2, 0x7610);
                    }
                    const uint32_t S = prmt(TS[kk], TS[kh], inQ);
                    const uint32_t A = Lx + Lv + CA, A2 = Lx2 + Lv + CA2, B = up_y + up_u + CA, B2 = up_y2 + up_u + CA2;
                    uint32_t Z = __vimax3_s16x2(S, A, B);
                    Z = __vimax3_s16x2(Z, A2, B2);
                    // t - z + 0x8000 per half: bit 15 set exactly when t is the maximum (two-input adds: FMA pipe)
                    const uint32_t NZ = FSUB(DK, Z);
                    const uint32_t DS = FADD(S, NZ), DA = FADD(A, NZ), DB = FADD(B, NZ), DA2 = FADD(A2, NZ), DB2 = FADD(B2, NZ);
                    const uint32_t nu = FSUB(Z, Lv), nv = FSUB(Z, up_u);
                    // max(t - z - e, -(q + e)) in the offset form: bit 15 set exactly when the gap continues
                    const uint32_t nx = __viaddmax_u16x2(DA, QM1, XFLOOR), ny = __viaddmax_u16x2(DB, QM1, XFLOOR);
                    const uint32_t nx2 = __viaddmax_u16x2(DA2, Q2M1, XFLOOR), ny2 = __viaddmax_u16x2(DB2, Q2M1, XFLOOR);
                    // 8 top bits per cell -> one byte per cell (a flag is the COMPLEMENT of the top bit: "below the maximum",
                    // "does not continue").  Two iterations are gathered together: one PRMT per flag replicates the top bits of
                    // (k lo, k hi, k+1 lo, k+1 hi) into the bytes r_f of a word (0x00 / 0xff).  sum_f r_f 2^f = 255 T (mod 2^32),
                    // T = the word of top-bit bytes (each 0xff byte is 256 - 1: the carries telescope), so the flag word
                    // ~T = -1 - 255 T / 255 = (sum_f r_f 2^f) * 0x01010101 - 1: a tree of multiply-adds and one multiply, all IMADs.
                    if (k == FC) {                  // last iteration (even): only its hi cell (FC-1, j+1) exists
                        uint32_t acc = ~prmt(DS, DA, 0xFDB9) & 0x02020101u;
                        acc |= ~prmt(DB, DA2, 0xFDB9) & 0x08080404u;
                        acc |= ~prmt(nx, ny, 0xFDB9) & 0x20201010u;
                        acc |= ~prmt(nx2, ny2, 0xFDB9) & 0x80804040u;
                        W[FC / 2] = acc | (acc >> 16);          // byte 1
                    } else if (!(k & 1)) {
                        eDS = DS; eDA = DA; eDB = DB; eDA2 = DA2; eNX = nx; eNY = ny; eNX2 = nx2; eNY2 = ny2;
                    } else {
                        const uint32_t acc = flag_sum(prmt(eDS, DS, 0xFDB9), prmt(eDA, DA, 0xFDB9), prmt(eDB, DB, 0xFDB9), prmt(eDA2, DA2, 0xFDB9),
                                                      prmt(eNX, nx, 0xFDB9), prmt(eNY, ny, 0xFDB9), prmt(eNX2, nx2, 0xFDB9), prmt(eNY2, ny2, 0xFDB9),
                                                      TWO, FOUR, SIXTEEN);
                        W[k >> 1] = imad(acc, K01, 0xffffffffu);
                    }
                    if (k >= 1) { Uu[k - 1] = nu; Uy[k - 1] = ny; Uy2[k - 1] = ny2; }      // hi halves: cell (k-1, j+1) = up input of the next row pair
                    if (k == 0) Ufirst = nu;
                    if (k == FC - 1) { kV = nv; kX = nx; kX2 = nx2; }
                    pu = nu; py = ny; py2 = ny2; Lv = nv; Lx = nx; Lx2 = nx2;
                }
                outV = __byte_perm(kV, Lv, 0x7610); outX = __byte_perm(kX, Lx, 0x7610); outX2 = __byte_perm(kX2, Lx2, 0x7610);
                if (t0 < stride) {
                    uint8_t *r0 = dir + (int64_t)j * stride + t0, *r1 = r0 + stride;
                    uint32_t w0[FC / 4], w1[FC / 4];
#pragma unroll
                    for (int w = 0; w < FC / 4; ++w) {
                        w0[w] = __byte_perm(W[2 * w], W[2 * w + 1], 0x6420);                                        // row j: cells (4w .. 4w+3, j)
                        w1[w] = __byte_perm(__byte_perm(W[2 * w], W[2 * w + 1], 0x0753), W[2 * w + 2], 0x5210);      // row j+1
                    }
                    if (FC == 8) {
                        *reinterpret_cast<uint2 *>(r0) = make_uint2(w0[0], w0[1]);
                        if (j + 1 < qlen) *reinterpret_cast<uint2 *>(r1) = make_uint2(w1[0], w1[1]);
                    } else {
#pragma unroll
                        for (int w = 0; w < FC / 4; ++w) {
                            reinterpret_cast<uint32_t *>(r0)[w] = w0[w];
                            if (j + 1 < qlen) reinterpret_cast<uint32_t *>(r1)[w] = w1[w];
                        }
                    }
                }
                if (lane == 31 && !last_pass) { bnd[3 * m] = outV; bnd[3 * m + 1] = outX; bnd[3 * m + 2] = outX2; }
            }
            // systolic hand-over to the next lane (used by it in the next step)
            const uint32_t sV = __shfl_up_sync(FULL, outV, 1), sX = __shfl_up_sync(FULL, outX, 1), sX2 = __shfl_up_sync(FULL, outX2, 1);
            if (lane > 0) { inV = sV; inX = sX; inX2 = sX2; }
        }
        // u of the last query row for this lane's columns (score = boundary + sum of u along the last row)
        if (live) {
#pragma unroll
            for (int k = 0; k < FC; ++k) {
                if (t0 + k >= tlen) continue;
                if (qlen & 1) usum += (k == 0 ? lo16(Ufirst) : lo16(Uu[k - 1])) - FB;     // last real row = first row of the last pair
                else usum += hi16(Uu[k]) - FB;
            }
        }
        __syncwarp();       // bnd[] written by lane 31 is read by lane 0 in the next pass
    }
#undef BNDF
#pragma unroll
    for (int d = 16; d; d >>= 1) usum += __shfl_xor_sync(FULL, usum, d);
    if (lane == 0) {
        R.score = bnd_sum(qlen, qe, e, e2, LT, LD) + usum;
        atomicAdd(cells_acc, (unsigned long long)qlen * (unsigned long long)tlen);
    }
    __syncwarp();
}

// forward pass; returns false (nothing written) when an ambiguous base is present
__device__ bool warp_fill_fast(const Opt &o, const DpTask &T, DpRes &R, uint8_t *dir, uint32_t *bnd, unsigned long long *cells_acc, bool wide_ok = false)
{
    const int lane = threadIdx.x & 31;
    {   // ambiguous bases take the general path (their score is not match/mismatch)
        bool n = false;
        for (int i = lane; i < T.qlen; i += 32) n |= T.q[i] > 3;
        for (int i = lane; i < T.tlen; i += 32) n |= T.t[i] > 3;
        if (__any_sync(0xffffffffu, n)) return false;
    }
    {   // the offset representation needs every stored half-word in [0, 2 FB + a] (score table = positive int8)
        const int qm = o.q + o.e > o.q2 + o.e2 ? o.q + o.e : o.q2 + o.e2;
        if (o.a < 0 || o.b < 0 || 2 * FB + o.a > 127 || o.a + o.b + qm > FB) return false;
    }
#if TELR_FILL_FC12
    // the 12-column instance halves the passes of 257..384-column fills but doubles the hot code: it pays only where it has SMs
    // of its own (k_al_queue, role 2); sharing an instruction cache with the 8-column instance it costs 35 % (profiles/README.md)
    if (wide_ok && fill_width(T.tlen) == 12) warp_fill_fwd<12>(o, T, R, dir, bnd, cells_acc);
    else
#endif
    warp_fill_fwd<8>(o, T, R, dir, bnd, cells_acc);
    return true;
}

// ksw_backtrack over the fast path's row-major direction bytes (global alignment, no band).
// Warp-cooperative and run-at-a-time: the lanes stage a window of 56 rows x 64 columns that ends at the current cell
// into shared memory (coalesced 8-byte loads, two rows per lane); then every step resolves a whole CIGAR run: lane k
// looks at the k-th cell ahead along the current direction (diagonal, left or up), one ballot finds where the run stops.
// All lanes carry the same (i, j, state, open run); only lane 0 writes CIGAR words.
constexpr int TBW = 64, TBR = 56, TBP = 72;        // window columns, rows, row pitch in bytes (72: column walks hit 16 banks)
struct TbSmem { uint2 w[TBR][TBP / 8]; };

__device__ void fill_traceback(const DpTask &T, DpRes &R, const uint8_t *dir, uint32_t *ezcig, int ezcap, int32_t *err, TbSmem &tb)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int qlen = T.qlen, tlen = T.tlen, stride = fill_stride(tlen);
    uint32_t *c = ezcig; int n = 0;
    int i = tlen - 1, j = qlen - 1, state = 0;
    int run_op = -1; uint32_t run_len = 0;          // the open CIGAR run lives in registers; memory is touched once per run
#define FLUSH() do { if (run_len) { if (n < ezcap && lane == 0) c[n] = run_len << 4 | (uint32_t)run_op; ++n; } } while (0)
#define PUSH(op, len) do { if ((op) == run_op) run_len += (uint32_t)(len); else { FLUSH(); run_op = (op); run_len = (uint32_t)(len); } } while (0)
    const uint8_t *wb = reinterpret_cast<const uint8_t *>(&tb.w[0][0]);
    while (i >= 0 && j >= 0) {
        const int c0 = (i - (TBW - 8)) > 0 ? ((i - (TBW - 8)) & ~7) : 0;      // window columns [c0, c0+64), rows [jtop, j0]
        const int j0 = j, jtop = j - (TBR - 1) > 0 ? j - (TBR - 1) : 0;
        {   // Every load first, then the stores: through generic pointers the compiler must assume that a store to the window may
            // alias the next load, and would wait for each load in turn (16 dependent L2 round trips per window and lane).
            static_assert(TBR <= 64, "two window rows per lane");
            uint2 t[2][TBW / 8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = lane + 32 * h, row = j0 - rr;
                const uint8_t *src = dir + (int64_t)row * stride + c0;
#pragma unroll
                for (int k = 0; k < TBW / 8; ++k)
                    t[h][k] = rr < TBR && row >= 0 && c0 + 8 * k < stride ? *reinterpret_cast<const uint2 *>(src + 8 * k) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = lane + 32 * h;
                if (rr < TBR) {
#pragma unroll
                    for (int k = 0; k < TBW / 8; ++k) tb.w[rr][k] = t[h][k];
                }
            }
        }
        __syncwarp();
        while (i >= c0 && j >= jtop) {
            if (state == 0) {           // diagonal run: cells (i-k, j-k) while "s is the maximum"
                const int ii = i - lane, jj = j - lane;
                const bool ok = ii >= c0 && jj >= jtop;
                const uint32_t b = ok ? wb[(j0 - jj) * TBP + (ii - c0)] : 1u;
                const unsigned m = __ballot_sync(FULL, b & 1u);
                const int L = m ? __ffs(m) - 1 : 32;
                if (L > 0) { PUSH(0, L); i -= L; j -= L; continue; }
                const uint32_t b0 = __shfl_sync(FULL, b, 0);      // a gap opens here
                state = !(b0 & 2) ? 1 : !(b0 & 4) ? 2 : !(b0 & 8) ? 3 : 4;
                if (state == 1 || state == 3) { PUSH(2, 1); --i; } else { PUSH(1, 1); --j; }
            } else {                    // gap run: cells to the left (deletion) or above (insertion) while the gap state continues
                const bool del = state == 1 || state == 3;
                const int ii = del ? i - lane : i, jj = del ? j : j - lane;
                const bool ok = ii >= c0 && jj >= jtop;
                const uint32_t b = ok ? wb[(j0 - jj) * TBP + (ii - c0)] : 0xffu;
                const unsigned m = __ballot_sync(FULL, (b >> (3 + state)) & 1u);
                const unsigned okm = __ballot_sync(FULL, ok);
                const int L = m ? __ffs(m) - 1 : 32;
                if (L > 0) { PUSH(del ? 2 : 1, L); if (del) i -= L; else j -= L; }
                if (L < 32 && ((okm >> L) & 1u)) state = 0;       // the gap ends on a cell inside the window; its own direction is read next
            }
        }
        __syncwarp();
    }
    if (i >= 0) PUSH(2, i + 1);
    if (j >= 0) PUSH(1, j + 1);
    FLUSH();
#undef PUSH
#undef FLUSH
    __syncwarp();
    if (n > ezcap) { if (lane == 0) atomicOr(err, TELR_ERR_CIGCAP); n = 0; }
    for (int k = lane; k < n >> 1; k += 32) { uint32_t t = c[k]; c[k] = c[n - 1 - k]; c[n - 1 - k] = t; }
    if (lane == 0) { R.n_cigar = n; R.cigar = ezcig; R.reach_end = 0; }
    __syncwarp();
}

}  // namespace telr
