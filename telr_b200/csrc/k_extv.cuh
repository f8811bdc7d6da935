// k_extv.cuh — vectorised form of the general two-piece-affine DP (kernel (d)): extensions with exact max tracking and
// z-drop, banded or not, left- or right-aligned gaps (ksw_extd2_sse as called by mm_align1; reference call site
// TELR_te.py:505).  Anti-diagonal sweep over a shared-memory circular window of the six difference arrays.  Every lane
// owns a group of 4 consecutive columns per step: state is moved as 32-bit words (4 x int8), unpacked to two 16x2
// registers and run through the packed recurrence of k_fill.cuh.  Sequence codes are staged 128 at a time as packed
// 2-bit streams (the query reversed, so that an anti-diagonal reads both forwards); the XOR of 4+4 codes indexes a
// shared table of packed match/mismatch scores.  (An 8-column variant executes fewer instructions but its larger loop
// body costs more in instruction-cache misses than it saves: measured 12 % slower end to end.)  In exact mode the running H row is updated in the same pass and the
// row maximum (with the reference's tie order) is found with one integer key per cell and a single warp reduction.
// The shared-memory state is held in the offset form of k_fill.cuh — bytes u + FB, v + FB, x + (q + e), y + (q + e),
// x2 + (q2 + e2), y2 + (q2 + e2), all non-negative — so that the recurrence runs on 32-bit adds (IADD3 / IMAD) without packed
// negations, and the flag gather is the Horner form described there.
// Direction bytes use the sign-bit format:
//   bits 0-3: "below the maximum" for (s,a,b,a2) [left-aligned gaps] or (b2,a,b,a2) [right-aligned]
//   bits 4-7: "gap does not continue" for x,y,x2,y2
// Requests that contain an ambiguous base, or whose band does not fit the window, return false and take the scalar path.
#pragma once
#include "k_fill.cuh"

namespace telr {

constexpr int VSC = 1024;             // shared-memory state window (columns), power of two
constexpr int VCW = 1024;             // window of staged sequence codes (2 bits each), power of two
constexpr int VEC_MAX_NCOL = VSC - 136;   // widest live band: leaves room for one 128-code staging step + group slack (default band: 752)
__device__ __forceinline__ int vwr(int i) { return i & (VSC - 1); }   // window slot of column i
struct VecSmem { int8_t st[6][VSC]; int32_t H[VSC]; uint8_t tb[VCW / 4]; uint8_t qb[VCW / 4]; };
// Score lookup: index = XOR of four packed 2-bit target codes with the four query codes they meet; value = the packed
// match/mismatch scores of cells (0,1) and (2,3).  256 entries per CTA, filled once at kernel start.
__device__ __forceinline__ void vec_fill_stab(uint2 *stab, const Opt &o)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int s0 = 2 * FB + ((i & 3) ? -o.b : o.a), s1 = 2 * FB + (((i >> 2) & 3) ? -o.b : o.a), s2 = 2 * FB + (((i >> 4) & 3) ? -o.b : o.a), s3 = 2 * FB + (((i >> 6) & 3) ? -o.b : o.a);
        stab[i] = make_uint2(((uint32_t)s0 & 0xffffu) | ((uint32_t)s1 << 16), ((uint32_t)s2 & 0xffffu) | ((uint32_t)s3 << 16));
    }
    __syncthreads();
}

__host__ __device__ __forceinline__ int vec_ncol(int qlen, int tlen, int w_in)
{
    int w = w_in < 0 ? (tlen > qlen ? tlen : qlen) : w_in;
    int ncol = qlen < tlen ? qlen : tlen;
    if (ncol > w + 1) ncol = w + 1;
    return ncol;
}
// one row of direction bytes: the 4-column groups that overlap [st, en]
__host__ __device__ __forceinline__ int vec_stride(int qlen, int tlen, int w_in) { return (vec_ncol(qlen, tlen, w_in) + 11) & ~3; }
__host__ __device__ __forceinline__ int64_t vec_dir_bytes(int qlen, int tlen, int w_in)
{
    if (qlen <= 0 || tlen <= 0) return 0;
    return (int64_t)(qlen + tlen - 1) * vec_stride(qlen, tlen, w_in);
}
// shape test of the vectorised path (ambiguous bases are detected while staging)
__host__ __device__ __forceinline__ bool vec_ok(const Opt &o, int qlen, int tlen, int w_in)
{
    if (qlen <= 0 || tlen <= 0 || vec_ncol(qlen, tlen, w_in) > VEC_MAX_NCOL) return false;
    // |H| must stay below 2^20 for the packed (score, rank) keys
    const int qm = o.q + o.e > o.q2 + o.e2 ? o.q + o.e : o.q2 + o.e2;
    if (o.a < 0 || o.b < 0 || o.a + o.b + qm > FB) return false;      // offset form: every stored byte stays in [0, 2 FB + a]
    const long long mab = o.a > o.b ? o.a : o.b, me = o.e > o.e2 ? o.e : o.e2;
    return mab * (qlen < tlen ? qlen : tlen) + o.q + o.q2 + me * ((long long)qlen + tlen) < (1 << 20) - 64;
}

__device__ bool warp_extd2_vec(const Opt &o, const DpTask &T, DpRes &R, VecSmem &M, const uint2 *stab, uint8_t *p, unsigned long long *cells_acc)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int qlen = T.qlen, tlen = T.tlen, flag = T.flag;
    const bool RIGHT = flag & KSW_RIGHT;
    int q = o.q, e = o.e, q2 = o.q2, e2 = o.e2;
    if (q2 + e2 < q + e) { int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }
    const int qe = q + e, qe2 = q2 + e2;
    const int w = T.w < 0 ? (tlen > qlen ? tlen : qlen) : T.w;
    const int vstride = vec_stride(qlen, tlen, T.w);
    int LT = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
    if (q2 + e2 + LT * e2 > q + e + LT * e) ++LT;
    const int LD = LT * (e - e2) - (q2 - q) - e2;
    const bool approx = flag & KSW_APPROX_MAX;
    const uint32_t QC1 = pk1(q - 1 + (RIGHT ? 1 : 0)), QC2 = pk1(q2 - 1 + (RIGHT ? 1 : 0)), Q1 = pk1(q), Q2 = pk1(q2);
    const uint32_t CA = (uint32_t)(FB - qe) * 65537u, CA2 = (uint32_t)(FB - qe2) * 65537u, DK = 0x80008000u;
    const uint32_t ONE = opaque_one(), MONE = 0u - ONE, TWO = ONE + ONE, FOUR = TWO + TWO, SIXTEEN = FOUR * FOUR, K01 = 0x01010101u * ONE;
    uint8_t *u = reinterpret_cast<uint8_t *>(M.st[0]), *v = reinterpret_cast<uint8_t *>(M.st[1]), *x = reinterpret_cast<uint8_t *>(M.st[2]);
    uint8_t *y = reinterpret_cast<uint8_t *>(M.st[3]), *x2 = reinterpret_cast<uint8_t *>(M.st[4]), *y2 = reinterpret_cast<uint8_t *>(M.st[5]);
    int32_t *H = M.H;
    // seed roles: lanes 0..5 write one array each when a column enters the band; lanes 0..2 write the left neighbour of st
    uint8_t *const enter_arr = reinterpret_cast<uint8_t *>(M.st[lane < 6 ? lane : 0]);
    const int enter_val = lane < 2 ? FB - qe : 0;                // u, v = -q-e; x, y = -q-e and x2, y2 = -q2-e2 are 0 in the offset form
    uint8_t *const left_arr = lane == 0 ? x : lane == 1 ? x2 : v;
    int pst = -1, pen = -1, t_loaded = 0, q_loaded = 0;
    int32_t ez_max = 0, ez_max_t = -1, ez_max_q = -1, ez_mqe = KSW_NEG_INF, ez_mqe_t = -1, ez_mte = KSW_NEG_INF, ez_mte_q = -1;
    int32_t ez_score = KSW_NEG_INF, zdropped = 0, H0 = 0, last_H0_t = 0;
    unsigned long long cells = 0;
    const int nr = qlen + tlen - 1;
    bool bail = false;
    const bool bounded = (flag & KSW_EXTZ_ONLY) && !approx && T.end_bonus <= 0;
    const int mlen = tlen < qlen ? tlen : qlen, r_over = 2 * (mlen - 1);      // beyond r_over one sequence is exhausted
    for (int r = 0; r < nr; ++r) {
        // bounded extension: no cell of this or a later anti-diagonal can beat the maximum found so far (ext_bound_stop)
        if (bounded && r > r_over && ext_bound_stop(o.a, q, e, q2, e2, qlen, tlen, r, ez_max)) { zdropped = 1; break; }
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
        if (en > (r + w) >> 1) en = (r + w) >> 1;
        if (st > en) { zdropped = 1; break; }
        cells += (unsigned long long)(en - st + 1);
        // ---- stage sequence codes, 128 at a time: byte = four 2-bit codes ----
        while (t_loaded <= en) {
            uint32_t pk = 0; bool amb = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = t_loaded + 4 * lane + k, c = i < tlen ? dp_base(T.t, T.tstep, 0, i) : 0;
                amb |= c > 3; pk |= (uint32_t)(c & 3) << (2 * k);
            }
            if (__any_sync(FULL, amb)) { bail = true; break; }
            M.tb[((t_loaded >> 2) + lane) & (VCW / 4 - 1)] = (uint8_t)pk;
            t_loaded += 128;
        }
        while (!bail && q_loaded <= r - st) {      // the query runs backwards along an anti-diagonal: row i sits at stream position 3 - i
            uint32_t pk = 0; bool amb = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = q_loaded + 4 * lane + k, c = i < qlen ? dp_base(T.q, T.qstep, T.qcomp, i) : 0;
                amb |= c > 3; pk |= (uint32_t)(c & 3) << (2 * (3 - k));
            }
            if (__any_sync(FULL, amb)) { bail = true; break; }
            M.qb[(-((q_loaded >> 2) + lane)) & (VCW / 4 - 1)] = (uint8_t)pk;
            q_loaded += 128;
        }
        if (bail) break;
        // ---- seeds: entering column, top boundary, left neighbour of the first cell ----
        const int bnd = r == 0 ? -q - e : r < LT ? -e : r == LT ? LD : -e2;
        if (en > pen && lane < 6) enter_arr[vwr(en)] = (uint8_t)enter_val;          // u,v,x,y = -q-e; x2,y2 = -q2-e2
        if (en == r && lane == 0) u[vwr(r)] = (uint8_t)(bnd + FB);                    // (en == r implies en > pen: y, y2 are already seeded)
        if ((st == 0 || !(st - 1 >= pst && st - 1 <= pen)) && lane < 3)
            left_arr[vwr(st - 1)] = (uint8_t)(lane < 2 ? 0 : FB + (st == 0 ? bnd : -qe));
        int32_t Hp = 0;
        if (!approx && r > 0) Hp = H[vwr(en > 0 ? en - 1 : 0)];                       // last row's H next to (or at) the entering end
        __syncwarp();
        const int gs = st >> 2, ge = en >> 2, nchunk = ((ge - gs) >> 5) + 1;
        uint8_t *pr = p + (int64_t)r * vstride - (gs << 2);
        const int en1 = st + ((en - st) >> 2 << 2);                                    // [st, en1): the reference's 4-wide groups; [en1, en): its scalar tail
        const unsigned n_upd = (unsigned)(en - st), n_trk = (unsigned)(en1 - st);
        int32_t kmax = INT32_MIN;
        int kc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) kc[k] = 2046 - ((k - st) & 3) * 256 - ((k - st) >> 2);
        for (int cb = nchunk - 1; cb >= 0; --cb) {
            const int g = gs + (cb << 5) + lane;
            const bool act = g <= ge;
            const int c0 = g << 2, gw = vwr(c0);
            uint32_t Wu = 0, Wv = 0, Wx = 0, Wy = 0, Wx2 = 0, Wy2 = 0, Lv4 = 0, Lx4 = 0, Lx24 = 0, X = 0;
            if (act) {
                Wu = *reinterpret_cast<const uint32_t *>(u + gw); Wy = *reinterpret_cast<const uint32_t *>(y + gw); Wy2 = *reinterpret_cast<const uint32_t *>(y2 + gw);
                Wv = *reinterpret_cast<const uint32_t *>(v + gw); Wx = *reinterpret_cast<const uint32_t *>(x + gw); Wx2 = *reinterpret_cast<const uint32_t *>(x2 + gw);
                const int gl = vwr(c0 - 4);
                Lv4 = __byte_perm(*reinterpret_cast<const uint32_t *>(v + gl), Wv, 0x6543);      // the same arrays, one column to the left
                Lx4 = __byte_perm(*reinterpret_cast<const uint32_t *>(x + gl), Wx, 0x6543);
                Lx24 = __byte_perm(*reinterpret_cast<const uint32_t *>(x2 + gl), Wx2, 0x6543);
                const int p0 = (c0 + 3 - r) & (VCW - 1), qi = p0 >> 2;      // stream position of the query row that meets the group's first column
                const uint32_t qh = (uint32_t)M.qb[qi] | (uint32_t)M.qb[(qi + 1) & (VCW / 4 - 1)] << 8;
                X = ((uint32_t)M.tb[g & (VCW / 4 - 1)] ^ (qh >> (2 * (p0 & 3)))) & 0xffu;
            }
            __syncwarp();
            if (act) {
                const uint2 SS = stab[X];
                uint32_t nU[2], nV[2], nX[2], nY[2], nX2[2], nY2[2], fO[2], fA[2], fB[2], fA2[2], fX[2], fY[2], fX2[2], fY2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t sxl = h ? 0x4342 : 0x4140;          // bytes (0, 1) or (2, 3), zero-extended to the two halves
                    const uint32_t up_u = prmt(Wu, 0u, sxl), up_y = prmt(Wy, 0u, sxl), up_y2 = prmt(Wy2, 0u, sxl);
                    const uint32_t Lv = prmt(Lv4, 0u, sxl), Lx = prmt(Lx4, 0u, sxl), Lx2 = prmt(Lx24, 0u, sxl);
                    const uint32_t S = h ? SS.y : SS.x;
                    const uint32_t A = Lx + Lv + CA, A2 = Lx2 + Lv + CA2, B = up_y + up_u + CA, B2 = up_y2 + up_u + CA2;
                    uint32_t Z = __vimax3_s16x2(S, A, B);
                    Z = __vimax3_s16x2(Z, A2, B2);
                    const uint32_t NZ = FSUB(DK, Z);                   // t + NZ = t - z + 0x8000 per half: top bit set when t is the maximum
                    const uint32_t DA = FADD(A, NZ), DB = FADD(B, NZ), DA2 = FADD(A2, NZ), DB2 = FADD(B2, NZ);
                    fO[h] = FADD(RIGHT ? B2 : S, NZ);          // the one candidate the two tie-break orders do not share
                    nU[h] = FSUB(Z, Lv); nV[h] = FSUB(Z, up_u);
                    // 0x8000 + max(t - z + q, 0): the low byte is the new gap state in the offset form
                    nX[h] = __viaddmax_u16x2(DA, Q1, DK); nY[h] = __viaddmax_u16x2(DB, Q1, DK);
                    nX2[h] = __viaddmax_u16x2(DA2, Q2, DK); nY2[h] = __viaddmax_u16x2(DB2, Q2, DK);
                    fA[h] = DA; fB[h] = DB; fA2[h] = DA2;
                    fX[h] = FADD(DA, QC1); fY[h] = FADD(DB, QC1); fX2[h] = FADD(DA2, QC2); fY2[h] = FADD(DB2, QC2);   // top bit set when the gap continues
                }
                // 8 top bits per cell, complemented: one PRMT per flag gathers the four cells of the group (bytes = columns c0..c0+3),
                // a tree of multiply-adds and one multiply put flag f on bit f (k_fill.cuh)
                uint32_t dw = flag_sum(prmt(fO[0], fO[1], 0xFDB9), prmt(fA[0], fA[1], 0xFDB9), prmt(fB[0], fB[1], 0xFDB9), prmt(fA2[0], fA2[1], 0xFDB9),
                                       prmt(fX[0], fX[1], 0xFDB9), prmt(fY[0], fY[1], 0xFDB9), prmt(fX2[0], fX2[1], 0xFDB9), prmt(fY2[0], fY2[1], 0xFDB9),
                                       TWO, FOUR, SIXTEEN);
                dw = imad(dw, K01, 0xffffffffu);
                // Whole words are written back: cells outside [st, en] hold garbage, which is never read (a column is
                // re-seeded when it enters the band, and the left neighbour of st is either last row's cell or a constant).
#define PACK8(a) __byte_perm((a)[0], (a)[1], 0x6420)
                *reinterpret_cast<uint32_t *>(u + gw) = PACK8(nU);
                *reinterpret_cast<uint32_t *>(v + gw) = PACK8(nV);
                *reinterpret_cast<uint32_t *>(x + gw) = PACK8(nX);
                *reinterpret_cast<uint32_t *>(y + gw) = PACK8(nY);
                *reinterpret_cast<uint32_t *>(x2 + gw) = PACK8(nX2);
                *reinterpret_cast<uint32_t *>(y2 + gw) = PACK8(nY2);
#undef PACK8
                *reinterpret_cast<uint32_t *>(pr + c0) = dw;
                if (!approx && r > 0) {
                    // H[t] += v[t] for t in [st, en); one integer key per cell orders (H, then the reference's scan order):
                    // (t - st) = 4 * idx + cls, and among equal maxima the reference keeps the smallest (cls, idx)
                    const int4 h4 = *reinterpret_cast<const int4 *>(H + gw);
                    int32_t hv[4] = {h4.x, h4.y, h4.z, h4.w};
                    const int d = c0 - st;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int32_t vk = (int32_t)((k & 1) ? nV[k >> 1] >> 16 : nV[k >> 1] & 0xffffu) - FB;
                        const int rel = d + k;
                        if ((unsigned)rel < n_upd) hv[k] += vk;
                        // rank code 2046 - cls * 256 - idx with cls = (k - st) & 3 and idx = g + ((k - st) >> 2): kc[k] is uniform over the row
                        const int32_t key = hv[k] * 2048 + (kc[k] - g);
                        if ((unsigned)rel < n_trk && key > kmax) kmax = key;
                    }
                    *reinterpret_cast<int4 *>(H + gw) = make_int4(hv[0], hv[1], hv[2], hv[3]);
                }
            }
            __syncwarp();
        }
        if (!approx) {
            int32_t max_H, max_t, Hen, Hst;
            if (r > 0) {
                Hen = Hp + (en > 0 ? (int32_t)u[vwr(en)] : (int32_t)v[vwr(0)]) - FB;
                if (lane < 3 && en1 + lane < en) {                    // the scalar tail of the reference's row scan
                    const int32_t key = H[vwr(en1 + lane)] * 2048 + (2046 - 1024 - lane);
                    if (key > kmax) kmax = key;
                }
                {   // the entering end is compared last with >=: it wins every tie
                    const int32_t key = Hen * 2048 + 2047;
                    if (key > kmax) kmax = key;
                }
                const int32_t kall = __reduce_max_sync(FULL, kmax);
                max_H = kall >> 11;
                const int rk = 2047 - (kall & 2047);
                if (rk == 0) max_t = en;
                else { const int cls = (rk - 1) >> 8, idx = (rk - 1) & 255; max_t = cls < 4 ? st + 4 * idx + cls : en1 + idx; }
                Hst = st == en ? Hen : H[vwr(st)];
                if (lane == 0) H[vwr(en)] = Hen;
            } else {
                Hen = Hst = (int32_t)v[0] - FB - qe;
                if (lane == 0) H[0] = Hen;
                max_H = Hen, max_t = 0;
            }
            __syncwarp();           // H[en] is read by every lane at the top of the next row
            if (en == tlen - 1 && Hen > ez_mte) ez_mte = Hen, ez_mte_q = r - en;
            if (r - st == qlen - 1 && Hst > ez_mqe) ez_mqe = Hst, ez_mqe_t = st;
            bool stop = false;
            if (max_H > ez_max) ez_max = max_H, ez_max_t = max_t, ez_max_q = r - max_t;
            else if (max_t >= ez_max_t && r - max_t >= ez_max_q) {
                int tl = max_t - ez_max_t, ql = (r - max_t) - ez_max_q, l = tl > ql ? tl - ql : ql - tl;
                if (T.zdrop >= 0 && ez_max - max_H > T.zdrop + l * e2) zdropped = 1, stop = true;
            }
            if (stop) break;
            if (r == nr - 1 && en == tlen - 1) ez_score = Hen;
        } else {
            if (r > 0) {
                if (last_H0_t >= st && last_H0_t <= en && last_H0_t + 1 >= st && last_H0_t + 1 <= en) {
                    int d0 = (int)v[vwr(last_H0_t)] - FB, d1 = (int)u[vwr(last_H0_t + 1)] - FB;
                    if (d0 > d1) H0 += d0; else H0 += d1, ++last_H0_t;
                } else if (last_H0_t >= st && last_H0_t <= en) H0 += (int)v[vwr(last_H0_t)] - FB;
                else ++last_H0_t, H0 += (int)u[vwr(last_H0_t)] - FB;
            } else H0 = (int32_t)v[0] - FB - qe, last_H0_t = 0;
            if (r == nr - 1 && en == tlen - 1) ez_score = H0;
        }
        pst = st, pen = en;
    }
    if (bail) return false;
    if (lane == 0) {
        atomicAdd(cells_acc, cells);
        res_reset(R);
        R.max = ez_max; R.max_t = ez_max_t; R.max_q = ez_max_q; R.mqe = ez_mqe; R.mqe_t = ez_mqe_t;
        R.mte = ez_mte; R.mte_q = ez_mte_q; R.score = ez_score; R.zdropped = zdropped;
    }
    __syncwarp();
    return true;
}


// ksw_backtrack over the sign-bit direction bytes of warp_extd2_vec; sequential, one thread
__device__ void extd2_traceback_vec(const DpTask &T, DpRes &R, const uint8_t *p, uint32_t *ezcig, int ezcap, int32_t *err)
{
    const int qlen = T.qlen, tlen = T.tlen, flag = T.flag;
    const bool right = flag & KSW_RIGHT;
    R.n_cigar = 0; R.cigar = ezcig; R.reach_end = 0;
    if (qlen <= 0 || tlen <= 0) return;
    const int w = T.w < 0 ? (tlen > qlen ? tlen : qlen) : T.w;
    const int vstride = vec_stride(qlen, tlen, T.w);
    int i0 = -1, j0 = -1;
    if (!R.zdropped && !(flag & KSW_EXTZ_ONLY)) i0 = tlen - 1, j0 = qlen - 1;
    else if (!R.zdropped && (flag & KSW_EXTZ_ONLY) && R.mqe + T.end_bonus > R.max) R.reach_end = 1, i0 = R.mqe_t, j0 = qlen - 1;
    else if (R.max_t >= 0 && R.max_q >= 0) i0 = R.max_t, j0 = R.max_q;
    if (i0 < 0 || j0 < 0) return;
    uint32_t *c = ezcig; int n = 0;
    int run_op = -1; uint32_t run_len = 0;          // the open CIGAR run lives in registers
#define FLUSH() do { if (run_len) { if (n < ezcap) c[n] = run_len << 4 | (uint32_t)run_op; ++n; } } while (0)
#define PUSH(op, len) do { if ((op) == run_op) run_len += (uint32_t)(len); else { FLUSH(); run_op = (op); run_len = (uint32_t)(len); } } while (0)
    int i = i0, j = j0, state = 0;
    while (i >= 0 && j >= 0) {
        int r = i + j, force_state = -1;
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
        if (en > (r + w) >> 1) en = (r + w) >> 1;
        if (i < st) force_state = 2;
        if (i > en) force_state = 1;
        int d = 0; uint32_t b = 0;
        if (force_state < 0) {
            b = p[(int64_t)r * vstride + (i - (st & ~3))];
            if (!right) d = !(b & 1) ? 0 : !(b & 2) ? 1 : !(b & 4) ? 2 : !(b & 8) ? 3 : 4;
            else d = !(b & 1) ? 4 : !(b & 8) ? 3 : !(b & 4) ? 2 : !(b & 2) ? 1 : 0;
        }
        if (state == 0) state = d;
        else if (force_state >= 0 || ((b >> (3 + state)) & 1)) state = 0;
        if (state == 0) state = d;
        if (force_state >= 0) state = force_state;
        if (state == 0) { PUSH(0, 1); --i; --j; }
        else if (state == 1 || state == 3) { PUSH(2, 1); --i; }
        else { PUSH(1, 1); --j; }
    }
    if (i >= 0) PUSH(2, i + 1);
    if (j >= 0) PUSH(1, j + 1);
    FLUSH();
#undef PUSH
#undef FLUSH
    if (n > ezcap) { atomicOr(err, TELR_ERR_CIGCAP); n = 0; }
    if (!(flag & KSW_REV_CIGAR))
        for (int k = 0; k < n >> 1; ++k) { uint32_t t = c[k]; c[k] = c[n - 1 - k]; c[n - 1 - k] = t; }
    R.n_cigar = n;
}

}  // namespace telr
