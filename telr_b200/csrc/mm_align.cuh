// mm_align.cuh — base-level alignment driver for one read x contig strand as a resumable state machine.
//
// minimap2's mm_align_skeleton/mm_align1 (align.c) interleave sequential bookkeeping with DP calls.
// On the GPU the DP is a warp-wide primitive, so the bookkeeping is written as a coroutine:
// aln_next() runs on one lane, consumes the result of the previous DP request and either emits the
// next request (left extension, gap fill, exact re-run, inversion probe, right extension) or reports
// completion.  tests/emu drives the same function on the host.
#pragma once
#include "mm_chain.cuh"
#include "mm_hit.cuh"

namespace telr {

enum AlnPhase {
    PH_START = 0, PH_REG_BEGIN, PH_LEFT_DONE, PH_FILL_NEXT, PH_FILL1_DONE, PH_FILL_PROBE_DONE, PH_FILL_DECIDE,
    PH_FILL2_DONE, PH_FILL_CONSUME, PH_RIGHT, PH_RIGHT_DONE, PH_REG_FINAL, PH_INV_LL_DONE, PH_INV_EXT_DONE,
    PH_NEXT_REG, PH_FINISH, PH_DONE
};

struct AlnCtx {
    const Opt *o;
    const uint8_t *tseq; int tlen;
    const uint8_t *qseq[2]; int qlen;
    Anchor *a; int n_a;
    Reg *regs; int n_regs, cap_regs;
    uint32_t *cig; uint32_t cig_top, cig_cap;
    HitScratch hs;
    int32_t *K; int capK;
    int err, rep_len;       // rep_len: query bases under filtered high-occurrence seeds (feeds MAPQ)
    int defer_finish;       // 1: leave reg_finish + the final region pass to aln_finish() (GPU rounds), 0: inline
    int64_t n_tasks;
    // ---- coroutine state ----
    int phase, ireg;
    int as1, cnt1, rev, bw, bw_long, bw1, dropped, zdrop_code, i;
    int rs, re, qs, qe, rs0, re0, qs0, qe0, rs1, qs1, re1, qe1;
    int max_zdrop, zpos[2][2];
    Reg r2;
    DpRes res;          // result being consumed (first pass kept across the probe)
    // inversion attempt
    int inv_ql, inv_tl, inv_qoff, inv_toff;
    const uint8_t *inv_q;
};

enum { TELR_ERR_REGCAP = 1, TELR_ERR_CIGCAP = 2, TELR_ERR_KCAP = 4, TELR_ERR_DIRCAP = 8 };

TELR_HD int sc_pair(const Opt &o, int tc, int qc) { return (tc > 3 || qc > 3) ? -o.sc_ambi : tc == qc ? o.a : -o.b; }

TELR_HD void res_reset(DpRes &r)
{
    r.max_q = r.max_t = r.mqe_t = r.mte_q = -1;
    r.max = 0; r.score = r.mqe = r.mte = KSW_NEG_INF;
    r.n_cigar = 0; r.zdropped = 0; r.reach_end = 0;
    r.ll_score = 0; r.ll_qe = r.ll_te = -1;
}

// append DP cigar to the region under construction (its cigar is the top of the arena)
TELR_HD void reg_append_cigar(AlnCtx &c, Reg &r, int n, const uint32_t *cg)
{
    if (n == 0) return;
    if (!r.has_p) { r.has_p = 1; r.dp_score = r.dp_max = r.dp_max2 = r.n_ambi = 0; r.n_cigar = 0; r.cig = c.cig_top; }
    if (c.cig_top + (uint32_t)n > c.cig_cap) { c.err |= TELR_ERR_CIGCAP; return; }
    uint32_t *dst = c.cig + r.cig;
    if (r.n_cigar > 0 && (dst[r.n_cigar - 1] & 0xf) == (cg[0] & 0xf)) {
        dst[r.n_cigar - 1] += (cg[0] >> 4) << 4;
        for (int k = 1; k < n; ++k) dst[r.n_cigar + k - 1] = cg[k];
        r.n_cigar += n - 1;
    } else {
        for (int k = 0; k < n; ++k) dst[r.n_cigar + k] = cg[k];
        r.n_cigar += n;
    }
    c.cig_top = r.cig + (uint32_t)r.n_cigar;
}

TELR_HDN void fix_bad_ends(const Reg &r, const Anchor *a, int bw, int min_match, int *as, int *cnt)
{
    *as = r.as, *cnt = r.cnt;
    if (r.cnt < 3) return;
    int m, l;
    m = l = (int)(a[r.as].y >> 32 & 0xff);
    for (int i = r.as + 1; i < r.as + r.cnt - 1; ++i) {
        int span = (int)(a[i].y >> 32 & 0xff);
        if (a[i].y & SEED_LONG_JOIN) break;
        int lr = (int32_t)a[i].x - (int32_t)a[i - 1].x, lq = (int32_t)a[i].y - (int32_t)a[i - 1].y;
        int mn = lr < lq ? lr : lq, mx = lr > lq ? lr : lq;
        if (mx - mn > l >> 1) *as = i;
        l += mn;
        m += mn < span ? mn : span;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r.mlen >> 1) break;
    }
    *cnt = r.as + r.cnt - *as;
    m = l = (int)(a[r.as + r.cnt - 1].y >> 32 & 0xff);
    for (int i = r.as + r.cnt - 2; i > *as; --i) {
        int span = (int)(a[i + 1].y >> 32 & 0xff);
        if (a[i + 1].y & SEED_LONG_JOIN) break;
        int lr = (int32_t)a[i + 1].x - (int32_t)a[i].x, lq = (int32_t)a[i + 1].y - (int32_t)a[i].y;
        int mn = lr < lq ? lr : lq, mx = lr > lq ? lr : lq;
        if (mx - mn > l >> 1) *cnt = i + 1 - *as;
        l += mn;
        m += mn < span ? mn : span;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r.mlen >> 1) break;
    }
}

TELR_HD int anchor_gap(const Anchor *a, int i) // (dq - dr) between anchors i-1 and i
{
    return ((int32_t)a[i].y - (int32_t)a[i - 1].y) - ((int32_t)a[i].x - (int32_t)a[i - 1].x);
}

TELR_HD int long_gaps(AlnCtx &c, int as1, int cnt1, int min_gap)
{
    int n = 0;
    const Anchor *a = c.a + as1;
    for (int i = 1; i < cnt1; ++i) {
        int gap = anchor_gap(a, i);
        if (gap < -min_gap || gap > min_gap) {
            if (n < c.capK) c.K[n] = i; else c.err |= TELR_ERR_KCAP;
            ++n;
        }
    }
    if (n > c.capK) n = c.capK;
    return n <= 1 ? 0 : n;
}

TELR_HDN void filter_bad_seeds(AlnCtx &c, int as1, int cnt1, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt)
{
    int n = long_gaps(c, as1, cnt1, min_gap);
    if (n == 0) return;
    Anchor *a = c.a + as1;
    const int32_t *K = c.K;
    int mx = 0, max_st = -1, max_en = -1;
    for (int k = 0;; ++k) {
        int n_ins = 0, n_del = 0, max_diff = 0, max_diff_l = -1;
        if (k == n || k >= max_en) {
            if (max_en > 0)
                for (int i = K[max_st]; i < K[max_en]; ++i) a[i].y |= SEED_IGNORE;
            mx = 0, max_st = max_en = -1;
            if (k == n) break;
        }
        int i = K[k];
        int gap = anchor_gap(a, i);
        if (gap > 0) n_ins += gap; else n_del += -gap;
        int qs = (int32_t)a[i - 1].y, rs = (int32_t)a[i - 1].x;
        for (int l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
            int j = K[l];
            if ((int32_t)a[j].y - qs > max_ext_len || (int32_t)a[j].x - rs > max_ext_len) break;
            gap = anchor_gap(a, j);
            if (gap > 0) n_ins += gap; else n_del += -gap;
            int ad = n_ins - n_del; if (ad < 0) ad = -ad;
            int diff = n_ins + n_del - ad;
            if (max_diff < diff) max_diff = diff, max_diff_l = l;
        }
        if (max_diff > diff_thres && max_diff > mx) mx = max_diff, max_st = k, max_en = max_diff_l;
    }
}

TELR_HDN void filter_bad_seeds_alt(AlnCtx &c, int as1, int cnt1, int min_gap, int max_ext)
{
    int n = long_gaps(c, as1, cnt1, min_gap);
    if (n == 0) return;
    Anchor *a = c.a + as1;
    const int32_t *K = c.K;
    for (int k = 0; k < n;) {
        int i = K[k], l;
        int gap1 = anchor_gap(a, i);
        int re1 = (int32_t)a[i].x, qe1 = (int32_t)a[i].y;
        gap1 = gap1 > 0 ? gap1 : -gap1;
        for (l = k + 1; l < n; ++l) {
            int j = K[l];
            if ((int32_t)a[j].y - qe1 > max_ext || (int32_t)a[j].x - re1 > max_ext) break;
            int gap2 = anchor_gap(a, j);
            int span_pre = (int)(a[j - 1].y >> 32 & 0xff);
            int rs2 = (int32_t)a[j - 1].x + span_pre, qs2 = (int32_t)a[j - 1].y + span_pre;
            int m = rs2 - re1 < qs2 - qe1 ? rs2 - re1 : qs2 - qe1;
            gap2 = gap2 > 0 ? gap2 : -gap2;
            if (m > gap1 + gap2) break;
            re1 = (int32_t)a[j].x, qe1 = (int32_t)a[j].y;
            gap1 = gap2;
        }
        if (l > k + 1) {
            int end = K[l - 1];
            for (int j = K[k]; j < end; ++j) a[j].y |= SEED_IGNORE;
            a[end].y |= SEED_LONG_JOIN;
        }
        k = l;
    }
}

// position inside a minimizer at which DP windows are cut
TELR_HD void anchor_cut(const AlnCtx &c, const Anchor &an, int *r, int *q)
{
    if (c.o->hpc) {
        const uint8_t *qs = c.qseq[an.x >> 63];
        int i, cc;
        *q = (int32_t)an.y;
        for (i = *q - 1, cc = qs[*q]; i > 0; --i)
            if (qs[i] != cc) break;
        *q = i + 1;
        *r = (int32_t)an.x;
        for (i = *r - 1, cc = c.tseq[*r]; i > 0; --i)
            if (c.tseq[i] != cc) break;
        *r = i + 1;
    } else {
        *r = (int32_t)an.x - (c.o->k >> 1);
        *q = (int32_t)an.y - (c.o->k >> 1);
    }
}

// rescan a gap-fill CIGAR with single-affine costs: largest score drop along the path and where
TELR_HDN void scan_zdrop(AlnCtx &c, const uint8_t *qseq, const uint8_t *tseq, int n_cigar, const uint32_t *cigar)
{
    const Opt &o = *c.o;
    int32_t score = 0, mx = INT32_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
    int (*pos)[2] = c.zpos;
    pos[0][0] = pos[0][1] = pos[1][0] = pos[1][1] = -1;
    for (int k = 0; k < n_cigar; ++k) {
        int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
        if (op == 0) {
            for (int l = 0; l < len; ++l) {
                score += sc_pair(o, tseq[i + l], qseq[j + l]);
                if (score < mx) {
                    int li = i + l - max_i, lj = j + l - max_j;
                    int diff = li > lj ? li - lj : lj - li;
                    int z = mx - score - diff * o.e;
                    if (z > max_zdrop) { max_zdrop = z; pos[0][0] = max_i, pos[0][1] = max_j; pos[1][0] = i + l, pos[1][1] = j + l; }
                } else mx = score, max_i = i + l, max_j = j + l;
            }
            i += len, j += len;
        } else if (op == 1 || op == 2) {
            score -= o.q + o.e * len;
            if (op == 1) j += len; else i += len;
            if (score < mx) {
                int li = i - max_i, lj = j - max_j;
                int diff = li > lj ? li - lj : lj - li;
                int z = mx - score - diff * o.e;
                if (z > max_zdrop) { max_zdrop = z; pos[0][0] = max_i, pos[0][1] = max_j; pos[1][0] = i, pos[1][1] = j; }
            } else mx = score, max_i = i, max_j = j;
        }
    }
    c.max_zdrop = max_zdrop;
}

// Upper bound on everything scan_zdrop could report, from the CIGAR and the fill's DP score alone (no sequence access).
// With D diagonal cells, m mismatches, n ambiguous pairs and gap runs costing G2 under the DP's two-piece model:
//   score = a(D - m - n) - b m - sc_ambi n - G2   =>   (a+b) m + (a+sc_ambi) n = a D - score - G2 <= a D - score - G2min
// and no drop along the path can exceed the single-affine loss  b m + sc_ambi n + G1 <= floor(b X / (a+b)) + G1
// (valid while b/(a+b) >= sc_ambi/(a+sc_ambi)).  Returns INT32_MAX when no bound is available.
TELR_HDN int32_t zdrop_bound(const Opt &o, int n_cigar, const uint32_t *cigar, int32_t score)
{
    if (n_cigar <= 0 || score <= KSW_NEG_INF / 2) return INT32_MAX;
    if ((long long)o.b * (o.a + o.sc_ambi) < (long long)o.sc_ambi * (o.a + o.b)) return INT32_MAX;
    long long D = 0, G1 = 0, G2 = 0;
    for (int k = 0; k < n_cigar; ++k) {
        const int op = cigar[k] & 0xf; const long long len = cigar[k] >> 4;
        if (op == 0) D += len;
        else {
            const long long c1 = o.q + o.e * len, c2 = o.q2 + o.e2 * len;
            G1 += c1; G2 += c1 < c2 ? c1 : c2;
        }
    }
    const long long X = o.a * D - score - G2;
    if (X < 0) return INT32_MAX;
    const long long N = o.b * X / (o.a + o.b) + G1;
    return N < INT32_MAX ? (int32_t)N : INT32_MAX;
}

// CIGAR clean-up after all pieces are joined: indel left-alignment, I/D merging, leading I/D removal
TELR_HDN void fix_cigar(AlnCtx &c, Reg &r, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift)
{
    uint32_t *cg = c.cig + r.cig;
    int32_t toff = 0, qoff = 0, to_shrink = 0;
    *qshift = *tshift = 0;
    if (r.n_cigar <= 1) return;
    for (int k = 0; k < r.n_cigar; ++k) {
        uint32_t op = cg[k] & 0xf, len = cg[k] >> 4;
        if (len == 0) to_shrink = 1;
        if (op == 0) {
            toff += len, qoff += len;
        } else if (op == 1 || op == 2) {
            if (k > 0 && k < r.n_cigar - 1 && (cg[k - 1] & 0xf) == 0 && (cg[k + 1] & 0xf) == 0) {
                int l, prev_len = (int)(cg[k - 1] >> 4);
                if (op == 1) {
                    for (l = 0; l < prev_len; ++l)
                        if (qseq[qoff - 1 - l] != qseq[qoff + len - 1 - l]) break;
                } else {
                    for (l = 0; l < prev_len; ++l)
                        if (tseq[toff - 1 - l] != tseq[toff + len - 1 - l]) break;
                }
                if (l > 0) cg[k - 1] -= (uint32_t)l << 4, cg[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
                if (l == prev_len) to_shrink = 1;
            }
            if (op == 1) qoff += len; else toff += len;
        }
    }
    for (int k = 0; k < r.n_cigar - 2; ++k) {
        if ((cg[k] & 0xf) > 0 && (cg[k] & 0xf) + (cg[k + 1] & 0xf) == 3) {
            int l;
            uint32_t s[3] = {0, 0, 0};
            for (l = k; l < r.n_cigar; ++l) {
                uint32_t op = cg[l] & 0xf;
                if (op == 1 || op == 2 || cg[l] >> 4 == 0) s[op] += cg[l] >> 4;
                else break;
            }
            if (s[1] > 0 && s[2] > 0 && l - k > 2) {
                cg[k] = s[1] << 4 | 1;
                cg[k + 1] = s[2] << 4 | 2;
                for (k += 2; k < l; ++k) cg[k] &= 0xf;
                to_shrink = 1;
            }
            k = l;
        }
    }
    if (to_shrink) {
        int l = 0;
        for (int k = 0; k < r.n_cigar; ++k)
            if (cg[k] >> 4 != 0) cg[l++] = cg[k];
        r.n_cigar = l;
        l = 0;
        for (int k = 0; k < r.n_cigar; ++k)
            if (k == r.n_cigar - 1 || (cg[k] & 0xf) != (cg[k + 1] & 0xf)) cg[l++] = cg[k];
            else cg[k + 1] += cg[k] >> 4 << 4;
        r.n_cigar = l;
    }
    if ((cg[0] & 0xf) == 1 || (cg[0] & 0xf) == 2) {
        int l = (int)(cg[0] >> 4);
        if ((cg[0] & 0xf) == 1) {
            if (r.rev) r.qe -= l; else r.qs += l;
            *qshift = l;
        } else r.rs += l, *tshift = l;
        --r.n_cigar;
        for (int k = 0; k < r.n_cigar; ++k) cg[k] = cg[k + 1];
    }
}

// final per-region statistics from the joined CIGAR (mlen, blen, dp_max)
TELR_HDN void reg_finish(AlnCtx &c, Reg &r, const uint8_t *qseq, const uint8_t *tseq)
{
    const Opt &o = *c.o;
    if (!r.has_p) return;
    int qshift, tshift, toff = 0, qoff = 0;
    double s = 0.0, mx = 0.0;
    fix_cigar(c, r, qseq, tseq, &qshift, &tshift);
    qseq += qshift, tseq += tshift;
    const uint32_t *cg = c.cig + r.cig;
    r.blen = r.mlen = 0;
    for (int k = 0; k < r.n_cigar; ++k) {
        int op = cg[k] & 0xf, len = (int)(cg[k] >> 4);
        if (op == 0) {
            int n_ambi = 0, diff = 0;
            for (int l = 0; l < len; ++l) {
                int cq = qseq[qoff + l], ct = tseq[toff + l];
                if (ct > 3 || cq > 3) ++n_ambi;
                else if (ct != cq) ++diff;
                s += sc_pair(o, ct, cq);
                if (s < 0) s = 0; else mx = mx > s ? mx : s;
            }
            r.blen += len - n_ambi, r.mlen += len - (n_ambi + diff), r.n_ambi += n_ambi;
            toff += len, qoff += len;
        } else if (op == 1) {
            int n_ambi = 0;
            for (int l = 0; l < len; ++l)
                if (qseq[qoff + l] > 3) ++n_ambi;
            r.blen += len - n_ambi, r.n_ambi += n_ambi;
            s -= o.q + o.e * len;
            if (s < 0) s = 0;
            qoff += len;
        } else if (op == 2) {
            int n_ambi = 0;
            for (int l = 0; l < len; ++l)
                if (tseq[toff + l] > 3) ++n_ambi;
            r.blen += len - n_ambi, r.n_ambi += n_ambi;
            s -= o.q + o.e * len;
            if (s < 0) s = 0;
            toff += len;
        }
    }
    r.dp_max = (int32_t)(mx + .499);
}

// re-rank near-equal full-length hits by an identity-aware score (align.c mm_update_dp_max)
TELR_HDN void regs_update_dp_max(AlnCtx &c)
{
    const Opt &o = *c.o;
    int n_regs = c.n_regs;
    Reg *regs = c.regs;
    int32_t mx = -1, mx2 = -1, max_i = -1;
    if (n_regs < 2) return;
    for (int i = 0; i < n_regs; ++i) {
        const Reg &r = regs[i];
        if (!r.has_p) continue;
        if (r.dp_max > mx) mx2 = mx, mx = r.dp_max, max_i = i;
        else if (r.dp_max > mx2) mx2 = r.dp_max;
    }
    if (max_i < 0 || mx < 0 || mx2 < 0) return;
    if (regs[max_i].qe - regs[max_i].qs < TELR_DMUL((double)c.qlen, (double)o.rank_frac)) return;
    if (mx2 < TELR_DMUL((double)mx, (double)o.rank_frac)) return;
    double div = TELR_DADD(1., -((double)regs[max_i].mlen / regs[max_i].blen));
    if (div < 0.02) div = 0.02;
    double b2 = 0.5 / div;
    if (TELR_DMUL(b2, (double)o.a) < o.b) b2 = (double)o.a / o.b;
    for (int i = 0; i < n_regs; ++i) {
        Reg &r = regs[i];
        if (!r.has_p) continue;
        const uint32_t *cg = c.cig + r.cig;
        int n_gap = 0;
        double gap_cost = 0.0;
        for (int k = 0; k < r.n_cigar; ++k) {
            int op = cg[k] & 0xf, len = (int)(cg[k] >> 4);
            if (op == 1 || op == 2) {
                gap_cost = TELR_DADD(gap_cost, TELR_DADD(b2, (double)fast_log2(TELR_FADD(1.0f, (float)len))));
                n_gap += len;
            }
        }
        int n_mis = r.blen + r.n_ambi - r.mlen - n_gap;
        double inner = TELR_DADD(TELR_DADD((double)r.mlen, -TELR_DMUL(b2, (double)n_mis)), -gap_cost);
        r.dp_max = (int32_t)TELR_DADD(TELR_DMUL((double)o.a, inner), .499);
        if (r.dp_max < 0) r.dp_max = 0;
    }
}

TELR_HD bool regs_insert(AlnCtx &c, const Reg &r, int i)
{
    if (c.n_regs >= c.cap_regs) { c.err |= TELR_ERR_REGCAP; return false; }
    for (int k = c.n_regs - 1; k > i; --k) c.regs[k + 1] = c.regs[k];
    c.regs[i + 1] = r;
    ++c.n_regs;
    return true;
}

TELR_HD void task_ext(DpTask &t, const uint8_t *q, int qstep, int qlen, const uint8_t *tt, int tstep, int tlen, int w,
                      int end_bonus, int zdrop, int flag)
{
    t.kind = 0; t.q = q; t.t = tt; t.qstep = qstep; t.tstep = tstep; t.qcomp = 0;
    t.qlen = qlen; t.tlen = tlen; t.w = w; t.zdrop = zdrop; t.end_bonus = end_bonus; t.flag = flag;
}

// returns true and fills `task` when a DP request is pending; false when the problem is finished.
// `in` is the result of the previously emitted request (ignored on the first call).
TELR_HDN bool aln_next(AlnCtx &c, const DpRes &in, DpTask &task)
{
    const Opt &o = *c.o;
    for (;;) {
        switch (c.phase) {
        case PH_START: {
            c.n_a = regs_squeeze_anchors(c.n_regs, c.regs, c.a, c.hs);
            c.ireg = 0;
            c.cig_top = 0;
            c.phase = PH_REG_BEGIN;
            break;
        }
        case PH_REG_BEGIN: {
            if (c.ireg >= c.n_regs) { c.phase = PH_FINISH; break; }
            Reg &r = c.regs[c.ireg];
            const Anchor *a = c.a;
            c.r2.cnt = 0;
            if (r.cnt == 0) { c.phase = PH_NEXT_REG; break; }
            c.rev = (int)(a[r.as].x >> 63);
            c.dropped = 0;
            c.bw = (int)(o.bw * 1.5 + 1.);
            c.bw_long = (int)(o.bw_long * 1.5 + 1.);
            if (c.bw_long < c.bw) c.bw_long = c.bw;
            fix_bad_ends(r, a, o.bw, o.min_chain_score * 2, &c.as1, &c.cnt1);
            filter_bad_seeds(c, c.as1, c.cnt1, 10, 40, o.max_gap >> 1, 10);
            filter_bad_seeds_alt(c, c.as1, c.cnt1, 30, o.max_gap >> 1);
            anchor_cut(c, a[c.as1], &c.rs, &c.qs);
            anchor_cut(c, a[c.as1 + c.cnt1 - 1], &c.re, &c.qe);
            int rs = c.rs, qs = c.qs, re = c.re, qe = c.qe, l, i;
            int rs0, qs0, rs1, qs1, re0, qe0, re1, qe1;
            rs0 = (int32_t)a[r.as].x + 1 - (int32_t)(a[r.as].y >> 32 & 0xff);
            qs0 = (int32_t)a[r.as].y + 1 - (int32_t)(a[r.as].y >> 32 & 0xff);
            if (rs0 < 0) rs0 = 0;
            rs1 = qs1 = 0;
            for (i = r.as - 1, l = 0; i >= 0 && a[i].x >> 32 == a[r.as].x >> 32; --i) {
                int x = (int32_t)a[i].x + 1 - (int32_t)(a[i].y >> 32 & 0xff);
                int y = (int32_t)a[i].y + 1 - (int32_t)(a[i].y >> 32 & 0xff);
                if (x < rs0 && y < qs0) {
                    if (++l > o.min_cnt) {
                        l = rs0 - x > qs0 - y ? rs0 - x : qs0 - y;
                        rs1 = rs0 - l, qs1 = qs0 - l;
                        if (rs1 < 0) rs1 = 0;
                        break;
                    }
                }
            }
            if (qs > 0 && rs > 0) {
                l = qs < o.max_gap ? qs : o.max_gap;
                qs1 = qs1 > qs - l ? qs1 : qs - l;
                qs0 = qs0 < qs1 ? qs0 : qs1;
                l += l * o.a > o.q ? (l * o.a - o.q) / o.e : 0;
                l = l < o.max_gap ? l : o.max_gap;
                l = l < rs ? l : rs;
                rs1 = rs1 > rs - l ? rs1 : rs - l;
                rs0 = rs0 < rs1 ? rs0 : rs1;
                rs0 = rs0 < rs ? rs0 : rs;
            } else rs0 = rs, qs0 = qs;
            re0 = (int32_t)a[r.as + r.cnt - 1].x + 1;
            qe0 = (int32_t)a[r.as + r.cnt - 1].y + 1;
            re1 = c.tlen, qe1 = c.qlen;
            for (i = r.as + r.cnt, l = 0; i < c.n_a && a[i].x >> 32 == a[r.as].x >> 32; ++i) {
                int x = (int32_t)a[i].x + 1, y = (int32_t)a[i].y + 1;
                if (x > re0 && y > qe0) {
                    if (++l > o.min_cnt) {
                        l = x - re0 > y - qe0 ? x - re0 : y - qe0;
                        re1 = re0 + l, qe1 = qe0 + l;
                        break;
                    }
                }
            }
            if (qe < c.qlen && re < c.tlen) {
                l = c.qlen - qe < o.max_gap ? c.qlen - qe : o.max_gap;
                qe1 = qe1 < qe + l ? qe1 : qe + l;
                qe0 = qe0 > qe1 ? qe0 : qe1;
                l += l * o.a > o.q ? (l * o.a - o.q) / o.e : 0;
                l = l < o.max_gap ? l : o.max_gap;
                l = l < c.tlen - re ? l : c.tlen - re;
                re1 = re1 < re + l ? re1 : re + l;
                re0 = re0 > re1 ? re0 : re1;
            } else re0 = re, qe0 = qe;
            c.rs0 = rs0, c.qs0 = qs0, c.re0 = re0, c.qe0 = qe0;
            c.rs1 = rs1, c.qs1 = qs1, c.re1 = re1, c.qe1 = qe1;
            r.cig = c.cig_top;   // this region's CIGAR grows at the top of the arena
            if (qs > 0 && rs > 0) {
                int ql = qs - qs0, tl = rs - rs0;
                c.phase = PH_LEFT_DONE;
                if (o.max_sw_mat > 0 && (long long)tl * ql > o.max_sw_mat) { res_reset(c.res); c.res.zdropped = 1; break; }
                if (ql <= 0 || tl <= 0) { res_reset(c.res); break; }
                task_ext(task, &c.qseq[c.rev][qs - 1], -1, ql, &c.tseq[rs - 1], -1, tl, c.bw, o.end_bonus,
                         r.split_inv ? o.zdrop_inv : o.zdrop, KSW_EXTZ_ONLY | KSW_RIGHT | KSW_REV_CIGAR);
                ++c.n_tasks;
                c.phase = PH_LEFT_DONE + 100;    // +100: take the result from `in`
                return true;
            }
            c.rs1 = rs, c.qs1 = qs;
            c.re1 = rs, c.qe1 = qs;
            c.i = 1;
            c.phase = PH_FILL_NEXT;
            break;
        }
        case PH_LEFT_DONE + 100: c.res = in; c.phase = PH_LEFT_DONE; break;
        case PH_LEFT_DONE: {
            Reg &r = c.regs[c.ireg];
            if (c.res.n_cigar > 0) {
                reg_append_cigar(c, r, c.res.n_cigar, c.res.cigar);
                r.dp_score += c.res.max;
            }
            c.rs1 = c.rs - (c.res.reach_end ? c.res.mqe_t + 1 : c.res.max_t + 1);
            c.qs1 = c.qs - (c.res.reach_end ? c.qs - c.qs0 : c.res.max_q + 1);
            c.re1 = c.rs, c.qe1 = c.qs;
            c.i = 1;
            c.phase = PH_FILL_NEXT;
            break;
        }
        case PH_FILL_NEXT: {
            const Anchor *a = c.a;
            bool emitted = false;
            while (c.i < c.cnt1) {
                int i = c.i;
                if ((a[c.as1 + i].y & (SEED_IGNORE | SEED_TANDEM)) && i != c.cnt1 - 1) { ++c.i; continue; }
                anchor_cut(c, a[c.as1 + i], &c.re, &c.qe);
                c.re1 = c.re, c.qe1 = c.qe;
                if (i == c.cnt1 - 1 || (a[c.as1 + i].y & SEED_LONG_JOIN) ||
                    (c.qe - c.qs >= o.min_ksw_len && c.re - c.rs >= o.min_ksw_len)) {
                    c.bw1 = c.bw_long;
                    if (a[c.as1 + i].y & SEED_LONG_JOIN) c.bw1 = c.qe - c.qs > c.re - c.rs ? c.qe - c.qs : c.re - c.rs;
                    int ql = c.qe - c.qs, tl = c.re - c.rs;
                    c.phase = PH_FILL1_DONE;
                    emitted = true;
                    if (o.max_sw_mat > 0 && (long long)tl * ql > o.max_sw_mat) { res_reset(c.res); c.res.zdropped = 1; break; }
                    if (ql <= 0 || tl <= 0) { res_reset(c.res); break; }
                    task_ext(task, &c.qseq[c.rev][c.qs], 1, ql, &c.tseq[c.rs], 1, tl, c.bw1, -1, o.zdrop, KSW_APPROX_MAX);
                    ++c.n_tasks;
                    c.phase = PH_FILL1_DONE + 100;
                    return true;
                }
                ++c.i;
            }
            if (!emitted) c.phase = PH_RIGHT;
            break;
        }
        case PH_FILL1_DONE + 100: c.res = in; c.phase = PH_FILL1_DONE; break;
        case PH_FILL1_DONE: {
            {   // the common case: the whole path loses less than either threshold, so the per-base rescan cannot fire
                const int32_t lim = o.zdrop < o.zdrop_inv ? o.zdrop : o.zdrop_inv;
                if (zdrop_bound(o, c.res.n_cigar, c.res.cigar, c.res.score) <= lim) { c.max_zdrop = 0; c.zdrop_code = 0; c.phase = PH_FILL_DECIDE; break; }
            }
            scan_zdrop(c, &c.qseq[c.rev][c.qs], &c.tseq[c.rs], c.res.n_cigar, c.res.cigar);
            int q_len = c.zpos[1][1] - c.zpos[0][1], t_len = c.zpos[1][0] - c.zpos[0][0];
            if (c.max_zdrop > o.zdrop_inv && q_len < o.max_gap && t_len < o.max_gap) {
                if (q_len > 0 && t_len > 0) {
                    task.kind = 1;
                    task.q = &c.qseq[c.rev][c.qs + c.zpos[1][1] - 1]; task.qstep = -1; task.qcomp = 1; task.qlen = q_len;
                    task.t = &c.tseq[c.rs + c.zpos[0][0]]; task.tstep = 1; task.tlen = t_len;
                    task.w = -1; task.zdrop = -1; task.end_bonus = 0; task.flag = 0;
                    ++c.n_tasks;
                    c.phase = PH_FILL_PROBE_DONE;
                    return true;
                }
                c.zdrop_code = c.max_zdrop > o.zdrop ? 1 : 0;     // empty probe scores 0
            } else c.zdrop_code = c.max_zdrop > o.zdrop ? 1 : 0;
            c.phase = PH_FILL_DECIDE;
            break;
        }
        case PH_FILL_PROBE_DONE: {
            int score = in.ll_score;
            if (score >= o.min_chain_score * o.a && score >= o.min_dp_max) c.zdrop_code = 2;
            else c.zdrop_code = c.max_zdrop > o.zdrop ? 1 : 0;
            c.phase = PH_FILL_DECIDE;
            break;
        }
        case PH_FILL_DECIDE: {
            if (c.zdrop_code != 0) {
                int ql = c.qe - c.qs, tl = c.re - c.rs;
                task_ext(task, &c.qseq[c.rev][c.qs], 1, ql, &c.tseq[c.rs], 1, tl, c.bw1, -1,
                         c.zdrop_code == 2 ? o.zdrop_inv : o.zdrop, 0);
                ++c.n_tasks;
                c.phase = PH_FILL2_DONE;
                return true;
            }
            c.phase = PH_FILL_CONSUME;
            break;
        }
        case PH_FILL2_DONE: c.res = in; c.phase = PH_FILL_CONSUME; break;
        case PH_FILL_CONSUME: {
            Reg &r = c.regs[c.ireg];
            if (c.res.n_cigar > 0) reg_append_cigar(c, r, c.res.n_cigar, c.res.cigar);
            if (c.res.zdropped) {
                const Anchor *a = c.a;
                int j;
                if (!r.has_p) { r.has_p = 1; r.dp_score = r.dp_max = r.dp_max2 = r.n_ambi = 0; r.n_cigar = 0; r.cig = c.cig_top; }
                for (j = c.i - 1; j >= 0; --j)
                    if ((int32_t)a[c.as1 + j].x <= c.rs + c.res.max_t) break;
                c.dropped = 1;
                if (j < 0) j = 0;
                r.dp_score += c.res.max;
                c.re1 = c.rs + (c.res.max_t + 1);
                c.qe1 = c.qs + (c.res.max_q + 1);
                if (c.cnt1 - (j + 1) >= o.min_cnt) {
                    reg_split(r, c.r2, c.as1 + j + 1 - r.as, c.qlen, a);
                    if (c.zdrop_code == 2) c.r2.split_inv = 1;
                }
                c.phase = PH_REG_FINAL;
                break;
            }
            r.dp_score += c.res.score;
            c.rs = c.re, c.qs = c.qe;
            ++c.i;
            c.phase = PH_FILL_NEXT;
            break;
        }
        case PH_RIGHT: {
            if (!c.dropped && c.qe < c.qe0 && c.re < c.re0) {
                int ql = c.qe0 - c.qe, tl = c.re0 - c.re;
                c.phase = PH_RIGHT_DONE;
                if (o.max_sw_mat > 0 && (long long)tl * ql > o.max_sw_mat) { res_reset(c.res); c.res.zdropped = 1; break; }
                task_ext(task, &c.qseq[c.rev][c.qe], 1, ql, &c.tseq[c.re], 1, tl, c.bw, o.end_bonus, o.zdrop, KSW_EXTZ_ONLY);
                ++c.n_tasks;
                c.phase = PH_RIGHT_DONE + 100;
                return true;
            }
            c.phase = PH_REG_FINAL;
            break;
        }
        case PH_RIGHT_DONE + 100: c.res = in; c.phase = PH_RIGHT_DONE; break;
        case PH_RIGHT_DONE: {
            Reg &r = c.regs[c.ireg];
            if (c.res.n_cigar > 0) {
                reg_append_cigar(c, r, c.res.n_cigar, c.res.cigar);
                r.dp_score += c.res.max;
            }
            c.re1 = c.re + (c.res.reach_end ? c.res.mqe_t + 1 : c.res.max_t + 1);
            c.qe1 = c.qe + (c.res.reach_end ? c.qe0 - c.qe : c.res.max_q + 1);
            c.phase = PH_REG_FINAL;
            break;
        }
        case PH_REG_FINAL: {
            Reg &r = c.regs[c.ireg];
            r.rs = c.rs1, r.re = c.re1;
            if (c.rev) r.qs = c.qlen - c.qe1, r.qe = c.qlen - c.qs1;
            else r.qs = c.qs1, r.qe = c.qe1;
            if (r.has_p) {
                // statistics of the joined CIGAR are only needed right away when an inversion test follows
                if (c.defer_finish && !r.split_inv && !(c.r2.cnt > 0 && c.r2.split_inv)) r.need_fin = 1, r.fin_q = c.qs1, r.fin_t = c.rs1;
                else reg_finish(c, r, &c.qseq[r.rev][c.qs1], &c.tseq[c.rs1]);
            }
            if (c.r2.cnt > 0) regs_insert(c, c.r2, c.ireg);
            c.phase = PH_NEXT_REG;
            if (c.ireg > 0 && c.regs[c.ireg].split_inv) {      // inversion between the two halves of a split?
                const Reg &r1 = c.regs[c.ireg - 1], &rr = c.regs[c.ireg];
                bool ok = (r1.split & 1) && (rr.split & 2);
                if (ok && r1.id != r1.parent && r1.parent != PARENT_TMP_PRI) ok = false;
                if (ok && rr.id != rr.parent && rr.parent != PARENT_TMP_PRI) ok = false;
                if (ok && r1.rev != rr.rev) ok = false;
                if (ok) {
                    int ql = r1.rev ? r1.qs - rr.qe : rr.qs - r1.qe, tl = rr.rs - r1.re;
                    if (ql < o.min_chain_score || ql > o.max_gap || tl < o.min_chain_score || tl > o.max_gap) ok = false;
                    if (ok) {
                        c.inv_ql = ql, c.inv_tl = tl;
                        c.inv_q = r1.rev ? &c.qseq[0][rr.qe] : &c.qseq[1][c.qlen - rr.qs];
                        task.kind = 1;
                        task.q = c.inv_q + ql - 1; task.qstep = -1; task.qcomp = 0; task.qlen = ql;
                        task.t = &c.tseq[r1.re + tl - 1]; task.tstep = -1; task.tlen = tl;
                        task.w = -1; task.zdrop = -1; task.end_bonus = 0; task.flag = 0;
                        ++c.n_tasks;
                        c.phase = PH_INV_LL_DONE;
                        return true;
                    }
                }
            }
            break;
        }
        case PH_INV_LL_DONE: {
            c.phase = PH_NEXT_REG;
            if (in.ll_score < o.min_dp_max) break;
            const Reg &r1 = c.regs[c.ireg - 1];
            c.inv_qoff = c.inv_ql - (in.ll_qe + 1), c.inv_toff = c.inv_tl - (in.ll_te + 1);
            int ql = c.inv_ql - c.inv_qoff, tl = c.inv_tl - c.inv_toff;
            if (o.max_sw_mat > 0 && (long long)tl * ql > o.max_sw_mat) break;
            task_ext(task, c.inv_q + c.inv_qoff, 1, ql, &c.tseq[r1.re + c.inv_toff], 1, tl, (int)(o.bw * 1.5), -1, o.zdrop, KSW_EXTZ_ONLY);
            ++c.n_tasks;
            c.phase = PH_INV_EXT_DONE;
            return true;
        }
        case PH_INV_EXT_DONE: {
            c.phase = PH_NEXT_REG;
            if (in.n_cigar == 0) break;
            const Reg r1 = c.regs[c.ireg - 1], rr = c.regs[c.ireg];
            Reg ri;
            reg_clear(ri);
            reg_append_cigar(c, ri, in.n_cigar, in.cigar);
            ri.dp_score = in.max;
            ri.id = -1;
            ri.parent = PARENT_UNSET;
            ri.inv = 1;
            ri.rev = !r1.rev;
            if (ri.rev == 0) {
                ri.qs = rr.qe + c.inv_qoff;
                ri.qe = ri.qs + in.max_q + 1;
            } else {
                ri.qe = rr.qs - c.inv_qoff;
                ri.qs = ri.qe - (in.max_q + 1);
            }
            ri.rs = r1.re + c.inv_toff;
            ri.re = ri.rs + in.max_t + 1;
            reg_finish(c, ri, c.inv_q + c.inv_qoff, &c.tseq[r1.re + c.inv_toff]);
            if (regs_insert(c, ri, c.ireg)) ++c.ireg;   // skip the inserted inversion record
            break;
        }
        case PH_NEXT_REG: ++c.ireg; c.phase = PH_REG_BEGIN; break;
        case PH_FINISH: {
            if (c.defer_finish) return false;        // the caller runs aln_finish() once every DP request is served
            regs_filter(o, c.qlen, &c.n_regs, c.regs);
            if (c.qlen >= o.rank_min_len) {
                regs_update_dp_max(c);
                regs_filter(o, c.qlen, &c.n_regs, c.regs);
            }
            regs_sort(&c.n_regs, c.regs, c.hs);
            regs_set_parent(o, c.n_regs, c.regs, c.hs);
            regs_select_sub(o, 0, &c.n_regs, c.regs, c.hs, c.cap_regs);
            regs_set_sam_pri(c.n_regs, c.regs);
            c.phase = PH_DONE;
            return false;
        }
        default: return false;
        }
    }
}

// deferred tail of the coroutine: per-region statistics, then the final filter / sort / primary assignment
TELR_HDN void aln_finish(AlnCtx &c)
{
    const Opt &o = *c.o;
    for (int i = 0; i < c.n_regs; ++i) {
        Reg &r = c.regs[i];
        if (r.need_fin) { r.need_fin = 0; reg_finish(c, r, &c.qseq[r.rev][r.fin_q], &c.tseq[r.fin_t]); }
    }
    regs_filter(o, c.qlen, &c.n_regs, c.regs);
    if (c.qlen >= o.rank_min_len) {
        regs_update_dp_max(c);
        regs_filter(o, c.qlen, &c.n_regs, c.regs);
    }
    regs_sort(&c.n_regs, c.regs, c.hs);
    regs_set_parent(o, c.n_regs, c.regs, c.hs);
    regs_select_sub(o, 0, &c.n_regs, c.regs, c.hs, c.cap_regs);
    regs_set_sam_pri(c.n_regs, c.regs);
    regs_set_mapq(o, c.n_regs, c.regs, c.rep_len, c.hs);
    c.phase = PH_DONE;
}

}  // namespace telr
