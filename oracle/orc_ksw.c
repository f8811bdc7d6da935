/*
 * oracle/orc_ksw.c — base-level DP restatements.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (orc.h).
 *
 * orc_ksw_extd2 follows [UP] ksw2_extd2_sse.c ksw_extd2_sse (minimap2 2.22), reached from the
 * reference at TELR_te.py:505 through mm_align1 -> mm_align_pair: Suzuki-Kasahara difference
 * recurrence on anti-diagonals with two-piece affine gaps, direction bytes, exact / approximate
 * max tracking, z-drop, and ksw_backtrack.
 *
 * Stated deviation ("clean band"): the SSE code rounds each anti-diagonal's band [st0,en0] out to
 * 16-lane vectors and computes scratch cells there from stale inputs; those cells can feed the two
 * edge cells of the next anti-diagonal.  This restatement computes only [st0,en0] and gives cells
 * outside the previous band the seeds the SSE code uses for never-computed columns
 * (x=v=u=y=-q-e, x2=y2=-q2-e2).  The two agree unless the optimal path runs along the band edge.
 */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "orc.h"

int orc_ext_bound = 1;     /* bounded extension rule on (see orc_ksw_extd2); tests switch it off to show that no output depends on it */
void orc_set_ext_bound(int on) { orc_ext_bound = on; }

#define KSW_NEG_INF (-0x40000000)

static void ez_reset(orc_ez_t *ez)
{
    ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
    ez->max = 0; ez->score = ez->mqe = ez->mte = KSW_NEG_INF;
    ez->n_cigar = 0; ez->zdropped = 0; ez->reach_end = 0;
}

static void push_cigar(orc_ez_t *ez, uint32_t op, int len)
{
    if (ez->n_cigar == 0 || op != (ez->cigar[ez->n_cigar - 1] & 0xf)) {
        if (ez->n_cigar == ez->m_cigar) {
            ez->m_cigar = ez->m_cigar ? ez->m_cigar << 1 : 4;
            ez->cigar = (uint32_t *)realloc(ez->cigar, (size_t)ez->m_cigar * 4);
        }
        ez->cigar[ez->n_cigar++] = (uint32_t)len << 4 | op;
    } else ez->cigar[ez->n_cigar - 1] += (uint32_t)len << 4;
}

/* [UP] ksw2.h ksw_apply_zdrop (is_rot = 1) */
static int apply_zdrop(orc_ez_t *ez, int32_t H, int r, int t, int zdrop, int e)
{
    if (H > ez->max) {
        ez->max = H, ez->max_t = t, ez->max_q = r - t;
    } else if (t >= ez->max_t && r - t >= ez->max_q) {
        int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l;
        l = tl > ql ? tl - ql : ql - tl;
        if (zdrop >= 0 && ez->max - H > zdrop + l * e) {
            ez->zdropped = 1;
            return 1;
        }
    }
    return 0;
}

/* [UP] ksw2.h ksw_backtrack (is_rot = 1, min_intron_len = 0) */
static void backtrack(orc_ez_t *ez, int is_rev, const uint8_t *p, const int64_t *poff, const int *off,
                      const int *off_end, int i0, int j0, int qlen, int tlen)
{
    int i = i0, j = j0, r, state = 0, near_edge = 0;
    uint32_t tmp;
    ez->n_cigar = 0;
    while (i >= 0 && j >= 0) {
        int force_state = -1;
        r = i + j;
        if (i < off[r]) force_state = 2;
        if (i > off_end[r]) force_state = 1;
        {   /* deviation counter: the path is within one cell of an edge that the band (not a sequence end) set */
            int lo = r - qlen + 1 > 0 ? r - qlen + 1 : 0, hi = r < tlen - 1 ? r : tlen - 1;
            if ((off[r] > lo && i <= off[r] + 1) || (off_end[r] < hi && i >= off_end[r] - 1)) near_edge = 1;
        }
        tmp = force_state < 0 ? p[poff[r] + (i - off[r])] : 0;
        if (state == 0) state = tmp & 7;
        else if (!(tmp >> (state + 2) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (force_state >= 0) state = force_state;
        if (state == 0) push_cigar(ez, 0, 1), --i, --j;
        else if (state == 1 || state == 3) push_cigar(ez, 2, 1), --i;
        else push_cigar(ez, 1, 1), --j;
    }
    if (i >= 0) push_cigar(ez, 2, i + 1);
    if (j >= 0) push_cigar(ez, 1, j + 1);
    if (near_edge) {
#pragma omp atomic
        ++orc_dev[2];
    }
    if (!is_rev)
        for (i = 0; i < ez->n_cigar >> 1; ++i)
            tmp = ez->cigar[i], ez->cigar[i] = ez->cigar[ez->n_cigar - 1 - i], ez->cigar[ez->n_cigar - 1 - i] = tmp;
}

void orc_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int sc_a, int sc_b, int sc_ambi, int q, int e, int q2, int e2,
                   int w, int zdrop, int end_bonus, int flag, orc_ez_t *ez)
{
    int r, t, qe, qe2, long_thres, long_diff, pst = -1, pen = -1;
    int approx_max = !!(flag & ORC_KSW_APPROX_MAX), right = !!(flag & ORC_KSW_RIGHT);
    int32_t *u, *v, *x, *y, *x2, *y2, *H = 0, H0 = 0, last_H0_t = 0;
    uint8_t *p;
    int64_t *poff, ptot = 0, pcap;
    int *off, *off_end;

    ez_reset(ez);
    ez->cells = 0;
    if (qlen <= 0 || tlen <= 0) return;
    if (q2 + e2 < q + e) t = q, q = q2, q2 = t, t = e, e = e2, e2 = t;
    qe = q + e, qe2 = q2 + e2;
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    if (w < (tlen > qlen ? tlen : qlen)) {
#pragma omp atomic
        ++orc_dev[5];
    }
    {   /* -min_sc > 2*(q+e): no mismatch would ever be seen */
        int min_sc = -sc_b < -sc_ambi ? -sc_b : -sc_ambi;
        if (-min_sc > 2 * (q + e)) return;
    }
    long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
    if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
    long_diff = long_thres * (e - e2) - (q2 - q) - e2;

    u = (int32_t *)malloc((size_t)tlen * 6 * 4);
    v = u + tlen, x = v + tlen, y = x + tlen, x2 = y + tlen, y2 = x2 + tlen;
    if (!approx_max) {
        H = (int32_t *)malloc((size_t)tlen * 4);
        for (t = 0; t < tlen; ++t) H[t] = KSW_NEG_INF;
    }
    {
        int ncol = qlen < tlen ? qlen : tlen;
        if (ncol > w + 1) ncol = w + 1;
        pcap = (int64_t)(qlen + tlen - 1) * ncol;
    }
    p = (uint8_t *)malloc((size_t)pcap + 1);
    poff = (int64_t *)malloc((size_t)(qlen + tlen) * 8);
    off = (int *)malloc((size_t)(qlen + tlen) * 2 * sizeof(int));
    off_end = off + qlen + tlen;

    for (r = 0; r < qlen + tlen - 1; ++r) {
        int st = 0, en = tlen - 1;
        int32_t x1, x21, v1;
        uint8_t *pr;
        /* Bounded extension (an exact work-saving rule of this project, not of ksw2; orc_ext_bound = 0 switches it off).  An
         * extension only reports its maximum (max, max_t, max_q; align.c never reads zdropped of an extension, and with
         * end_bonus <= 0 the query-end rule `mqe + end_bonus > max` cannot hold since mqe <= max).  A cell (i, j) of
         * anti-diagonal r scores at most sc_a * min(i + 1, j + 1) - gap(|i - j|): that many matches at best, and the offset
         * between the two coordinates has to be paid for by gaps, a single gap being the cheapest way.  Once a sequence is
         * exhausted, |i - j| >= r - 2 (len - 1) grows with r, so as soon as the bound is not above the maximum found so far
         * no later anti-diagonal can change the result: stop.  Typical case: a read overhanging the contig end by thousands
         * of bases, where ksw2 keeps ~750 anti-diagonals of a few dozen cells alive until the band runs out. */
        if (orc_ext_bound && (flag & ORC_KSW_EXTZ_ONLY) && !approx_max && end_bonus <= 0) {
            int dmin = r - 2 * (tlen - 1) > r - 2 * (qlen - 1) ? r - 2 * (tlen - 1) : r - 2 * (qlen - 1);
            if (dmin > 0) {
                int mcap = tlen < qlen ? tlen : qlen;
                int64_t g1 = (int64_t)q + (int64_t)e * dmin, g2 = (int64_t)q2 + (int64_t)e2 * dmin, ub;
                if ((r >> 1) + 1 < mcap) mcap = (r >> 1) + 1;
                ub = (int64_t)sc_a * mcap - (g1 < g2 ? g1 : g2);
                if (ub <= ez->max) { ez->zdropped = 1; break; }
            }
        }
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
        if (en > (r + w) >> 1) en = (r + w) >> 1;
        if (st > en) {
            ez->zdropped = 1;
            break;
        }
        ez->cells += en - st + 1;
        /* left neighbour of the first cell */
        if (st > 0) {
            if (st - 1 >= pst && st - 1 <= pen) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
            else x1 = -q - e, x21 = -q2 - e2, v1 = -q - e;
        } else {
            x1 = -q - e, x21 = -q2 - e2;
            v1 = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
        }
        off[r] = st, off_end[r] = en, poff[r] = ptot;
        pr = p + ptot;
        ptot += en - st + 1;
        assert(ptot <= pcap);
        for (t = st; t <= en; ++t) {
            int32_t ut, yt, y2t, z, a, b, a2, b2, tmp;
            uint8_t d;
            int qc = query[r - t], tc = target[t];
            if (t >= pst && t <= pen) ut = u[t], yt = y[t], y2t = y2[t];
            else if (t == r) {
                yt = -q - e, y2t = -q2 - e2;
                ut = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
            } else ut = -q - e, yt = -q - e, y2t = -q2 - e2;
            z = (qc > 3 || tc > 3) ? -sc_ambi : qc == tc ? sc_a : -sc_b;
            a = x1 + v1, b = yt + ut, a2 = x21 + v1, b2 = y2t + ut;
            /* carry this column's old x, x2, v to the next cell before overwriting them */
            {
                int32_t nx1, nx21, nv1;
                if (t >= pst && t <= pen) nx1 = x[t], nx21 = x2[t], nv1 = v[t];
                else nx1 = -q - e, nx21 = -q2 - e2, nv1 = -q - e;
                if (!right) {
                    d = a > z ? 1 : 0;  z = z > a ? z : a;
                    d = b > z ? 2 : d;  z = z > b ? z : b;
                    d = a2 > z ? 3 : d; z = z > a2 ? z : a2;
                    d = b2 > z ? 4 : d; z = z > b2 ? z : b2;
                } else {
                    d = z > a ? 0 : 1;  z = z > a ? z : a;
                    d = z > b ? d : 2;  z = z > b ? z : b;
                    d = z > a2 ? d : 3; z = z > a2 ? z : a2;
                    d = z > b2 ? d : 4; z = z > b2 ? z : b2;
                }
                if (z > sc_a) z = sc_a;
                u[t] = z - v1;
                v[t] = z - ut;
                tmp = z - q;  a -= tmp, b -= tmp;
                tmp = z - q2; a2 -= tmp, b2 -= tmp;
                if (!right) {
                    x[t] = (a > 0 ? a : 0) - qe;     d |= a > 0 ? 0x08 : 0;
                    y[t] = (b > 0 ? b : 0) - qe;     d |= b > 0 ? 0x10 : 0;
                    x2[t] = (a2 > 0 ? a2 : 0) - qe2; d |= a2 > 0 ? 0x20 : 0;
                    y2[t] = (b2 > 0 ? b2 : 0) - qe2; d |= b2 > 0 ? 0x40 : 0;
                } else {
                    x[t] = (a >= 0 ? a : 0) - qe;     d |= a >= 0 ? 0x08 : 0;
                    y[t] = (b >= 0 ? b : 0) - qe;     d |= b >= 0 ? 0x10 : 0;
                    x2[t] = (a2 >= 0 ? a2 : 0) - qe2; d |= a2 >= 0 ? 0x20 : 0;
                    y2[t] = (b2 >= 0 ? b2 : 0) - qe2; d |= b2 >= 0 ? 0x40 : 0;
                }
                pr[t - st] = d;
                x1 = nx1, x21 = nx21, v1 = nv1;
            }
        }
        if (!approx_max) {
            int32_t max_H, max_t;
            if (r > 0) {
                int32_t HH[4], tt[4], en1 = st + (en - st) / 4 * 4, i;
                max_H = H[en] = en > 0 ? H[en - 1] + u[en] : H[en] + v[en];
                max_t = en;
                for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
                for (t = st; t < en1; t += 4)
                    for (i = 0; i < 4; ++i) {
                        H[t + i] += v[t + i];
                        if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
                    }
                for (i = 0; i < 4; ++i)
                    if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
                for (; t < en; ++t) {
                    H[t] += v[t];
                    if (H[t] > max_H) max_H = H[t], max_t = t;
                }
            } else H[0] = v[0] - qe, max_H = H[0], max_t = 0;
            if (en == tlen - 1 && H[en] > ez->mte) ez->mte = H[en], ez->mte_q = r - en;
            if (r - st == qlen - 1 && H[st] > ez->mqe) ez->mqe = H[st], ez->mqe_t = st;
            if (apply_zdrop(ez, max_H, r, max_t, zdrop, e2)) break;
            if (r == qlen + tlen - 2 && en == tlen - 1) ez->score = H[tlen - 1];
        } else {
            if (r > 0) {
                if (last_H0_t >= st && last_H0_t <= en && last_H0_t + 1 >= st && last_H0_t + 1 <= en) {
                    int32_t d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
                    if (d0 > d1) H0 += d0;
                    else H0 += d1, ++last_H0_t;
                } else if (last_H0_t >= st && last_H0_t <= en) {
                    H0 += v[last_H0_t];
                } else {
                    ++last_H0_t, H0 += u[last_H0_t];
                }
            } else H0 = v[0] - qe, last_H0_t = 0;
            if (r == qlen + tlen - 2 && en == tlen - 1) ez->score = H0;
        }
        pst = st, pen = en;
    }
    free(u);
    free(H);
    {
        int rev_cigar = !!(flag & ORC_KSW_REV_CIGAR);
        if (!ez->zdropped && !(flag & ORC_KSW_EXTZ_ONLY)) {
            backtrack(ez, rev_cigar, p, poff, off, off_end, tlen - 1, qlen - 1, qlen, tlen);
        } else if (!ez->zdropped && (flag & ORC_KSW_EXTZ_ONLY) && ez->mqe + end_bonus > ez->max) {
            ez->reach_end = 1;
            backtrack(ez, rev_cigar, p, poff, off, off_end, ez->mqe_t, qlen - 1, qlen, tlen);
        } else if (ez->max_t >= 0 && ez->max_q >= 0) {
            backtrack(ez, rev_cigar, p, poff, off, off_end, ez->max_t, ez->max_q, qlen, tlen);
        }
    }
    free(p); free(poff); free(off);
}

/* [UP] ksw2_ll_sse.c ksw_ll_i16: local affine Smith-Waterman (gap of length l costs gapo + l*gape).
 * Only the score is consumed on the inversion-probe path (align.c mm_test_zdrop); mm_align1_inv also
 * consumes the end coordinates, for which the striped SSE code's tie order is restated as "first
 * maximum in target-major, query-minor scan" (stated deviation; unobservable unless maxima tie). */
int orc_ksw_ll(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
               int sc_a, int sc_b, int sc_ambi, int gapo, int gape, int *qe_, int *te_)
{
    int i, j, gmax = 0, gqe = -1, gte = -1, gapoe = gapo + gape, best = 0, n_best = 0;
    int32_t *H, *E;
    if (qe_) *qe_ = -1;
    if (te_) *te_ = -1;
    if (qlen <= 0 || tlen <= 0) return 0;
    H = (int32_t *)calloc((size_t)qlen + 1, 4);
    E = (int32_t *)calloc((size_t)qlen + 1, 4);
    for (i = 0; i < tlen; ++i) {
        int32_t f = 0, hdiag = 0, tc = target[i], imax = 0, iqe = -1;
        for (j = 0; j < qlen; ++j) {
            int qc = query[j];
            int32_t s = (qc > 3 || tc > 3) ? -sc_ambi : qc == tc ? sc_a : -sc_b;
            int32_t h = hdiag + s, eij = E[j + 1];
            hdiag = H[j + 1];
            if (h < eij) h = eij;
            if (h < f) h = f;
            if (h < 0) h = 0;
            H[j + 1] = h;
            if (h > imax) imax = h, iqe = j;
            if (h > best) best = h, n_best = 1; else if (h == best) ++n_best;
            eij -= gape; if (eij < h - gapoe) eij = h - gapoe; if (eij < 0) eij = 0;
            E[j + 1] = eij;
            f -= gape; if (f < h - gapoe) f = h - gapoe; if (f < 0) f = 0;
        }
        if (imax > gmax) gmax = imax, gte = i, gqe = iqe;
    }
    free(H); free(E);
    {
#pragma omp atomic
        ++orc_dev[7];
    }
    if (best > 0 && n_best > 1) {
#pragma omp atomic
        ++orc_dev[3];
    }
    if (qe_) *qe_ = gqe;
    if (te_) *te_ = gte;
    return gmax;
}
