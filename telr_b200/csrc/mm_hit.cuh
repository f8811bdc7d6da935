// mm_hit.cuh — region bookkeeping for one read x contig strand (minimap2 hit.c semantics): region
// generation from chains, primary/secondary assignment, sub-optimal selection, filtering, sorting.
// Sequential; runs on one lane between the warp-parallel stages (or on the host in tests/emu).
#pragma once
#include <math.h>
#include "mm_sort.cuh"

namespace telr {

struct HitScratch {
    Anchor *z;        // [cap]
    uint64_t *cov;    // [cap]
    int32_t *w;       // [cap]
    Reg *tmp;         // [cap]
    int32_t *sortws;  // rs_scratch_words(cap)
};
TELR_HD size_t hit_scratch_bytes(size_t cap)
{
    size_t s = (cap + 1) * (sizeof(Anchor) + 8 + 4 + sizeof(Reg)) + (size_t)rs_scratch_words((int)cap) * 4;
    return (s + 255) / 256 * 256;
}
TELR_HD void hit_scratch_carve(HitScratch &s, uint8_t *base, size_t cap)
{
    s.z = (Anchor *)base; base += (cap + 1) * sizeof(Anchor);
    s.cov = (uint64_t *)base; base += (cap + 1) * 8;
    s.tmp = (Reg *)base; base += (cap + 1) * sizeof(Reg);
    s.w = (int32_t *)base; base += (cap + 1) * 4;
    s.sortws = (int32_t *)base;
}

TELR_HD void reg_set_coor(Reg &r, int qlen, const Anchor *a)
{
    int k = r.as, span = (int)(a[k].y >> 32 & 0xff);
    r.rev = (uint8_t)(a[k].x >> 63);
    r.rs = (int32_t)a[k].x + 1 > span ? (int32_t)a[k].x + 1 - span : 0;
    r.re = (int32_t)a[k + r.cnt - 1].x + 1;
    if (!r.rev) {
        r.qs = (int32_t)a[k].y + 1 - span;
        r.qe = (int32_t)a[k + r.cnt - 1].y + 1;
    } else {
        r.qs = qlen - ((int32_t)a[k + r.cnt - 1].y + 1);
        r.qe = qlen - ((int32_t)a[k].y + 1 - span);
    }
}

TELR_HD void reg_fuzzy_len(Reg &r, const Anchor *a)
{
    r.mlen = r.blen = 0;
    if (r.cnt <= 0) return;
    r.mlen = r.blen = (int32_t)(a[r.as].y >> 32 & 0xff);
    for (int i = r.as + 1; i < r.as + r.cnt; ++i) {
        int span = (int)(a[i].y >> 32 & 0xff);
        int tl = (int32_t)a[i].x - (int32_t)a[i - 1].x;
        int ql = (int32_t)a[i].y - (int32_t)a[i - 1].y;
        r.blen += tl > ql ? tl : ql;
        r.mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
    }
}

TELR_HD void reg_clear(Reg &r)
{
    r.id = r.cnt = r.score = r.qs = r.qe = r.rs = r.re = r.parent = r.subsc = r.as = r.mlen = r.blen = r.n_sub = r.score0 = 0;
    r.hash = 0;
    r.rev = r.inv = r.sam_pri = r.split = r.split_inv = r.strand_retained = r.has_p = r.need_fin = 0;
    r.fin_q = r.fin_t = 0;
    r.dp_score = r.dp_max = r.dp_max2 = r.n_ambi = r.n_cigar = 0;
    r.cig = 0;
}

// chains (u, a) -> regions sorted by score (ties by a hash of anchors and read name)
TELR_HDN void regs_from_chains(uint32_t hash, int qlen, int n_u, const uint64_t *u, const Anchor *a, Reg *r, HitScratch &s)
{
    Anchor *z = s.z;
    int k = 0;
    for (int i = 0; i < n_u; ++i) {
        uint32_t h = (uint32_t)mix64((mix64(a[k].x) + mix64(a[k].y)) ^ hash);
        z[i].x = u[i] ^ h;
        z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
        k += (int32_t)u[i];
    }
    rs_sort_emul(z, n_u, KeyX(), s.sortws);
    for (int i = 0; i < n_u >> 1; ++i) { Anchor t = z[i]; z[i] = z[n_u - 1 - i]; z[n_u - 1 - i] = t; }
    for (int i = 0; i < n_u; ++i) {
        Reg &ri = r[i];
        reg_clear(ri);
        ri.id = i;
        ri.parent = PARENT_UNSET;
        ri.score = ri.score0 = (int32_t)(z[i].x >> 32);
        ri.hash = (uint32_t)z[i].x;
        ri.cnt = (int32_t)z[i].y;
        ri.as = (int32_t)(z[i].y >> 32);
        reg_set_coor(ri, qlen, a);
        reg_fuzzy_len(ri, a);
    }
}

TELR_HD void regs_set_sam_pri(int n, Reg *r)
{
    int n_pri = 0;
    for (int i = 0; i < n; ++i)
        if (r[i].id == r[i].parent) {
            ++n_pri;
            r[i].sam_pri = (n_pri == 1);
        } else r[i].sam_pri = 0;
}

TELR_HDN void regs_sync(int n, Reg *regs, HitScratch &s, int cap)
{
    if (n <= 0) return;
    int max_id = -1;
    int32_t *tmp = s.w;
    for (int i = 0; i < n; ++i) max_id = max_id > regs[i].id ? max_id : regs[i].id;
    int n_tmp = max_id + 1;
    if (n_tmp > cap) n_tmp = cap;
    for (int i = 0; i < n_tmp; ++i) tmp[i] = -1;
    for (int i = 0; i < n; ++i)
        if (regs[i].id >= 0 && regs[i].id < n_tmp) tmp[regs[i].id] = i;
    for (int i = 0; i < n; ++i) {
        Reg &r = regs[i];
        r.id = i;
        if (r.parent == PARENT_TMP_PRI) r.parent = i;
        else if (r.parent >= 0 && r.parent < n_tmp && tmp[r.parent] >= 0) r.parent = tmp[r.parent];
        else r.parent = PARENT_UNSET;
    }
    regs_set_sam_pri(n, regs);
}

TELR_HDN void regs_set_parent(const Opt &o, int n, Reg *r, HitScratch &s)
{
    if (n <= 0) return;
    const int sub_diff = o.a * 2 + o.b;
    uint64_t *cov = s.cov;
    int32_t *w = s.w;
    for (int i = 0; i < n; ++i) r[i].id = i;
    w[0] = 0, r[0].parent = 0;
    int k = 1;
    for (int i = 1; i < n; ++i) {
        Reg &ri = r[i];
        int si = ri.qs, ei = ri.qe, n_cov = 0, uncov_len = 0, j;
        for (j = 0; j < k; ++j) {
            const Reg &rp = r[w[j]];
            int sj = rp.qs, ej = rp.qe;
            if (ej <= si || sj >= ei) continue;
            if (sj < si) sj = si;
            if (ej > ei) ej = ei;
            cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
        }
        if (n_cov > 0) {
            int x = si;
            rs_sort_emul(cov, n_cov, KeyId(), s.sortws);
            for (int jj = 0; jj < n_cov; ++jj) {
                if ((int)(cov[jj] >> 32) > x) uncov_len += (int)(cov[jj] >> 32) - x;
                x = (int32_t)cov[jj] > x ? (int32_t)cov[jj] : x;
            }
            if (ei > x) uncov_len += ei - x;
            for (j = 0; j < k; ++j) {
                Reg &rp = r[w[j]];
                int sj = rp.qs, ej = rp.qe;
                if (ej <= si || sj >= ei) continue;
                int mn = ej - sj < ei - si ? ej - sj : ei - si;
                int mx = ej - sj > ei - si ? ej - sj : ei - si;
                int ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
                if ((float)ol / mn - (float)uncov_len / mx > o.mask_level && uncov_len <= o.mask_len) {
                    int cnt_sub = 0, sci = ri.score;
                    ri.parent = rp.parent;
                    rp.subsc = rp.subsc > sci ? rp.subsc : sci;
                    if (ri.cnt >= rp.cnt) cnt_sub = 1;
                    if (rp.has_p && ri.has_p && (rp.rs != ri.rs || rp.re != ri.re || ol != mn)) {
                        sci = ri.dp_max;
                        rp.dp_max2 = rp.dp_max2 > sci ? rp.dp_max2 : sci;
                        if (rp.dp_max - ri.dp_max <= sub_diff) cnt_sub = 1;
                    }
                    if (cnt_sub) ++rp.n_sub;
                    break;
                }
            }
        } else j = k;
        if (j == k) w[k++] = i, ri.parent = i, ri.n_sub = 0;
    }
}

TELR_HDN void regs_select_sub(const Opt &o, int check_strand, int *n_, Reg *r, HitScratch &s, int cap)
{
    const int min_diff = o.k * 2, min_strand_sc = (int)(o.max_gap * 0.8);
    if (o.pri_ratio > 0.0f && *n_ > 0) {
        int k = 0, n = *n_, n_2nd = 0;
        for (int i = 0; i < n; ++i) {
            int p = r[i].parent;
            if (p == i || r[i].inv) {
                r[k++] = r[i];
            } else if ((r[i].score >= r[p].score * o.pri_ratio || r[i].score + min_diff >= r[p].score) && n_2nd < o.best_n) {
                if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rs == r[p].rs && r[i].re == r[p].re)) r[k++] = r[i], ++n_2nd;
            } else if (check_strand && n_2nd < o.best_n && r[i].score > min_strand_sc && r[p].rev != r[i].rev) {
                r[i].strand_retained = 1;
                r[k++] = r[i], ++n_2nd;
            }
        }
        if (k != n) regs_sync(k, r, s, cap);
        *n_ = k;
    }
}

TELR_HDN void regs_filter(const Opt &o, int qlen, int *n_regs, Reg *regs)
{
    int k = 0;
    for (int i = 0; i < *n_regs; ++i) {
        Reg &r = regs[i];
        int flt = 0;
        if (!r.inv && r.cnt < o.min_cnt) flt = 1;
        if (r.has_p) {
            if (r.mlen < o.min_chain_score) flt = 1;
            else if (r.dp_max < o.min_dp_max) flt = 1;
            else if (r.qs > qlen * o.max_clip_ratio && qlen - r.qe > qlen * o.max_clip_ratio) flt = 1;
        }
        if (!flt) {
            if (k < i) regs[k++] = regs[i];
            else ++k;
        }
    }
    *n_regs = k;
}

TELR_HDN void regs_sort(int *n_regs, Reg *r, HitScratch &s)
{
    int n = *n_regs, n_aux = 0;
    if (n <= 1) return;
    Anchor *aux = s.z;
    for (int i = 0; i < n; ++i)
        if (r[i].inv || r[i].cnt > 0) {
            int score = r[i].has_p ? r[i].dp_max : r[i].score;
            aux[n_aux].x = (uint64_t)(uint32_t)score << 32 | r[i].hash;
            aux[n_aux++].y = (uint64_t)i;
        }
    rs_sort_emul(aux, n_aux, KeyX(), s.sortws);
    for (int i = n_aux - 1; i >= 0; --i) s.tmp[n_aux - 1 - i] = r[aux[i].y];
    for (int i = 0; i < n_aux; ++i) r[i] = s.tmp[i];
    *n_regs = n_aux;
}

// minimap2 hit.c mm_set_mapq (long reads) + mm_set_inv_mapq.  logf = correctly rounded float logarithm through fp64,
// the same on the host and on the device.
TELR_HD float logf_cr(float x) { return (float)log((double)x); }
TELR_HDN void regs_set_mapq(const Opt &o, int n_regs, Reg *regs, int rep_len, HitScratch &s)
{
    const float q_coef = 40.0f;
    const int min_chain_sc = o.min_chain_score, match_sc = o.a;
    long long sum_sc = 0;
    if (n_regs == 0) return;
    for (int i = 0; i < n_regs; ++i)
        if (regs[i].parent == regs[i].id) sum_sc += regs[i].score;
    const float uniq_ratio = (float)sum_sc / (float)(sum_sc + rep_len);
    for (int i = 0; i < n_regs; ++i) {
        Reg &r = regs[i];
        if (r.inv) r.mapq = 0;
        else if (r.parent == r.id) {
            int mq;
            float pen_s1 = TELR_FMUL(r.score > 100 ? 1.0f : TELR_FMUL(0.01f, (float)r.score), uniq_ratio);
            float pen_cm = r.cnt > 10 ? 1.0f : TELR_FMUL(0.1f, (float)r.cnt);
            pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
            const int subsc = r.subsc > min_chain_sc ? r.subsc : min_chain_sc;
            if (r.has_p && r.dp_max2 > 0 && r.dp_max > 0) {
                const float identity = (float)r.mlen / (float)r.blen;
                const float x = TELR_FMUL((float)r.dp_max2, (float)subsc) / (float)r.dp_max / (float)r.score0;
                const float one_m = TELR_FADD(1.0f, -TELR_FMUL(x, x));
                mq = (int)TELR_FMUL(TELR_FMUL(TELR_FMUL(TELR_FMUL(identity, pen_cm), q_coef), one_m), logf_cr((float)r.dp_max / (float)match_sc));
                const int mq_alt = (int)TELR_FADD(TELR_FMUL(TELR_FMUL(TELR_FMUL(6.02f, identity), identity), (float)(r.dp_max - r.dp_max2)) / (float)match_sc, .499f);
                mq = mq < mq_alt ? mq : mq_alt;
            } else {
                const float x = (float)subsc / (float)r.score0;
                const float one_m = TELR_FADD(1.0f, -x);
                if (r.has_p) {
                    const float identity = (float)r.mlen / (float)r.blen;
                    mq = (int)TELR_FMUL(TELR_FMUL(TELR_FMUL(TELR_FMUL(identity, pen_cm), q_coef), one_m), logf_cr((float)r.dp_max / (float)match_sc));
                } else mq = (int)TELR_FMUL(TELR_FMUL(TELR_FMUL(pen_cm, q_coef), one_m), logf_cr((float)r.score));
            }
            mq -= (int)TELR_FADD(TELR_FMUL(4.343f, logf_cr((float)(r.n_sub + 1))), .499f);
            mq = mq > 0 ? mq : 0;
            r.mapq = mq < 60 ? mq : 60;
            if (r.has_p && r.dp_max > r.dp_max2 && r.mapq == 0) r.mapq = 1;
        } else r.mapq = 0;
    }
    if (n_regs >= 3) {          // an inversion piece takes the smaller MAPQ of its neighbours on the target
        int i;
        for (i = 0; i < n_regs; ++i) if (regs[i].inv) break;
        if (i < n_regs) {
            Anchor *aux = s.z;
            int n_aux = 0;
            for (i = 0; i < n_regs; ++i)
                if (regs[i].parent == i || regs[i].parent < 0) aux[n_aux].y = (uint64_t)i, aux[n_aux++].x = (uint64_t)(uint32_t)regs[i].rs;
            rs_sort_emul(aux, n_aux, KeyX(), s.sortws);
            for (i = 1; i < n_aux - 1; ++i)
                if (regs[aux[i].y].inv) {
                    const int l = regs[aux[i - 1].y].mapq, rr = regs[aux[i + 1].y].mapq;
                    regs[aux[i].y].mapq = l < rr ? l : rr;
                }
        }
    }
}

// squeeze anchors not referenced by any region; returns number of anchors kept
TELR_HDN int regs_squeeze_anchors(int n_regs, Reg *regs, Anchor *a, HitScratch &s)
{
    int as = 0;
    uint64_t *aux = s.cov;
    for (int i = 0; i < n_regs; ++i) aux[i] = (uint64_t)(uint32_t)regs[i].as << 32 | (uint32_t)i;
    rs_sort_emul(aux, n_regs, KeyId(), s.sortws);
    for (int i = 0; i < n_regs; ++i) {
        Reg &r = regs[(int32_t)aux[i]];
        if (r.as != as) {
            for (int c = 0; c < r.cnt; ++c) a[as + c] = a[r.as + c];
            r.as = as;
        }
        as += r.cnt;
    }
    return as;
}

TELR_HD void reg_split(Reg &r, Reg &r2, int n, int qlen, const Anchor *a)
{
    if (n <= 0 || n >= r.cnt) return;
    r2 = r;
    r2.id = -1;
    r2.sam_pri = 0;
    r2.has_p = 0, r2.n_cigar = 0, r2.cig = 0, r2.dp_score = r2.dp_max = r2.dp_max2 = r2.n_ambi = 0;
    r2.split_inv = 0;
    r2.cnt = r.cnt - n;
    r2.score = (int32_t)(r.score * ((float)r2.cnt / r.cnt) + .499);
    r2.as = r.as + n;
    if (r.parent == r.id) r2.parent = PARENT_TMP_PRI;
    reg_set_coor(r2, qlen, a);
    r.cnt -= r2.cnt;
    r.score -= r2.score;
    reg_set_coor(r, qlen, a);
    r.split |= 1, r2.split |= 2;
}

}  // namespace telr
