// tests/emu/emu.cpp — HOST harness for the product's sequential control logic (telr_b200/csrc/mm_*.cuh).
// TEST INFRASTRUCTURE ONLY: it drives the very same chaining / region / alignment-state-machine code the
// CUDA kernels run on one lane, with the warp-parallel primitives (candidate scan, DP) replaced by
// sequential stand-ins (chain_dp_seq / chain_rmq_seq from the product headers, the oracle's DP).
// It lets `pytest -m "not gpu"` check that logic against the oracle on a CPU.  Not shipped, not a fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../telr_b200/csrc/mm_align.cuh"
extern "C" {
#include "../../oracle/orc.h"
}
using namespace telr;

static void opt_from_orc(Opt &o, const orc_opt_t &p)
{
    memset(&o, 0, sizeof(o));
    o.k = p.k; o.w = p.w; o.hpc = p.hpc; o.a = p.a; o.b = p.b; o.q = p.q; o.e = p.e; o.q2 = p.q2; o.e2 = p.e2;
    o.sc_ambi = p.sc_ambi; o.zdrop = p.zdrop; o.zdrop_inv = p.zdrop_inv; o.end_bonus = p.end_bonus;
    o.min_dp_max = p.min_dp_max; o.min_ksw_len = p.min_ksw_len; o.bw = p.bw; o.bw_long = p.bw_long; o.max_gap = p.max_gap;
    o.max_chain_skip = p.max_chain_skip; o.max_chain_iter = p.max_chain_iter; o.min_cnt = p.min_cnt;
    o.min_chain_score = p.min_chain_score; o.rmq_inner_dist = p.rmq_inner_dist; o.rmq_size_cap = p.rmq_size_cap;
    o.rmq_rescue_size = p.rmq_rescue_size; o.rmq_rescue_ratio = p.rmq_rescue_ratio;
    o.chn_pen_gap = (float)(p.chain_gap_scale * 0.01 * p.k); o.chn_pen_skip = (float)(p.chain_skip_scale * 0.01 * p.k);
    o.mask_level = p.mask_level; o.mask_len = p.mask_len; o.pri_ratio = p.pri_ratio; o.best_n = p.best_n;
    o.q_occ_frac = p.q_occ_frac; o.mid_occ_frac = p.mid_occ_frac; o.min_mid_occ = p.min_mid_occ; o.max_mid_occ = p.max_mid_occ; o.max_max_occ = p.max_max_occ; o.occ_dist = p.occ_dist;
    o.seed_term = wang_hash32((uint32_t)p.seed); o.max_sw_mat = p.max_sw_mat; o.rank_min_len = p.rank_min_len;
    o.rank_frac = p.rank_frac; o.max_clip_ratio = p.max_clip_ratio;
}

// run the product logic for one read x contig strand starting from SORTED anchors.
// out regs: 13 ints each (rs,re,qs,qe,rev,flag,dp_max,mlen,blen,n_cigar,cig_off,score,cnt); returns n regs or <0
extern "C" int emu_map_from_anchors(int preset, const uint8_t *contig, int clen, const uint8_t *read, int qlen,
                                    uint32_t name_hash, int64_t n_a, const uint64_t *anchors /* x,y pairs */,
                                    int32_t *out_regs, int out_cap, uint32_t *out_cigar, int64_t cigar_cap,
                                    int64_t *n_cigar_out, int32_t *chain_dump /* n_u, then u pairs hi/lo */, int chain_cap)
{
    orc_opt_t po; orc_opt_preset(&po, preset);
    Opt o; opt_from_orc(o, po);
    int n = (int)n_a;
    std::vector<Anchor> a(n + 1);
    for (int i = 0; i < n; ++i) a[i].x = anchors[2 * i], a[i].y = anchors[2 * i + 1];
    std::vector<uint8_t> cs(chain_scratch_bytes(n + 1));
    ChainScratch s; chain_scratch_carve(s, cs.data(), n + 1);
    int n_u = 0, n_v = 0;
    *n_cigar_out = 0;
    if (n == 0) return 0;
    chain_dp_seq(o, n, a.data(), s);
    chain_backtrack(n, s, o.min_cnt, o.min_chain_score, o.bw, &n_u, &n_v);
    if (n_u == 0) return 0;
    chain_compact(n_u, n_v, s, a.data());
    if (o.bw_long > o.bw && n_u > 1) {
        int32_t st = (int32_t)a[0].y, en = (int32_t)a[(int32_t)s.u[0] - 1].y;
        if (qlen - (en - st) > o.rmq_rescue_size || en - st > qlen * o.rmq_rescue_ratio) {
            int m = 0;
            for (int i = 0; i < n_u; ++i) m += (int32_t)s.u[i];
            rs_sort_emul(a.data(), m, KeyX(), s.sortws);
            chain_rmq_seq(o, m, a.data(), s);
            chain_backtrack(m, s, o.min_cnt, o.min_chain_score, o.bw_long, &n_u, &n_v);
            if (n_u == 0) return 0;
            chain_compact(n_u, n_v, s, a.data());
        }
    }
    if (chain_dump) {
        chain_dump[0] = n_u;
        for (int i = 0; i < n_u && 1 + 2 * i + 1 < chain_cap; ++i) chain_dump[1 + 2 * i] = (int32_t)(s.u[i] >> 32), chain_dump[2 + 2 * i] = (int32_t)s.u[i];
    }
    int cap_regs = 2 * (n / 3) + 4;
    std::vector<Reg> regs(cap_regs + 1);
    std::vector<uint8_t> hsb(hit_scratch_bytes(cap_regs + 1));
    AlnCtx c; memset(&c, 0, sizeof(c));
    hit_scratch_carve(c.hs, hsb.data(), cap_regs + 1);
    uint32_t hash = name_hash;
    hash ^= wang_hash32((uint32_t)qlen) + o.seed_term;
    hash = wang_hash32(hash);
    regs_from_chains(hash, qlen, n_u, s.u, a.data(), regs.data(), c.hs);
    int n_regs = n_u;
    regs_set_parent(o, n_regs, regs.data(), c.hs);
    regs_select_sub(o, 1, &n_regs, regs.data(), c.hs, cap_regs);
    if (n_regs == 0) return 0;
    // alignment
    std::vector<uint8_t> qrc(qlen);
    for (int i = 0; i < qlen; ++i) qrc[qlen - 1 - i] = read[i] < 4 ? 3 - read[i] : 4;
    std::vector<uint32_t> cig((size_t)4 * (qlen + clen) + 1024);
    std::vector<int32_t> K(n + 4);
    c.o = &o; c.tseq = contig; c.tlen = clen; c.qseq[0] = read; c.qseq[1] = qrc.data(); c.qlen = qlen;
    c.a = a.data(); c.regs = regs.data(); c.n_regs = n_regs; c.cap_regs = cap_regs;
    c.cig = cig.data(); c.cig_cap = (uint32_t)cig.size(); c.K = K.data(); c.capK = n + 4;
    c.phase = PH_START;
    DpRes res; res_reset(res);
    DpTask t;
    orc_ez_t ez; memset(&ez, 0, sizeof(ez));
    std::vector<uint8_t> qb, tb;
    while (aln_next(c, res, t)) {
        qb.resize(t.qlen); tb.resize(t.tlen);
        for (int i = 0; i < t.qlen; ++i) { uint8_t b = t.q[(ptrdiff_t)i * t.qstep]; qb[i] = t.qcomp ? (b >= 4 ? 4 : 3 - b) : b; }
        for (int i = 0; i < t.tlen; ++i) tb[i] = t.t[(ptrdiff_t)i * t.tstep];
        res_reset(res);
        if (t.kind == 0) {
            orc_ksw_extd2(t.qlen, qb.data(), t.tlen, tb.data(), o.a, o.b, o.sc_ambi, o.q, o.e, o.q2, o.e2, t.w, t.zdrop, t.end_bonus, t.flag, &ez);
            res.max = ez.max; res.max_q = ez.max_q; res.max_t = ez.max_t; res.mqe = ez.mqe; res.mqe_t = ez.mqe_t;
            res.mte = ez.mte; res.mte_q = ez.mte_q; res.score = ez.score; res.zdropped = ez.zdropped; res.reach_end = ez.reach_end;
            res.n_cigar = ez.n_cigar; res.cigar = ez.cigar;
        } else {
            int qe, te;
            res.ll_score = orc_ksw_ll(t.qlen, qb.data(), t.tlen, tb.data(), o.a, o.b, o.sc_ambi, o.q, o.e, &qe, &te);
            res.ll_qe = qe; res.ll_te = te;
            res.n_cigar = ez.n_cigar; res.cigar = ez.cigar;     // untouched
        }
    }
    int ret = c.err ? -100 - c.err : 0;
    if (!ret) {
        int64_t nc = 0;
        for (int i = 0; i < c.n_regs; ++i) {
            const Reg &r = regs[i];
            if (i >= out_cap || nc + r.n_cigar > cigar_cap) { ret = -1; break; }
            int32_t *o2 = out_regs + 13 * i;
            o2[0] = r.rs; o2[1] = r.re; o2[2] = r.qs; o2[3] = r.qe; o2[4] = r.rev;
            o2[5] = (r.rev ? 0x10 : 0) | (r.parent != r.id ? 0x100 : !r.sam_pri ? 0x800 : 0);
            o2[6] = r.dp_max; o2[7] = r.mlen; o2[8] = r.blen; o2[9] = r.n_cigar; o2[10] = (int32_t)nc; o2[11] = r.score; o2[12] = r.cnt;
            memcpy(out_cigar + nc, c.cig + r.cig, (size_t)r.n_cigar * 4);
            nc += r.n_cigar;
        }
        *n_cigar_out = nc;
        if (!ret) ret = c.n_regs;
    }
    free(ez.cigar);
    return ret;
}

extern "C" void emu_radix_sort_128x(uint64_t *pairs, int n)
{
    std::vector<int32_t> ws(rs_scratch_words(n) + 8);
    rs_sort_emul((Anchor *)pairs, n, KeyX(), ws.data());
}
