"""Oracle self-consistency and the host emulation of the product's control logic (CPU only)."""
import ctypes as C

import numpy as np
import pytest

from telr_b200 import synth
from tests import orc, util


def rand_seq(rng, n, p_n=0.0):
    s = rng.integers(0, 4, n).astype(np.uint8)
    if p_n:
        s[rng.random(n) < p_n] = 4
    return s


def revcomp(s):
    return np.where(s[::-1] < 4, 3 - s[::-1], 4).astype(np.uint8)


@pytest.mark.parametrize("w,k", [(10, 15), (19, 19), (5, 11)])
def test_sketch_strand_symmetry(built, w, k):
    """The minimizer multiset of a sequence and of its reverse complement are mirror images."""
    rng = np.random.default_rng(w * 100 + k)
    for n in (k, k + w - 1, 200, 5000):
        s = rand_seq(rng, n)
        x1, y1 = orc.sketch(s, w, k)
        x2, y2 = orc.sketch(revcomp(s), w, k)
        assert sorted(x1.tolist()) == sorted(x2.tolist())
        p1 = sorted(((y >> 1), int(x)) for x, y in zip(x1.tolist(), y1.tolist()))
        p2 = sorted((n - 1 - (y >> 1) + (k - 1), int(x)) for x, y in zip(x2.tolist(), y2.tolist()))
        assert p1 == p2


def test_sketch_short_and_ambiguous(built):
    rng = np.random.default_rng(3)
    assert len(orc.sketch(rand_seq(rng, 10), 10, 15)[0]) == 0             # shorter than k
    assert len(orc.sketch(rand_seq(rng, 20), 10, 15)[0]) == 1             # shorter than w+k-1: one minimizer at the end flush
    s = rand_seq(rng, 400)
    s[::7] = 4                                                            # no 15-mer without N
    assert len(orc.sketch(s, 10, 15)[0]) == 0
    x, y = orc.sketch(rand_seq(rng, 3000, 0.002), 10, 15)
    assert (np.diff((y >> np.uint64(1)).astype(np.int64)) > 0).all()       # emitted in position order


def test_sketch_hpc_spans(built):
    rng = np.random.default_rng(4)
    s = np.repeat(rand_seq(rng, 600), rng.integers(1, 5, 600))
    x, y = orc.sketch(s, 10, 19, 1)
    assert len(x) > 0 and ((x & np.uint64(0xff)) >= 19).all() and ((x & np.uint64(0xff)) < 256).all()


def score_cigar(q, t, cig, o):
    i = j = sc = 0
    for c in cig.tolist():
        op, ln = c & 0xf, c >> 4
        if op == 0:
            for z in range(ln):
                a, b = int(t[i + z]), int(q[j + z])
                sc += -o.sc_ambi if (a > 3 or b > 3) else (o.a if a == b else -o.b)
            i += ln; j += ln
        elif op == 1:
            sc -= min(o.q + o.e * ln, o.q2 + o.e2 * ln); j += ln
        else:
            sc -= min(o.q + o.e * ln, o.q2 + o.e2 * ln); i += ln
    return sc, i, j


def test_ksw_global_score_equals_cigar_score(built):
    rng = np.random.default_rng(11)
    o = orc.opt(0)
    for n in (1, 7, 60, 300):
        t = rand_seq(rng, n)
        q = t.copy()
        for _ in range(max(1, n // 12)):
            p = rng.integers(0, len(q))
            r = rng.random()
            q = np.delete(q, p) if (r < .3 and len(q) > 2) else (np.insert(q, p, rng.integers(0, 4)) if r < .6 else q)
            if r >= .6:
                q[p] = (q[p] + 1) % 4
        for flag in (0, orc_flag("APPROX")):
            r = orc.ksw_extd2(q, t, o, -1, -1, -1, flag)
            sc, i, j = score_cigar(q, t, r["cigar"], o)
            assert (i, j) == (len(t), len(q))
            assert sc == r["score"], (n, flag, sc, r["score"])
    same = rand_seq(rng, 100)
    r = orc.ksw_extd2(same, same, o, -1, -1, -1, 0)
    assert r["cigar"].tolist() == [100 << 4] and r["score"] == 100 * o.a


def orc_flag(name):
    return {"APPROX": 0x08, "EXTZ": 0x40, "RIGHT": 0x02, "REV": 0x80}[name]


def test_ksw_extension_zdrop_and_cells(built):
    rng = np.random.default_rng(12)
    o = orc.opt(0)
    t = rand_seq(rng, 3000)
    q = np.concatenate([t[:300], rand_seq(rng, 2000)])          # homology ends after 300 bases
    r = orc.ksw_extd2(q, t, o, 751, 400, -1, orc_flag("EXTZ"))
    assert r["zdropped"] == 1 and 280 <= r["max_t"] <= 320 and r["cells"] < 2000 * 3000 // 4
    sc, i, j = score_cigar(q, t, r["cigar"], o)
    assert (i, j) == (r["max_t"] + 1, r["max_q"] + 1) and sc == r["max"]


def test_ksw_ll_matches_bruteforce(built):
    rng = np.random.default_rng(13)
    o = orc.opt(0)
    for _ in range(10):
        q, t = rand_seq(rng, int(rng.integers(1, 40))), rand_seq(rng, int(rng.integers(1, 40)))
        qe, te = C.c_int(), C.c_int()
        sc = orc.lib().orc_ksw_ll(len(q), q.ctypes.data, len(t), t.ctypes.data, o.a, o.b, o.sc_ambi, o.q, o.e, C.byref(qe), C.byref(te))
        H = np.zeros((len(t) + 1, len(q) + 1), int); E = np.zeros_like(H); F = np.zeros_like(H)
        best = 0
        for i in range(1, len(t) + 1):
            for j in range(1, len(q) + 1):
                E[i, j] = max(E[i - 1, j] - o.e, H[i - 1, j] - o.q - o.e, 0)
                F[i, j] = max(F[i, j - 1] - o.e, H[i, j - 1] - o.q - o.e, 0)
                s = o.a if q[j - 1] == t[i - 1] else -o.b
                H[i, j] = max(0, H[i - 1, j - 1] + s, E[i, j], F[i, j])
                best = max(best, H[i, j])
        assert sc == best


def test_radix_sort_emulation_matches_oracle(built):
    rng = np.random.default_rng(1)
    for n in (2, 64, 65, 300, 5000):
        for trial in range(6):
            x = rng.integers(0, 1 << 20 if trial < 3 else 40, n).astype(np.uint64) << np.uint64(int(rng.integers(0, 40)))
            a = np.stack([x, np.arange(n, dtype=np.uint64)], 1).copy()
            b = a.copy()
            orc.lib().orc_radix_sort_128x(a.ctypes.data, a.ctypes.data + 16 * n)
            util.emu().emu_radix_sort_128x(b.ctypes.data, n)
            assert (a == b).all()
            assert (np.diff(a[:, 0].astype(np.int64)) >= 0).all()


@pytest.mark.parametrize("cfg,preset", [("ont_3k_50x", 0), ("clr_3k_40x", 1), ("hifi_3k_40x", 2)])
def test_product_logic_emulated_on_host_matches_oracle(built, cfg, preset):
    """telr_b200/csrc/mm_*.cuh (chaining, regions, alignment coroutine) driven on the CPU vs the oracle."""
    b = synth.generate(cfg, 0, 2, depth=12)
    o = orc.opt(preset)
    E = util.emu()
    n_prob = 0
    for l in range(b.n_loci):
        ctg = b.unpack(int(b.contig_off[l]), int(b.contig_len[l]))
        for cs in (ctg, (3 - ctg[::-1]).astype(np.uint8)):
            for r in range(b.locus_read_begin[l], b.locus_read_begin[l + 1]):
                rd = b.unpack(int(b.read_off[r]), int(b.read_len[r]))
                ref = orc.map_one(o, cs, rd, int(b.read_hash[r]))
                anch = np.ascontiguousarray(ref["anchors"])
                regs = np.zeros((64, 13), np.int32); cig = np.zeros(len(rd) * 4 + 1024, np.uint32); nc = C.c_int64(0)
                n = E.emu_map_from_anchors(preset, cs.ctypes.data, len(cs), rd.ctypes.data, len(rd), int(b.read_hash[r]), len(anch),
                                           anch.ctypes.data, regs.ctypes.data, 64, cig.ctypes.data, len(cig), C.byref(nc), None, 0)
                ra = ref["aln"]
                assert n == len(ra)
                for i in range(n):
                    e, a = regs[i], ra[i]
                    assert tuple(e[:10]) == (a["rs"], a["re"], a["qs"], a["qe"], a["rev"], a["flag"], a["dp_max"], a["mlen"], a["blen"], a["n_cigar"])
                    assert (cig[e[10]:e[10] + e[9]] == ref["cigar"][a["cigar_off"]:a["cigar_off"] + a["n_cigar"]]).all()
                n_prob += 1
    assert n_prob > 20


def test_oracle_alignment_invariants(built):
    b = synth.generate("ont_3k_50x", 0, 2, depth=15)
    r = orc.af_run(b, threads=0)
    al = r.alns
    depth_from_blocks = np.zeros_like(r.depth)
    off = np.concatenate([[0], np.cumsum(2 * b.contig_len.astype(np.int64))])
    locus_of_read = np.repeat(np.arange(b.n_loci), np.diff(b.locus_read_begin))
    for i, a in enumerate(al):
        c = r.cigar_of(i); op, ln = c & 0xf, c >> 4
        assert op[0] == 0 and op[-1] == 0
        assert ln[(op == 0) | (op == 2)].sum() == a["re"] - a["rs"] and ln[(op == 0) | (op == 1)].sum() == a["qe"] - a["qs"]
        if a["flag"] & 0x100:
            continue
        l = locus_of_read[a["read"]]; pos = a["rs"]
        base = off[l] + a["strand"] * b.contig_len[l]
        for o_, n_ in zip(op.tolist(), ln.tolist()):
            if o_ == 0:
                depth_from_blocks[base + pos: base + pos + n_] += 1; pos += n_
            elif o_ == 2:
                pos += n_
    assert (depth_from_blocks == r.depth).all()          # depth == sum of M blocks of non-secondary records


def test_config1_fixture_matches_committed_oracle_outputs(built):
    import hashlib, json, os
    b = util.load_config1()
    gold = json.load(open(os.path.join(util.ROOT, "tests", "golden", "config1_oracle.json")))
    r = orc.af_run(b, threads=0)
    assert r.cov2x.tolist() == gold["cov2x"] and int(r.c.dp_cells) == gold["dp_cells"] and int(r.c.n_aln) == gold["n_aln"]
    assert hashlib.sha1(r.depth.tobytes()).hexdigest() == gold["depth_sha1"]
    assert [[int(a[f]) for f in ("read", "strand", "rs", "re", "qs", "qe", "rev", "flag", "dp_max", "n_cigar")] for a in r.alns] == gold["aln"]


def textbook_two_piece_global(q, t, o):
    """Independent O(nm) global alignment score under the two-piece affine gap cost min(q + l*e, q2 + l*e2), written from the
    textbook recurrence with explicit H/E/F/E2/F2 matrices -- not from the difference formulation of orc_ksw.c."""
    n, m = len(t), len(q)
    NEG = -10 ** 9
    gap = lambda l: 0 if l == 0 else -min(o.q + o.e * l, o.q2 + o.e2 * l)
    Hp = [gap(j) for j in range(m + 1)]
    E1 = [NEG] * (m + 1); E2 = [NEG] * (m + 1)          # gaps that consume target bases (vertical), per query column
    for i in range(1, n + 1):
        H = [gap(i)] + [NEG] * m
        F1 = F2 = NEG                                   # gaps that consume query bases (horizontal)
        ti = int(t[i - 1])
        for j in range(1, m + 1):
            E1[j] = max(Hp[j] - o.q, E1[j]) - o.e
            E2[j] = max(Hp[j] - o.q2, E2[j]) - o.e2
            F1 = max(H[j - 1] - o.q, F1) - o.e
            F2 = max(H[j - 1] - o.q2, F2) - o.e2
            qj = int(q[j - 1])
            s = -o.sc_ambi if (ti > 3 or qj > 3) else (o.a if ti == qj else -o.b)
            H[j] = max(Hp[j - 1] + s, E1[j], E2[j], F1, F2)
        Hp = H
    return Hp[m]


@pytest.mark.parametrize("preset", [0, 2])
def test_ksw_global_score_is_optimal(built, preset):
    """The oracle's global (gap-fill) score equals the optimum of an independently written dynamic programme, for the exact and
    the approximate-max variants, and its CIGAR rescoring reaches that optimum (so the traceback is an optimal path)."""
    rng = np.random.default_rng(70 + preset)
    o = orc.opt(preset)
    for trial in range(14):
        n = int(rng.choice([1, 2, 9, 40, 90, 160]))
        t = rand_seq(rng, n, p_n=0.02 if trial % 5 == 0 else 0.0)
        q = t.copy()
        for _ in range(int(rng.integers(0, max(2, n // 6)))):
            p = int(rng.integers(0, len(q)))
            r = rng.random()
            if r < .25 and len(q) > 2:
                q = np.delete(q, slice(p, p + int(rng.integers(1, 30))))           # long deletions exercise the second gap piece
            elif r < .5:
                q = np.insert(q, p, rng.integers(0, 4, int(rng.integers(1, 30))))
            else:
                q[p] = (q[p] + 1) % 4
        if len(q) == 0:
            q = rand_seq(rng, 1)
        best = textbook_two_piece_global(q, t, o)
        for flag in (0, orc_flag("APPROX")):
            r = orc.ksw_extd2(q, t, o, -1, -1, -1, flag)
            sc, i, j = score_cigar(q, t, r["cigar"], o)
            assert (i, j) == (len(t), len(q))
            assert r["score"] == best == sc, (trial, n, len(q), flag, r["score"], best, sc)


def test_deviation_reach_counters(built):
    """How often the five configurations reach a spot where the restatement knowingly differs from minimap2 2.22
    (DESIGN.md section 3): plain mid_occ filter, RMQ priority ties, band-edge paths, ksw_ll ties, index buckets > 64."""
    from telr_b200 import synth
    out = (C.c_int64 * 8)()
    tot = np.zeros(8, np.int64)
    for cfg, first, n in (("ont_3k_50x", 0, 12), ("clr_3k_40x", 0, 8), ("hifi_3k_40x", 0, 8), ("ont_30k_30x", 500, 16), ("poly_10k_200x", 0, 2)):
        orc.lib().orc_dev_counters(out, 1)
        orc.af_run(synth.generate(cfg, first, n), threads=0, want_depth=False, want_aln=False)
        orc.lib().orc_dev_counters(out, 1)
        tot += np.array(list(out), np.int64)
    assert tot[5] > 1000 and tot[6] > 10000                    # the counted paths did run: banded DP calls, RMQ queries
    assert tot[0] == 0 and tot[2] == 0 and tot[3] == 0 and tot[4] == 0, tot.tolist()
    assert tot[1] <= 1e-5 * tot[6], tot.tolist()               # measured: 7 ties in 5.8 M queries over 530 loci (profiles/README.md)


def test_bounded_extension_changes_no_output(built):
    """The bounded-extension rule (stop an extension once no later anti-diagonal can beat the maximum found so far) saves rows,
    never results: with the rule switched off -- every anti-diagonal computed, as ksw2 does -- the oracle gives the same alignment
    records, CIGARs, depth and coverage integers; only the cell count differs."""
    from telr_b200 import synth
    from tests import util
    saved = []
    try:
        for cfg, first, n in (("ont_3k_50x", 30, 6), ("clr_3k_40x", 5, 4), ("hifi_3k_40x", 5, 4)):
            b = synth.generate(cfg, first, n)
            orc.set_ext_bound(False)
            r0 = orc.af_run(b, threads=0)
            orc.set_ext_bound(True)
            r1 = orc.af_run(b, threads=0)
            util.assert_same_results(r1, r0)
            assert r1.c.dp_cells < r0.c.dp_cells
            saved.append(1 - int(r1.c.dp_cells) / int(r0.c.dp_cells))
        # single calls: an overhanging query against a short target, with and without the rule
        rng = np.random.default_rng(5)
        o = orc.opt(0)
        for tl in (8, 30, 120, 400):
            t = rand_seq(rng, tl)
            q = np.concatenate([t[: tl * 3 // 4], rand_seq(rng, 3000)])
            for flag in (orc_flag("EXTZ"), orc_flag("EXTZ") | orc_flag("RIGHT") | orc_flag("REV")):
                orc.set_ext_bound(False)
                a = orc.ksw_extd2(q, t, o, 751, 400, -1, flag)
                orc.set_ext_bound(True)
                bnd = orc.ksw_extd2(q, t, o, 751, 400, -1, flag)
                assert all(a[k] == bnd[k] for k in ("max", "max_q", "max_t", "reach_end")) and (a["cigar"] == bnd["cigar"]).all()
                assert bnd["cells"] < a["cells"] * (0.6 if tl <= 120 else 1.0)
    finally:
        orc.set_ext_bound(True)
    assert min(saved) > 0.005
