"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bit-exact for alignment coordinates/CIGARs, per-base depth, coverage integers, DP cell counts; AF within 1e-9."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from telr_b200 import lib, stage4, synth
from telr_b200.batch import pack_sequences
from tests import orc, util
from tests.test_host import _make_stage3_artifacts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = lib.Context(0)
    yield c
    c.close()


def _seqs(rng, hpc, with_n):
    seqs = []
    for ln in (1, 5, 14, 15, 24, 25, 26, 100, 2047, 2048, 2049, 4096, 4100, 10000, 33333):
        s = rng.integers(0, 4, ln).astype(np.uint8)
        if hpc:
            for _ in range(ln // 50):
                p = rng.integers(0, ln); s[p:p + rng.integers(2, 12)] = s[p]
        if ln > 300:
            s[200:290] = np.tile(s[200:207], 13)[:90]           # tandem array: identical hashes inside windows
        b = np.array(list(b"ACGT"), np.uint8)[s]
        if with_n and ln > 30:
            for _ in range(max(1, ln // 400)):
                b[rng.integers(0, ln)] = ord("N")
        seqs.append(bytes(b))
    return seqs


@pytest.mark.parametrize("w,k,hpc", [(10, 15, 0), (10, 19, 1), (19, 19, 0)])
@pytest.mark.parametrize("with_n", [False, True])
def test_sketch_matches_oracle(ctx, w, k, hpc, with_n):
    rng = np.random.default_rng(7 + w + k + hpc)
    seqs = _seqs(rng, hpc, with_n)
    seq2, nmask, offs, lens = pack_sequences(seqs)
    x, y, off = ctx.sketch(seq2, nmask, offs, lens, w, k, hpc)
    for i, s in enumerate(seqs):
        nt4 = np.array([{65: 0, 67: 1, 71: 2, 84: 3}.get(c, 4) for c in s], np.uint8)
        ox, oy = orc.sketch(nt4, w, k, hpc)
        assert (ox == x[off[i]:off[i + 1]]).all() and (oy == y[off[i]:off[i + 1]]).all(), (len(s), w, k, hpc)


def test_sketch_two_pass_kernel_agrees(built, monkeypatch):
    """TELR_SKETCH_TILES=0: the first-generation CTA-per-sequence two-pass kernel, kept as a second implementation."""
    monkeypatch.setenv("TELR_SKETCH_TILES", "0")
    c = lib.Context(0)
    try:
        rng = np.random.default_rng(11)
        for w, k, hpc in ((10, 15, 0), (10, 19, 1)):
            seqs = _seqs(rng, hpc, True)
            seq2, nmask, offs, lens = pack_sequences(seqs)
            x, y, off = c.sketch(seq2, nmask, offs, lens, w, k, hpc)
            for i, s in enumerate(seqs):
                nt4 = np.array([{65: 0, 67: 1, 71: 2, 84: 3}.get(ch, 4) for ch in s], np.uint8)
                ox, oy = orc.sketch(nt4, w, k, hpc)
                assert (ox == x[off[i]:off[i + 1]]).all() and (oy == y[off[i]:off[i + 1]]).all(), (len(s), hpc)
    finally:
        c.close()


def test_sketch_tile_boundaries_and_ambiguous_runs(ctx):
    """The tile kernel's seams: lengths around multiples of 2048, runs of N across a seam, N at both ends, all-N."""
    rng = np.random.default_rng(23)
    seqs = []
    for ln in (2047, 2048, 2049, 2048 + 14, 2048 + 15, 4095, 4096, 4097, 6144 + 25, 3 * 2048):
        b = np.array(list(b"ACGT"), np.uint8)[rng.integers(0, 4, ln)]
        seqs.append(bytes(b))
        b2 = b.copy(); b2[2040:2060] = ord("N"); seqs.append(bytes(b2))          # a run of N across the first seam
        b3 = b.copy(); b3[0] = b3[-1] = ord("N"); b3[min(2047, ln - 2)] = ord("N"); seqs.append(bytes(b3))
        b4 = b.copy(); b4[2048 - 24:2048 - 10] = b4[2048 - 38:2048 - 24]; seqs.append(bytes(b4))   # identical k-mers next to the seam
    seqs.append(b"N" * 5000)
    seqs.append(b"ACGT" * 1500)          # period-4 tandem: every window holds identical hashes
    seqs.append(b"A" * 300 + b"C" * 300 + b"ACGTTGCA" * 300)     # runs longer than a byte-sized span when compressed
    # every kernel instance: (32-bit key, w=10), (64-bit, w=19), both run-time-w forms, and the homopolymer-compressed forms
    for w, k, hpc in ((10, 15, 0), (19, 19, 0), (5, 11, 0), (10, 19, 0), (10, 19, 1), (5, 11, 1)):
        seq2, nmask, offs, lens = pack_sequences(seqs)
        x, y, off = ctx.sketch(seq2, nmask, offs, lens, w, k, hpc)
        for i, s in enumerate(seqs):
            nt4 = np.array([{65: 0, 67: 1, 71: 2, 84: 3}.get(ch, 4) for ch in s], np.uint8)
            ox, oy = orc.sketch(nt4, w, k, hpc)
            assert len(ox) == off[i + 1] - off[i] and (ox == x[off[i]:off[i + 1]]).all() and (oy == y[off[i]:off[i + 1]]).all(), (i, len(s), w, k, hpc)


def test_sketch_empty_batch(ctx):
    x, y, off = ctx.sketch(np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.int64), np.zeros(0, np.int32), 10, 15)
    assert len(x) == 0 and off.tolist() == [0]


@pytest.mark.parametrize("mode", ["0", "1"])
def test_depth_and_af_stage(built, mode, monkeypatch):
    monkeypatch.setenv("TELR_DEPTH_MODE", mode)          # 0: shared-memory atomics per base, 1: boundary marks + scan
    c = lib.Context(0)
    rng = np.random.default_rng(3)
    nl = 48
    clen = rng.integers(600, 9000, nl).astype(np.int32)
    ts = np.array([rng.integers(0, L - 150) for L in clen], np.int32)
    te = np.array([min(L, s + rng.integers(20, 4000)) for L, s in zip(clen, ts)], np.int32)
    ts[0] = 300; ts[1] = 299; te[2] = ts[2] + 60; ts[3] = -1
    bl, bs, bn = [], [], []
    for l in range(nl):
        for s in range(2):
            for _ in range(rng.integers(0, 120) if l != 5 else 0):
                st = rng.integers(0, clen[l]); ln = rng.integers(1, 3000)
                bl.append(2 * l + s); bs.append(st); bn.append(min(ln, clen[l] - st))
    depth, cov, af = c.depth_af(clen, ts, te, np.array(bl), np.array(bs), np.array(bn))
    off = 0
    for l in range(nl):
        L = int(clen[l]); d = [np.zeros(L, np.int32), np.zeros(L, np.int32)]
        for q, st, ln in zip(bl, bs, bn):
            if q >> 1 == l:
                d[q & 1][st:st + ln] += 1
        assert (depth[off:off + L] == d[0]).all() and (depth[off + L:off + 2 * L] == d[1]).all()
        off += 2 * L
        if ts[l] < 0:
            assert (cov[l] == -2).all() and np.isnan(af[l])
            continue
        c8 = np.zeros(8, np.int32); a = C.c_double()
        orc.lib().orc_cov_af(d[0].ctypes.data, d[1].ctypes.data, L, int(ts[l]), int(te[l]), 100, 200, 50, 50, c8.ctypes.data, C.byref(a))
        assert (c8 == cov[l]).all(), (l, c8, cov[l])
        assert (np.isnan(a.value) and np.isnan(af[l])) or a.value == af[l]
    c.close()


def _mut(rng, s, rate):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(ch); out.append(rng.integers(0, 4)); continue
        out.append((ch + 1 + rng.integers(0, 3)) % 4 if r < rate else ch)
    return np.array(out, np.uint8)


@pytest.mark.parametrize("preset", [0, 2])
def test_dp_stage_matches_oracle(ctx, preset):
    rng = np.random.default_rng(21 + preset)
    o = orc.opt(preset)
    tasks, qs, ts_ = [], [], []
    qo = to = 0
    spec = [(50, 0x08, 30001, 400, -1, .1), (300, 0x08, 30001, 400, -1, .12), (1000, 0x08, 30001, 400, -1, .12), (300, 0, 30001, 400, -1, .12),
            (700, 0x40, 751, 400, -1, .12), (700, 0xC2, 751, 400, -1, .12), (2500, 0x40, 751, 400, -1, .15), (2500, 0xC2, 751, 200, -1, .15),
            (33, 0x40, 751, 400, -1, .5), (1, 0x08, 30001, 400, -1, 0.), (400, 0x40, 751, 400, -1, .9), (3000, 0x40, 100, 400, -1, .1),
            (1500, 0x08, 1200, 400, -1, .1), (1300, 0x40, 40, 400, 10, .05)] * 3
    for (ql, flag, w, zd, eb, rate) in spec:
        t = rng.integers(0, 4, max(1, int(ql * rng.uniform(.7, 1.4)))).astype(np.uint8)
        q = _mut(rng, t, rate)[:ql] if rate < .8 else rng.integers(0, 4, ql).astype(np.uint8)
        if len(q) == 0:
            q = np.zeros(1, np.uint8)
        if ql == 300 and flag == 0:
            t = np.concatenate([t[:100], rng.integers(0, 4, 2500).astype(np.uint8), t[100:]])
        if rng.random() < .3 and len(q) > 40:
            q[rng.integers(0, len(q))] = 4
        tasks.append((qo, to, len(q), len(t), w, zd, eb, flag)); qs.append(q); ts_.append(t); qo += len(q); to += len(t)
    out, cig = ctx.dp(preset, np.array(tasks, lib.DPTASK_DTYPE), np.concatenate(qs), np.concatenate(ts_))
    for i, (q, t) in enumerate(zip(qs, ts_)):
        w, zd, eb, flag = tasks[i][4:]
        ref = orc.ksw_extd2(q, t, o, w, zd, eb, flag); g = out[i]
        names = ["zdropped", "reach_end", "cells"]
        if not flag & 0x08:
            names += ["max", "max_q", "max_t"]
        if not (flag & 0x40) and not ref["zdropped"]:
            names += ["score"]
        if flag & 0x40 and not ref["zdropped"]:
            names += ["mqe", "mqe_t"]
        assert all(int(ref[n]) == int(g[n]) for n in names), (i, {n: (int(ref[n]), int(g[n])) for n in names})
        gc = cig[g["cigar_off"]: g["cigar_off"] + g["n_cigar"]]
        assert len(gc) == len(ref["cigar"]) and (gc == ref["cigar"]).all(), i


@pytest.mark.parametrize("flag", [0x40, 0xC2, 0x00, 0x42, 0x08])
def test_dp_small_shapes_match_oracle(ctx, flag):
    """Every lane-group / window edge of the DP kernels: 1x1 up to band-limited shapes, both gap alignments,
    extension / global / approximate-max flags (ksw_extd2 flag bits, minimap2 ksw2.h)."""
    rng = np.random.default_rng(5 + flag)
    o = orc.opt(0)
    shapes = [(1, 1, 751), (1, 7, 751), (3, 5, 751), (4, 4, 751), (5, 9, 751), (8, 8, 751), (9, 9, 751), (9, 12, 751), (12, 9, 751),
              (13, 13, 751), (16, 16, 751), (9, 17, 751), (33, 40, 751), (64, 64, 751), (100, 130, 751), (129, 200, 751),
              (255, 256, 751), (256, 257, 751), (300, 350, 751), (700, 900, 751), (300, 300, 20), (900, 1000, 100), (40, 2000, 751), (2000, 40, 751)]
    tasks, qs, ts_ = [], [], []
    qo = to = 0
    for (ql, tl, w) in shapes:
        for rep in range(2):
            t = rng.integers(0, 4, tl).astype(np.uint8)
            q = _mut(rng, np.resize(t, max(ql * 2, 4)), .12)[:ql]
            if len(q) == 0:
                q = np.zeros(1, np.uint8)
            ww = -1 if flag == 0x08 else w
            tasks.append((qo, to, len(q), len(t), ww, 400, -1, flag)); qs.append(q); ts_.append(t); qo += len(q); to += len(t)
    out, cig = ctx.dp(0, np.array(tasks, lib.DPTASK_DTYPE), np.concatenate(qs), np.concatenate(ts_))
    for i, (q, t) in enumerate(zip(qs, ts_)):
        w, zd, eb, fl = tasks[i][4:]
        ref = orc.ksw_extd2(q, t, o, w, zd, eb, fl); g = out[i]
        names = ["zdropped", "cells"] + ([] if fl & 0x08 else ["max", "max_q", "max_t"])
        assert all(int(ref[n]) == int(g[n]) for n in names), (i, len(q), len(t), {n: (int(ref[n]), int(g[n])) for n in names})
        gc = cig[g["cigar_off"]: g["cigar_off"] + g["n_cigar"]]
        assert len(gc) == len(ref["cigar"]) and (gc == ref["cigar"]).all(), (i, len(q), len(t))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_dp_fuzz_matches_oracle(ctx, seed):
    """Random shapes, bands, z-drop thresholds, end bonuses and flag combinations through all three DP kernels
    (systolic fill, vectorised general DP, scalar fallback for ambiguous bases)."""
    rng = np.random.default_rng(1000 + seed)
    preset = [0, 1, 2][seed % 3]
    o = orc.opt(preset)
    tasks, qs, ts_ = [], [], []
    qo = to = 0
    for _ in range(160):
        tl = int(rng.choice([rng.integers(1, 40), rng.integers(40, 300), rng.integers(300, 900)]))
        rate = float(rng.choice([0.0, 0.05, 0.15, 0.4]))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = _mut(rng, np.resize(t, max(4, int(tl * rng.uniform(.5, 1.6)))), rate)
        q = q[: max(1, int(len(q) * rng.uniform(.6, 1.0)))]
        if rng.random() < .15 and len(q) > 8:
            q[rng.integers(0, len(q))] = 4
        if rng.random() < .1 and len(t) > 8:
            t[rng.integers(0, len(t))] = 4
        flag = int(rng.choice([0x08, 0x08, 0x00, 0x40, 0x42, 0xC2, 0x80, 0x48]))
        w = int(rng.choice([-1, 3, 17, 100, 751])) if flag != 0x08 else int(rng.choice([-1, 30001]))
        zd = int(rng.choice([-1, 40, 400])); eb = int(rng.choice([-1, 0, 10]))
        tasks.append((qo, to, len(q), len(t), w, zd, eb, flag)); qs.append(q); ts_.append(t); qo += len(q); to += len(t)
    out, cig = ctx.dp(preset, np.array(tasks, lib.DPTASK_DTYPE), np.concatenate(qs), np.concatenate(ts_))
    for i, (q, t) in enumerate(zip(qs, ts_)):
        w, zd, eb, flag = tasks[i][4:]
        ref = orc.ksw_extd2(q, t, o, w, zd, eb, flag); g = out[i]
        names = ["zdropped", "reach_end", "cells"]
        if not flag & 0x08:
            names += ["max", "max_q", "max_t"]
        if not (flag & 0x40) and not ref["zdropped"]:
            names += ["score"]
        if flag & 0x40 and not (flag & 0x08) and not ref["zdropped"]:
            names += ["mqe", "mqe_t"]
        assert all(int(ref[n]) == int(g[n]) for n in names), (i, tasks[i], {n: (int(ref[n]), int(g[n])) for n in names})
        gc = cig[g["cigar_off"]: g["cigar_off"] + g["n_cigar"]]
        assert len(gc) == len(ref["cigar"]) and (gc == ref["cigar"]).all(), (i, tasks[i])


@pytest.mark.parametrize("cfg,n,kw", [("ont_3k_50x", 10, {}), ("clr_3k_40x", 6, {}), ("hifi_3k_40x", 6, {}),
                                      ("poly_10k_200x", 2, dict(depth=60)), ("ont_3k_50x", 4, dict(p_n=0.002)),
                                      # wider samples of the configurations BASELINE.json names (rare paths: z-drop re-runs,
                                      # inversion probes, long joins, two-pass fills), config 5 at its full 200x depth
                                      ("ont_30k_30x", 40, {}), ("clr_3k_40x", 20, dict(first=100)), ("poly_10k_200x", 3, {})])
def test_full_pipeline_matches_oracle(ctx, cfg, n, kw):
    kw = dict(kw)
    b = synth.generate(cfg, kw.pop("first", 0), n, **kw)
    r = ctx.run(b, want_depth=True, want_aln=True)
    ro = orc.af_run(b, threads=0)
    util.assert_same_results(r, ro)
    assert r.c.dp_cells == ro.c.dp_cells and r.c.n_anchors == ro.c.n_anchors


@pytest.mark.parametrize("cfg,n,env", [("ont_3k_50x", 8, {"TELR_AL_QUEUE": "0"}), ("ont_3k_50x", 8, {"TELR_AL_QUEUE": "1", "TELR_AL_EXT8": "3", "TELR_AL_WIDE8": "1"}),
                                       ("clr_3k_40x", 5, {"TELR_AL_QUEUE": "1", "TELR_AL_EXT8": "5", "TELR_AL_WIDE8": "0"}),
                                       ("hifi_3k_40x", 5, {"TELR_AL_QUEUE": "1", "TELR_AL_EXT8": "0", "TELR_AL_WIDE8": "4"})])
def test_both_alignment_kernels_match_oracle(built, monkeypatch, cfg, n, env):
    """The alignment stage has two kernels and a per-preset default (k_al_queue — SM roles + device-wide task rings, every problem
    migrating between warps through global memory, the 12-column fill instance in its own role — for map-ont, k_al_fused for map-pb and
    map-hifi).  The rest of the suite runs the defaults; this test forces the other kernel of each preset and non-default role splits."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    b = synth.generate(cfg, 20, n)
    c = lib.Context(0)
    r = c.run(b, want_depth=True, want_aln=True)
    c.close()
    ro = orc.af_run(b, threads=0)
    util.assert_same_results(r, ro)
    assert r.c.dp_cells == ro.c.dp_cells


def test_config1_repo_fixture(ctx):
    """BASELINE.json configs[0]: the locus built from the reference's test FASTAs (map-pb)."""
    b = util.load_config1()
    gold = json.load(open(os.path.join(util.ROOT, "tests", "golden", "config1_oracle.json")))
    r = ctx.run(b, want_depth=True, want_aln=True)
    assert r.cov2x.tolist() == gold["cov2x"] and int(r.c.dp_cells) == gold["dp_cells"] and int(r.c.n_aln) == gold["n_aln"]
    assert hashlib.sha1(r.depth.tobytes()).hexdigest() == gold["depth_sha1"]
    assert [[int(a[f]) for f in ("read", "strand", "rs", "re", "qs", "qe", "rev", "flag", "dp_max", "n_cigar")] for a in r.alns] == gold["aln"]
    assert [None if np.isnan(v) else float(v) for v in r.af] == gold["af"]


def test_edge_cases_empty_and_ragged(ctx):
    b = synth.generate("ont_3k_50x", 0, 5, depth=8)
    # locus 1 loses its reads, locus 2 its contig, locus 3 its annotation
    keep = np.ones(b.n_reads, bool); keep[b.locus_read_begin[1]:b.locus_read_begin[2]] = False
    lrb = np.concatenate([[0], np.cumsum([int(keep[b.locus_read_begin[l]:b.locus_read_begin[l + 1]].sum()) for l in range(5)])]).astype(np.int32)
    from telr_b200.batch import Batch
    b2 = Batch(b.preset, b.seq2, b.nmask, b.read_off[keep].copy(), b.read_len[keep].copy(), b.read_hash[keep].copy(), lrb, b.contig_off.copy(),
               b.contig_len.copy(), b.te_start.copy(), b.te_end.copy())
    b2.contig_len[2] = 0
    b2.te_start[3] = -1
    r = ctx.run(b2, want_depth=True, want_aln=True)
    ro = orc.af_run(b2, threads=0)
    util.assert_same_results(r, ro)
    assert (r.cov2x[2] == -2).all() and (r.cov2x[3] == -2).all() and r.cov2x[1, 0] == 0
    empty = synth.generate("ont_3k_50x", 0, 1, depth=8).subset([])
    r0 = ctx.run(empty)
    assert r0.cov2x.shape == (0, 8)


def test_degenerate_reads_and_contigs(ctx):
    """Inputs the reference would still hand to minimap2/samtools: reads shorter than a k-mer, all-N reads, N every few
    bases, a read identical to the contig, N runs inside the contig, a TE annotation touching both contig ends."""
    from telr_b200.batch import Batch, pack_sequences, name_hash
    rng = np.random.default_rng(99)
    base = synth.generate("ont_3k_50x", 0, 3, depth=6)
    ACGT = np.frombuffer(b"ACGTN", np.uint8)
    seqs, read_len, lrb = [], [], [0]
    contigs = []
    for l in range(3):
        c = base.unpack(int(base.contig_off[l]), int(base.contig_len[l])).copy()
        if l == 1:
            c[500:520] = 4; c[1200] = 4                       # N run + single N in the contig
        contigs.append(c)
        reads = [base.unpack(int(base.read_off[r]), int(base.read_len[r])) for r in range(base.locus_read_begin[l], base.locus_read_begin[l + 1])]
        reads += [np.array([0], np.uint8), rng.integers(0, 4, 14).astype(np.uint8), rng.integers(0, 4, 15).astype(np.uint8),
                  np.full(300, 4, np.uint8), c.copy(), c[::-1].copy() ^ 3 if l != 1 else c[:40].copy()]
        nr = reads[0].copy(); nr[::7] = 4; reads.append(nr)          # an ambiguous base every 7: no k-mer survives
        nr = reads[1].copy(); nr[100:103] = 4; nr[900] = 4; reads.append(nr)
        seqs += [bytes(ACGT[r]) for r in reads]
        lrb.append(len(seqs))
    n_reads = len(seqs)
    seqs += [bytes(ACGT[c]) for c in contigs]
    seq2, nmask, offs, lens = pack_sequences(seqs)
    te_s = base.te_start.copy(); te_e = base.te_end.copy()
    te_s[2] = 0; te_e[2] = int(base.contig_len[2])               # TE covers the whole contig: every flank window is clamped
    b = Batch(base.preset, seq2, nmask, offs[:n_reads].copy(), lens[:n_reads].copy(),
              np.array([name_hash("r%d" % i) for i in range(n_reads)], np.uint32), np.array(lrb, np.int32),
              offs[n_reads:].copy(), lens[n_reads:].copy(), te_s, te_e)
    for k in ("flank_len", "flank_off", "te_len", "te_off"):
        if hasattr(base, k):
            setattr(b, k, getattr(base, k))
    r = ctx.run(b, want_depth=True, want_aln=True)
    ro = orc.af_run(b, threads=0)
    util.assert_same_results(r, ro)
    assert r.c.dp_cells == ro.c.dp_cells


def _batch_from_sequences(preset, loci):
    """loci: [(contig nt4 array, [read nt4 arrays], te_start, te_end)] -> Batch"""
    from telr_b200.batch import Batch, pack_sequences, name_hash
    ACGT = np.frombuffer(b"ACGTN", np.uint8)
    seqs, lrb, n_reads = [], [0], 0
    for c, reads, _, _ in loci:
        seqs += [bytes(ACGT[r]) for r in reads]
        n_reads += len(reads)
        lrb.append(n_reads)
    seqs += [bytes(ACGT[c]) for c, _, _, _ in loci]
    seq2, nmask, offs, lens = pack_sequences(seqs)
    return Batch(preset, seq2, nmask, offs[:n_reads].copy(), lens[:n_reads].copy(), np.array([name_hash("rr%d" % i) for i in range(n_reads)], np.uint32),
                 np.array(lrb, np.int32), offs[n_reads:].copy(), lens[n_reads:].copy(), np.array([l[2] for l in loci], np.int32), np.array([l[3] for l in loci], np.int32))


def _repeat_motif_batch(preset, seed=300):
    """Contigs (36-45 kb) that carry one k-mer with a very small hash ~55 times, 700-800 bp apart: that single key exceeds
    mid_occ on the index side (in a single-contig index mid_occ sits at the 0.9998 quantile of the key counts, so only the top
    one or two keys ever can), while a read sees it too rarely for the query-side filter: mm_seed_select and the rep_len term
    of the MAPQ run.  The five synthetic configurations never reach this path (tests/test_oracle_props.py counters)."""
    rng = np.random.default_rng(seed + preset)
    o = orc.opt(preset)
    cands = rng.integers(0, 4, (4000, o.k)).astype(np.uint8)
    hashes = [int(orc.sketch(c, o.w, o.k, 0)[0][0] >> 8) if len(orc.sketch(c, o.w, o.k, 0)[0]) else 1 << 62 for c in cands]
    motif = cands[int(np.argmin(hashes))]
    loci = []
    for n_copies, gap in (((55, 700), (60, 760)) if preset == 0 else ((75, 800), (80, 820))):
        parts = [rng.integers(0, 4, 1200).astype(np.uint8)]
        for _ in range(n_copies):
            parts += [motif, rng.integers(0, 4, gap + int(rng.integers(0, 40))).astype(np.uint8)]
        contig = np.concatenate(parts)
        reads = []
        for _ in range(8):
            ln = int(rng.integers(5000, 9000)); a = int(rng.integers(0, len(contig) - ln))
            rd = _mut(rng, contig[a:a + ln], 0.05 if preset == 0 else 0.002)
            reads.append(rd if rng.random() < .5 else (3 - rd[::-1]).astype(np.uint8))
        loci.append((contig, reads, 5000, len(contig) - 5000))
    return _batch_from_sequences(preset, loci)


@pytest.mark.parametrize("preset", [0, 2])
def test_high_occurrence_seeds_are_selected_like_upstream(ctx, preset):
    import ctypes as C
    b = _repeat_motif_batch(preset)
    out = (C.c_int64 * 8)()
    orc.lib().orc_dev_counters(out, 1)
    ro = orc.af_run(b, threads=0)
    orc.lib().orc_dev_counters(out, 1)
    assert out[0] > 50, list(out)                        # high-occurrence seeds were met on the index side
    r = ctx.run(b, want_depth=True, want_aln=True)
    util.assert_same_results(r, ro)
    assert r.c.dp_cells == ro.c.dp_cells and r.c.n_anchors == ro.c.n_anchors


def test_long_read_config_matches_oracle(ctx):
    b = synth.generate("ont_30k_30x", 0, 2, depth=5)
    r = ctx.run(b, want_depth=True, want_aln=True)
    ro = orc.af_run(b, threads=0)
    util.assert_same_results(r, ro)
    assert r.c.dp_cells == ro.c.dp_cells


def test_long_contig_uses_the_global_index(ctx):
    """A 36 kb contig has more minimizers than the shared-memory index holds (4096): the global-memory instance takes over."""
    b = synth.generate("ont_3k_50x", 0, 2, depth=6, te_min=30000, te_max=31000, te_median=30500)
    assert int(b.contig_len.max()) > 30000
    r = ctx.run(b, want_depth=True, want_aln=True)
    ro = orc.af_run(b, threads=0)
    util.assert_same_results(r, ro)
    assert r.c.dp_cells == ro.c.dp_cells and r.c.n_anchors == ro.c.n_anchors


def test_chunking_and_rerun_are_identical(built, monkeypatch):
    b = synth.generate("ont_3k_50x", 0, 12, depth=12)
    c1 = lib.Context(0)
    r1 = c1.run(b, want_depth=True, want_aln=True)
    r1b = c1.run(b, want_depth=True, want_aln=True)            # same ctx again: workspace reuse
    monkeypatch.setenv("TELR_CHUNK_MBASES", "1")               # force several chunks of loci
    c2 = lib.Context(0)
    r2 = c2.run(b, want_depth=True, want_aln=True)
    for r in (r1b, r2):
        util.assert_same_results(r, r1)
        assert r.c.dp_cells == r1.c.dp_cells
    c1.close(); c2.close()


@pytest.fixture(scope="module")
def full_config2(built):
    """BASELINE.json configs[1] at its full size (3000 loci, 1.8 Gbase) through the default chunking, run once per session."""
    b = synth.generate("ont_3k_50x", 0, 3000)
    c = lib.Context(0)
    r = c.run(b)
    c.close()
    return b, r


def _random_sample_vs_oracle(b, r, n, seed):
    """`n` random loci of a full-size GPU result against the oracle run on exactly those loci (coverage integers bit-exact,
    AF within 1e-9 relative): large-offset paths (64-bit arenas, 0.9-Gbase chunks) checked against the CPU restatement."""
    idx = np.sort(np.random.default_rng(seed).choice(b.n_loci, n, replace=False))
    ro = orc.af_run(b.subset(idx), threads=0, want_depth=False, want_aln=False)
    bad = np.nonzero((r.cov2x[idx] != ro.cov2x).any(1))[0]
    assert len(bad) == 0, (idx[bad][:5], r.cov2x[idx[bad]][:5], ro.cov2x[bad][:5])
    ga = r.af[idx]
    assert np.array_equal(np.isnan(ga), np.isnan(ro.af))
    ok = ~np.isnan(ro.af)
    assert np.all(np.abs(ga[ok] - ro.af[ok]) <= 1e-9 * np.abs(ro.af[ok]))


def test_full_size_one_chunk_equals_many(full_config2, monkeypatch):
    """Full config 2: the default chunking and 384-Mbase chunks give the same coverage integers, AF values and DP cell count
    (guards the 64-bit offsets of the large-chunk layout)."""
    b, r1 = full_config2
    monkeypatch.setenv("TELR_CHUNK_MBASES", "384")
    c2 = lib.Context(0)
    r2 = c2.run(b)
    c2.close()
    assert (r1.cov2x == r2.cov2x).all() and r1.c.dp_cells == r2.c.dp_cells and r1.c.n_anchors == r2.c.n_anchors
    assert np.array_equal(r1.af, r2.af, equal_nan=True)
    ok = ~np.isnan(r1.af)
    assert ok.mean() > 0.8 and np.abs(np.minimum(r1.af[ok], 1) - b.meta["truth_af"][ok]).mean() < 0.12


def test_full_size_random_sample_matches_oracle_config2(full_config2):
    b, r = full_config2
    _random_sample_vs_oracle(b, r, 200, 20221103)


def test_full_size_random_sample_matches_oracle_config4(built):
    """BASELINE.json configs[3] at its full size (30 000 loci, 10.8 Gbase, the north-star batch) on one GPU: 200 random loci
    of the result against the oracle."""
    b = synth.generate("ont_30k_30x", 0, 30000)
    c = lib.Context(0)
    r = c.run(b)
    c.close()
    assert (r.cov2x[:, 0] != -2).all()
    _random_sample_vs_oracle(b, r, 200, 20221105)
    ok = ~np.isnan(r.af)
    assert ok.mean() > 0.8 and np.abs(np.minimum(r.af[ok], 1) - b.meta["truth_af"][ok]).mean() < 0.15


@pytest.mark.parametrize("cfg", ["ont_3k_50x", "clr_3k_40x", "hifi_3k_40x"])
def test_reads_align_to_their_simulated_origin(ctx, cfg):
    """Truth check independent of the oracle: the primary alignment of a simulated read on the forward contig lies on the
    stretch of the contig the read was drawn from, on the strand it was drawn from (>= 99 % of the reads that share
    at least 500 bases with the contig)."""
    b = synth.generate(cfg, 40, 40)
    r = ctx.run(b, want_aln=True)
    org = b.meta["read_origin"]
    locus_of_read = np.repeat(np.arange(b.n_loci), np.diff(b.locus_read_begin))
    prim = {}
    for a in r.alns:
        if a["strand"] == 0 and (a["flag"] & 0x900) == 0:
            prim[int(a["read"])] = a
    n_exp = n_ok = 0
    for rd in range(b.n_reads):
        l = locus_of_read[rd]
        L, ts, te = int(b.contig_len[l]), int(b.te_start[l]), int(b.te_end[l])
        start, ln, rc, ctx_len = (int(v) for v in org[rd])
        has_te = int(b.meta["read_truth"][rd])
        # haplotype coordinates -> forward-contig coordinates (polishing indels move them by a few bases at most)
        if has_te:
            segs = [(start - ctx_len + ts, start + ln - ctx_len + ts)]
        else:
            segs = [(start - ctx_len + ts, min(start + ln, ctx_len) - ctx_len + ts), (max(start, ctx_len) - ctx_len + te, start + ln - ctx_len + te)]
        segs = [(max(s, 0), min(e, L)) for s, e in segs]
        segs = [(s, e) for s, e in segs if e - s >= 500]
        if not segs:
            continue
        n_exp += 1
        a = prim.get(rd)
        if a is None or int(a["rev"]) != rc:
            continue
        n_ok += any(min(int(a["re"]), e) - max(int(a["rs"]), s) >= 0.8 * min(e - s, int(a["re"]) - int(a["rs"])) for s, e in segs)
    assert n_exp > 0.8 * b.n_reads and n_ok >= 0.99 * n_exp, (n_ok, n_exp, b.n_reads)


def test_get_af_dropin(built, tmp_path):
    """The stage API on files, GPU backend, against the same host code fed by the oracle."""
    b = synth.generate("ont_3k_50x", 0, 4, depth=10)
    kw = _make_stage3_artifacts(tmp_path, b, drop_contig=1)
    te_freq = stage4.get_af(**kw)
    ref = orc.af_run(b.subset([0, 2, 3]), threads=0, want_depth=False, want_aln=False)
    names = [f"chr1_{10000 * (l + 1)}_{10000 * (l + 1) + 1}" for l in range(4)]
    assert set(te_freq) == {names[0], names[2], names[3]}
    for j, l in enumerate((0, 2, 3)):
        d = te_freq[names[l]]
        got = [d[k] for k in ("te_5p_cov", "te_3p_cov", "flank_5p_cov", "flank_3p_cov", "te_5p_cov_rc", "te_3p_cov_rc", "flank_5p_cov_rc", "flank_3p_cov_rc")]
        assert got == [None if c == -1 else c / 2 for c in ref.cov2x[j].tolist()]
        g = ref.af[j]
        assert d["freq"] == (None if np.isnan(g) else round(1 if g > 1 else g, 3))


def test_full_size_properties(ctx):
    """Properties at a larger size than the oracle is run on: depth equals the M-block sum, rerun determinism."""
    b = synth.generate("ont_3k_50x", 100, 60)
    r = ctx.run(b, want_depth=True, want_aln=True)
    al = r.alns
    off = np.concatenate([[0], np.cumsum(2 * b.contig_len.astype(np.int64))])
    locus_of_read = np.repeat(np.arange(b.n_loci), np.diff(b.locus_read_begin))
    marks = np.zeros(len(r.depth) + 1, np.int64)
    for i in np.nonzero((al["flag"] & 0x100) == 0)[0]:
        a = al[i]; c = r.cigar_of(i); op, ln = c & 0xf, (c >> 4).astype(np.int64)
        l = locus_of_read[a["read"]]; base = off[l] + a["strand"] * b.contig_len[l]
        ref_adv = np.where((op == 0) | (op == 2), ln, 0)
        starts = base + a["rs"] + np.concatenate([[0], np.cumsum(ref_adv)[:-1]])
        m = op == 0
        np.add.at(marks, starts[m], 1); np.add.at(marks, starts[m] + ln[m], -1)
    assert (np.cumsum(marks)[:-1] == r.depth).all()
    assert int(r.c.n_aln_blocks) == int(sum(((r.cigar_of(i) & 0xf) == 0).sum() for i in np.nonzero((al["flag"] & 0x100) == 0)[0]))
    r2 = ctx.run(b, want_depth=True)
    assert (r2.depth == r.depth).all() and (r2.cov2x == r.cov2x).all()
    ok = ~np.isnan(r.af)
    assert ok.mean() > 0.8 and np.abs(np.minimum(r.af[ok], 1) - b.meta["truth_af"][ok]).mean() < 0.12      # estimates track the simulated truth


def test_loci_without_any_read(ctx):
    """n_loci > 0 and n_reads == 0 (every shard of a partition may look like this): zero coverage, AF None, no launch error."""
    from telr_b200.batch import Batch
    b = synth.generate("ont_3k_50x", 0, 3, depth=6)
    none = np.zeros(0, np.int64)
    b0 = Batch(b.preset, b.seq2, b.nmask, none, none.astype(np.int32), none.astype(np.uint32), np.zeros(4, np.int32),
               b.contig_off.copy(), b.contig_len.copy(), b.te_start.copy(), b.te_end.copy())
    r = ctx.run(b0, want_depth=True)
    ro = orc.af_run(b0, threads=0, want_aln=False)
    assert (r.cov2x == ro.cov2x).all() and (r.depth == 0).all() and np.isnan(r.af).all()


def test_contig_longer_than_the_shared_memory_depth_row(ctx):
    """A 70 kb local assembly (ordinary for Flye): the depth row moves to global memory, nothing fails, results equal the oracle."""
    b = synth.generate("ont_3k_50x", 0, 2, depth=5, te_min=64000, te_max=65000, te_median=64500, n_families=2)
    assert int(b.contig_len.max()) > 60000
    r = ctx.run(b, want_depth=True, want_aln=True)
    ro = orc.af_run(b, threads=0)
    util.assert_same_results(r, ro)
    assert r.c.dp_cells == ro.c.dp_cells


def _rec_key(r):
    return (r["qname"], r["flag"], r["tid"], r["pos"], r["mapq"], tuple(int(x) for x in r["cigar"]), bytes(r["seq"]) if not isinstance(r["seq"], str) else r["seq"], r["aux"])


def test_get_af_keeps_the_realign_bams(built, tmp_path, monkeypatch):
    """Row f2: with TELR_B200_KEEP_BAM=1 the drop-in leaves <locus>.realign.sort.bam / <locus>.revcomp.realign.sort.bam (+ .bai) in
    telr_reads/ like realignment() (TELR_te.py:502-515); their records equal the SAM records built from the oracle's regions
    (FLAG 0x10/0x100/0x800, POS, MAPQ, CIGAR with clips, SEQ, tags) and `samtools depth` over them gives the coverage integers."""
    from telr_b200 import realign
    from tests.test_host import _depth_from_records
    b = synth.generate("ont_3k_50x", 0, 3, depth=10)
    kw = _make_stage3_artifacts(tmp_path, b)
    monkeypatch.setenv("TELR_B200_KEEP_BAM", "1")
    te_freq = stage4.get_af(**kw)
    assert len(te_freq) == 3
    ro = orc.af_run(b, threads=0)
    names = [f"L{l:06d}_R{r - b.locus_read_begin[l]:04d}" for l in range(3) for r in range(b.locus_read_begin[l], b.locus_read_begin[l + 1])]
    off = np.concatenate([[0], np.cumsum(2 * b.contig_len.astype(np.int64))])
    for l in range(3):
        L = int(b.contig_len[l])
        for strand, sfx in ((0, ""), (1, ".revcomp")):
            p = os.path.join(kw["out"], "telr_reads", f"chr1_{10000 * (l + 1)}_{10000 * (l + 1) + 1}{sfx}.realign.sort.bam")
            assert os.path.isfile(p) and os.path.isfile(p + ".bai")
            _, refs, recs = realign.read_bam(p)
            assert refs == [("ctg1", L)]
            want = realign.strand_records(b, ro, l, strand, names)
            order = sorted(range(len(want)), key=lambda i: (want[i]["tid"] & 0xffffffff, want[i]["pos"]))      # samtools sort (stable)
            assert len(recs) == len(want)
            for g, i in zip(recs, order):
                w = want[i]
                assert (g["qname"], g["flag"], g["tid"], g["pos"], g["mapq"]) == (w["qname"], w["flag"], w["tid"], w["pos"], w["mapq"])
                assert (g["cigar"] == w["cigar"]).all() and g["seq"] == w["seq"].decode()
            assert (_depth_from_records(recs, L) == ro.depth[off[l] + strand * L: off[l] + (strand + 1) * L]).all()


def test_polishing_alignment_r2k_matches_oracle(ctx, tmp_path):
    """Row f1: `minimap2 -ax map-ont -r2k contig reads | samtools sort` (TELR_assembly.py:199-212) = the same kernels with bw 2000."""
    from telr_b200 import realign
    b0 = synth.generate("ont_3k_50x", 11, 1, depth=14)
    acgt = np.frombuffer(b"ACGTN", np.uint8)
    contig = acgt[b0.unpack(int(b0.contig_off[0]), int(b0.contig_len[0]))].tobytes()
    reads = [(f"r{r}", acgt[b0.unpack(int(b0.read_off[r]), int(b0.read_len[r]))].tobytes()) for r in range(b0.n_reads)]
    bam = str(tmp_path / "polish.bam")
    n = realign.align_to_bam("ctg1", contig, reads, "map-ont", bam, bw=2000, ctx=ctx)
    _, refs, recs = realign.read_bam(bam)
    assert refs == [("ctg1", len(contig))] and len(recs) == n and os.path.isfile(bam + ".bai")
    b, names = realign._batch_of("map-ont", [("ctg1", contig)], [reads])
    orc.set_bw(2000, 0)
    try:
        ro = orc.af_run(b, threads=0)
    finally:
        orc.set_bw(0, 0)
    want = realign.strand_records(b, ro, 0, 0, names)
    order = sorted(range(len(want)), key=lambda i: (want[i]["tid"] & 0xffffffff, want[i]["pos"]))
    assert len(recs) == len(want)
    for g, i in zip(recs, order):
        w = want[i]
        assert (g["qname"], g["flag"], g["pos"], g["mapq"]) == (w["qname"], w["flag"], w["pos"], w["mapq"]) and (g["cigar"] == w["cigar"]).all()
    # the option is per call: the next run on the same context uses the preset bandwidth again
    r_def = ctx.run(b, want_aln=True)
    ro_def = orc.af_run(b, threads=0)
    util.assert_same_results(r_def, ro_def)


def test_annotation_alignments_paf_match_oracle(ctx):
    """Row f4: `minimap2 -cx <preset> --secondary=no contig query` (VCF insertion sequence vs its contig, TELR_te.py:68-78) and
    `minimap2 -cx <preset> contig te_library` (TELR_te.py:119-132): PAF lines with cg:Z:, all contigs in one batch."""
    from telr_b200 import realign
    rng = np.random.default_rng(17)
    b0 = synth.generate("ont_3k_50x", 20, 4, depth=4)
    acgt = np.frombuffer(b"ACGTN", np.uint8)
    contigs, queries = [], []
    lib_te = [(f"TE{k}", acgt[rng.integers(0, 4, int(rng.integers(800, 3000)))].tobytes()) for k in range(3)]
    for l in range(4):
        c = b0.unpack(int(b0.contig_off[l]), int(b0.contig_len[l]))
        contigs.append((f"chr1_{l}_ctg", acgt[c].tobytes()))
        ts, te = int(b0.te_start[l]), int(b0.te_end[l])
        ins = _mut(rng, c[ts:te], 0.03)                                   # the insertion sequence Sniffles reported for this locus
        qs = [(f"ins{l}", acgt[ins].tobytes()), (f"ins{l}_rc", acgt[(3 - ins[::-1])].tobytes())] + lib_te + [(f"fam{l}", acgt[_mut(rng, c[ts:te], 0.1)].tobytes())]
        queries.append(qs)
    for secondary in (False, True):
        got = realign.align_to_paf(contigs, queries, "map-ont", secondary=secondary, ctx=ctx)
        b, names = realign._batch_of("map-ont", contigs, queries)
        ro = orc.af_run(b, threads=0)
        want = realign.paf_lines(b, ro, names, [c[0] for c in contigs], secondary=secondary)
        assert got == want and len(got) >= 8
        f = got[0].split("\t")
        assert len(f) >= 12 and f[4] in "+-" and f[-1].startswith("cg:Z:") and int(f[3]) - int(f[2]) > 500
    assert all("tp:A:S" not in ln for ln in realign.align_to_paf(contigs, queries, "map-ont", secondary=False, ctx=ctx))
