"""First-light GPU parity driver (run on the B200 box): stage by stage against the oracle, verbose diagnostics."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
from telr_b200 import synth, lib
from telr_b200.batch import pack_sequences
from tests import orc

def section(t): print("\n==== " + t, flush=True)

ctx = lib.Context(0)
rng = np.random.default_rng(7)

def check_sketch(w, k, hpc, with_n):
    seqs = []
    for ln in (5, 14, 15, 24, 25, 26, 100, 2047, 2048, 2049, 4096, 4100, 10000, 33333):
        s = rng.integers(0, 4, ln).astype(np.uint8)
        if hpc:  # add homopolymers
            for _ in range(ln // 50):
                p = rng.integers(0, ln); s[p:p + rng.integers(2, 12)] = s[p]
        if ln > 300:  # tandem repeat to force identical hashes
            s[200:290] = np.tile(s[200:207], 13)[:90]
        b = np.array(list(b"ACGT"), np.uint8)[s]
        if with_n and ln > 30:
            for _ in range(max(1, ln // 400)): b[rng.integers(0, ln)] = ord("N")
        seqs.append(bytes(b))
    seq2, nmask, offs, lens = pack_sequences(seqs)
    x, y, off = ctx.sketch(seq2, nmask, offs, lens, w, k, hpc)
    bad = 0
    for i, s in enumerate(seqs):
        nt4 = np.array([{65: 0, 67: 1, 71: 2, 84: 3}.get(c, 4) for c in s], np.uint8)
        ox, oy = orc.sketch(nt4, w, k, hpc)
        gx, gy = x[off[i]:off[i + 1]], y[off[i]:off[i + 1]]
        if len(ox) != len(gx) or not ((ox == gx).all() and (oy == gy).all()):
            bad += 1
            print(f"  sketch MISMATCH len={len(s)} w={w} k={k} hpc={hpc} N={with_n}: oracle {len(ox)} gpu {len(gx)}")
            m = min(len(ox), len(gx))
            d = np.nonzero((ox[:m] != gx[:m]) | (oy[:m] != gy[:m]))[0]
            if len(d): print("   first diff at", d[0], "oracle", hex(int(ox[d[0]])), int(oy[d[0]]), "gpu", hex(int(gx[d[0]])), int(gy[d[0]]))
    print(f"sketch w={w} k={k} hpc={hpc} N={with_n}: {'OK' if not bad else str(bad)+' BAD'}", flush=True)
    return bad

section("sketch")
tot_bad = 0
for (w, k, hpc) in ((10, 15, 0), (10, 19, 1), (19, 19, 0)):
    for with_n in (False, True):
        try: tot_bad += check_sketch(w, k, hpc, with_n)
        except Exception: traceback.print_exc(); tot_bad += 1

section("depth/af stage")
try:
    nl = 40
    clen = rng.integers(600, 9000, nl).astype(np.int32)
    ts = np.array([rng.integers(0, L - 150) for L in clen], np.int32); te = np.array([min(L, s + rng.integers(20, 4000)) for L, s in zip(clen, ts)], np.int32)
    ts[0] = 300; ts[1] = 299; te[2] = ts[2] + 60
    bl, bs, bn = [], [], []
    for l in range(nl):
        for s in range(2):
            for _ in range(rng.integers(0, 80)):
                st = rng.integers(0, clen[l]); ln = rng.integers(1, 3000)
                bl.append(2 * l + s); bs.append(st); bn.append(min(ln, clen[l] - st))
    depth, cov, af = ctx.depth_af(clen, ts, te, np.array(bl), np.array(bs), np.array(bn))
    # oracle
    import ctypes as C
    L_ = orc.lib(); off = 0; bad = 0
    for l in range(nl):
        L = int(clen[l]); d = [np.zeros(L, np.int32), np.zeros(L, np.int32)]
        for q, st, ln in zip(bl, bs, bn):
            if q >> 1 == l: d[q & 1][st:st + ln] += 1
        c8 = np.zeros(8, np.int32); a = C.c_double()
        L_.orc_cov_af(d[0].ctypes.data, d[1].ctypes.data, L, int(ts[l]), int(te[l]), 100, 200, 50, 50, c8.ctypes.data, C.byref(a))
        g0 = depth[off:off + L]; g1 = depth[off + L: off + 2 * L]; off += 2 * L
        ok = (g0 == d[0]).all() and (g1 == d[1]).all() and (c8 == cov[l]).all() and (np.isnan(a.value) == np.isnan(af[l])) and (np.isnan(af[l]) or a.value == af[l])
        if not ok:
            bad += 1; print("  depth/af MISMATCH locus", l, "cov oracle", c8, "gpu", cov[l], "af", a.value, af[l], "depth eq", (g0 == d[0]).all(), (g1 == d[1]).all())
    print("depth/af:", "OK" if not bad else f"{bad} BAD"); tot_bad += bad
except Exception: traceback.print_exc(); tot_bad += 1

section("DP stage")
try:
    o = orc.opt(0)
    tasks = []; qs = []; ts_ = []; qo = to = 0; cases = []
    def mut(s, rate):
        out = []
        for c in s:
            r = rng.random()
            if r < rate / 3: continue
            if r < 2 * rate / 3: out.append(c); out.append(rng.integers(0, 4)); continue
            if r < rate: out.append((c + 1 + rng.integers(0, 3)) % 4); continue
            out.append(c)
        return np.array(out, np.uint8)
    for (ql, flag, w, zd, eb, rate) in [(50, 0x08, 30001, 400, -1, .1), (300, 0x08, 30001, 400, -1, .12), (1000, 0x08, 30001, 400, -1, .12),
                                        (300, 0, 30001, 400, -1, .12), (700, 0x40, 751, 400, -1, .12), (700, 0x40 | 0x02 | 0x80, 751, 400, -1, .12),
                                        (2500, 0x40, 751, 400, -1, .15), (2500, 0x40 | 0x02 | 0x80, 751, 200, -1, .15), (33, 0x40, 751, 400, -1, .5),
                                        (1, 0x08, 30001, 400, -1, 0.), (400, 0x40, 751, 400, -1, .9), (3000, 0x40, 100, 400, -1, .1)] * 3:
        t = rng.integers(0, 4, max(1, int(ql * rng.uniform(.7, 1.4)))).astype(np.uint8)
        q = mut(t, rate)[:ql] if rate < .8 else rng.integers(0, 4, ql).astype(np.uint8)
        if len(q) == 0: q = np.zeros(1, np.uint8)
        if ql == 300 and flag == 0: t = np.concatenate([t[:100], rng.integers(0, 4, 2500).astype(np.uint8), t[100:]])   # big deletion, exact + zdrop
        if rng.random() < .3 and len(q) > 40: q[rng.integers(0, len(q))] = 4
        tasks.append((qo, to, len(q), len(t), w, zd, eb, flag)); qs.append(q); ts_.append(t); qo += len(q); to += len(t)
    ta = np.array(tasks, lib.DPTASK_DTYPE)
    out, cig = ctx.dp(0, ta, np.concatenate(qs), np.concatenate(ts_))
    bad = 0
    for i, (q, t) in enumerate(zip(qs, ts_)):
        w, zd, eb, flag = tasks[i][4:]
        ref = orc.ksw_extd2(q, t, o, w, zd, eb, flag); g = out[i]
        names = ["max", "max_q", "max_t", "zdropped", "reach_end", "cells"] + (["score"] if (flag & 0x40) == 0 and not ref["zdropped"] else []) + (["mqe", "mqe_t"] if flag & 0x40 and not ref["zdropped"] else [])
        if flag & 0x08: names = [n for n in names if n not in ("max", "max_q", "max_t")]
        ok = all(int(ref[n]) == int(g[n]) for n in names)
        gc = cig[g["cigar_off"]: g["cigar_off"] + g["n_cigar"]]
        ok = ok and len(gc) == len(ref["cigar"]) and (gc == ref["cigar"]).all()
        if not ok:
            bad += 1; print(f"  DP MISMATCH task {i} qlen={len(q)} tlen={len(t)} flag={flag:#x} w={w}:", {n: (int(ref[n]), int(g[n])) for n in names}, "ncig", len(ref['cigar']), int(g['n_cigar']))
    print("DP stage:", "OK" if not bad else f"{bad} BAD of {len(tasks)}"); tot_bad += bad
except Exception: traceback.print_exc(); tot_bad += 1

section("full pipeline, 12 ONT loci")
try:
    b = synth.generate("ont_3k_50x", 0, 12)
    t0 = time.time(); r = ctx.run(b, want_depth=True, want_aln=True); t1 = time.time()
    print("gpu run", t1 - t0, "s", r.stats())
    ro = orc.af_run(b, threads=0); print("oracle", time.time() - t1, "s cells", ro.c.dp_cells)
    print("cov equal:", (r.cov2x == ro.cov2x).all(), "af equal:", np.array_equal(r.af, ro.af, equal_nan=True), "depth equal:", (r.depth == ro.depth).all(),
          "cells", r.c.dp_cells, ro.c.dp_cells, "n_aln", r.c.n_aln, ro.c.n_aln)
    if not (r.cov2x == ro.cov2x).all(): print(r.cov2x, ro.cov2x)
    ga, oa = r.alns, ro.alns
    nb = 0
    if len(ga) == len(oa):
        for i in range(len(ga)):
            same = all(ga[i][f] == oa[i][f] for f in ("read", "strand", "rs", "re", "qs", "qe", "rev", "flag", "dp_max", "mlen", "blen", "n_cigar")) and (r.cigar_of(i) == ro.cigar_of(i)).all()
            if not same:
                nb += 1
                if nb < 6: print("  aln MISMATCH", i, ga[i], oa[i])
    else:
        nb = abs(len(ga) - len(oa)) + 1
        # find first differing read
        import collections
        cg = collections.Counter(zip(ga["read"].tolist(), ga["strand"].tolist())); co = collections.Counter(zip(oa["read"].tolist(), oa["strand"].tolist()))
        diff = [(k, cg[k], co[k]) for k in set(cg) | set(co) if cg[k] != co[k]][:10]
        print("  aln count differs; (read,strand): gpu n, oracle n:", diff)
    print("alignment records:", "OK" if nb == 0 else f"{nb} BAD"); tot_bad += nb
    tot_bad += int(not (r.cov2x == ro.cov2x).all()) + int(not (r.depth == ro.depth).all())
except Exception: traceback.print_exc(); tot_bad += 1
print("\nTOTAL BAD:", tot_bad)
