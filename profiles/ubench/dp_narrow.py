"""Narrow extensions in isolation (the shape census of profiles/README.md: 57 % of the general-DP tasks of config 2 have
min(qlen, tlen) < 64 — reads that overhang a contig end): exact max, band 500, z-drop 400, through telr_af_dp.
Run once per library variant (TELR_VEC_EXT=1 windowed path; a -DTELR_THIN_EXT=1 build with TELR_VEC_EXT=2 takes the
register-resident path for tlen <= 128)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from telr_b200 import lib

ctx = lib.Context(0)
rng = np.random.default_rng(2)

def run(name, n, ql, tl, w, zdrop, flag):
    qs, ts, tasks = [], [], []
    qo = to = 0
    for _ in range(n):
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = rng.integers(0, 4, ql).astype(np.uint8)
        q[:tl] = t                                    # the read matches the contig end, then runs off it
        sub = rng.random(tl) < 0.08
        q[:tl][sub] = (q[:tl][sub] + 1) % 4
        tasks.append((qo, to, len(q), len(t), w, zdrop, -1, flag))
        qs.append(q); ts.append(t); qo += len(q); to += len(t)
    tasks = np.array(tasks, lib.DPTASK_DTYPE)
    q = np.concatenate(qs); t = np.concatenate(ts)
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        out, cig = ctx.dp(0, tasks, q, t)
        best = min(best, time.perf_counter() - t0)
    cells = int(out["cells"].sum())
    print(f"{name}: {n} tasks, {cells/1e9:.3f} Gcells, best {best*1e3:.1f} ms wall (incl. H2D/D2H) -> {cells/best/1e9:.1f} GCUPS", flush=True)

run("ext 1900x30 w=500 exact", 40000, 1900, 30, 500, 400, 0x40)
run("ext 1900x100 w=500 exact", 20000, 1900, 100, 500, 400, 0x40)
