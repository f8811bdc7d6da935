"""Throughput of the two DP kernels in isolation (all warps run the same loop) through the stage entry point
telr_af_dp: gap fills (approximate-max global, 250x250) and extensions (exact max, band 751, 900x1000).
Compare with the cells/s the same kernels reach inside the mixed alignment kernel (bench.py)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from telr_b200 import lib

ctx = lib.Context(0)
rng = np.random.default_rng(1)

def mut(s, rate):
    r = rng.random(len(s))
    keep = r >= rate / 3
    out = s.copy()
    sub = (r >= rate / 3) & (r < 2 * rate / 3)
    out[sub] = (out[sub] + 1 + rng.integers(0, 3, sub.sum())) % 4
    return out[keep]

def run(name, n, ql, tl, w, zdrop, flag):
    qs, ts, tasks = [], [], []
    qo = to = 0
    for _ in range(n):
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = mut(t, .12)[:ql]
        tasks.append((qo, to, len(q), len(t), w, zdrop, -1, flag))
        qs.append(q); ts.append(t); qo += len(q); to += len(t)
    tasks = np.array(tasks, lib.DPTASK_DTYPE)
    q = np.concatenate(qs); t = np.concatenate(ts)
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        out, cig = ctx.dp(0, tasks, q, t)
        dt = time.perf_counter() - t0
        best = min(best, dt)
    cells = int(out["cells"].sum())
    print(f"{name}: {n} tasks, {cells/1e9:.2f} Gcells, best {best*1e3:.1f} ms wall (incl. H2D/D2H) -> {cells/best/1e9:.1f} GCUPS")

run("fill 250x250 approx", 60000, 250, 250, -1, -1, 0x08)
run("ext 900x1000 w=751 exact", 6000, 900, 1000, 751, 400, 0x40)
