"""Gap fills alone: warp-per-fill systolic kernel against one-fill-per-lane (TELR_FILL_LANES=1), telr_af_dp on N fills of ~q x t."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from telr_b200 import lib

def run(n, ql, tl, jitter, lanes):
    os.environ["TELR_FILL_LANES"] = "1" if lanes else "0"
    os.environ["TELR_CENSUS"] = "1"
    rng = np.random.default_rng(3)
    tasks, qs, ts = [], [], []
    qo = to = 0
    for _ in range(n):
        t = rng.integers(0, 4, int(tl * (1 + rng.uniform(-jitter, jitter)))).astype(np.uint8)
        q = rng.integers(0, 4, int(ql * (1 + rng.uniform(-jitter, jitter)))).astype(np.uint8)
        m = min(len(q), len(t)); keep = rng.random(m) < 0.9
        q[:m][keep] = t[:m][keep]
        tasks.append((qo, to, len(q), len(t), -1, 400, -1, 0x08)); qs.append(q); ts.append(t); qo += len(q); to += len(t)
    c = lib.Context(0)
    T = np.array(tasks, lib.DPTASK_DTYPE); Q = np.concatenate(qs); Tt = np.concatenate(ts)
    c.dp(0, T, Q, Tt)
    t0 = time.time(); out, cig = c.dp(0, T, Q, Tt); dt = time.time() - t0
    c.close()
    return int(out["cells"].sum()), dt, int(out["score"].sum()), len(cig)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    if len(sys.argv) > 2:           # single configuration for ncu: N lanes(0/1) q t jitter
        cells, dt, sc, nc = run(n, int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), int(sys.argv[2]))
        print(cells, dt)
        sys.exit(0)
    for ql, tl, jit in ((250, 250, 0.0), (230, 230, 0.15), (250, 300, 0.15)):
        for lanes in (0, 1):
            cells, dt, sc, nc = run(n, ql, tl, jit, lanes)
            print(f"q~{ql} t~{tl} jitter {jit} lanes={lanes}: {cells/1e9:.2f} Gcells wall {dt*1e3:.1f} ms (incl. copies) score-sum {sc} cigar words {nc}", flush=True)
