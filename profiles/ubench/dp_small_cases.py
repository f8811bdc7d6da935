"""Small-case sweep of the vectorised DP against the oracle (all flag combinations, tiny to band-limited shapes); run from the repo root."""
import sys; sys.path.insert(0, ".")
import numpy as np
from telr_b200 import lib
from tests import orc
ctx = lib.Context(0)
o = orc.opt(0)
rng = np.random.default_rng(5)
def mut(s, rate):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate / 3: continue
        if r < 2 * rate / 3: out.append(ch); out.append(rng.integers(0, 4)); continue
        out.append((ch + 1 + rng.integers(0, 3)) % 4 if r < rate else ch)
    return np.array(out, np.uint8)
import collections
bad = 0; byflag=collections.Counter(); small={}
for flag in (0x40, 0xC2, 0x0, 0x42):
    for (ql, tl, w) in [(1,1,751),(3,5,751),(4,4,751),(5,9,751),(8,8,751),(9,9,751),(9,12,751),(12,9,751),(10,10,751),(12,12,751),(13,13,751),(16,16,751),(9,17,751),(33,40,751),(64,64,751),(100,130,751),(129,200,751),(300,350,751),(700,900,751),(300,300,20),(900,1000,100)]:
        for rep in range(3):
            t = rng.integers(0, 4, tl).astype(np.uint8)
            q = mut(t, .12)[:ql]
            if len(q) == 0: q = np.zeros(1, np.uint8)
            task = np.array([(0, 0, len(q), len(t), w, 400, -1, flag)], lib.DPTASK_DTYPE)
            out, cig = ctx.dp(0, task, q, t)
            ref = orc.ksw_extd2(q, t, o, w, 400, -1, flag); g = out[0]
            names = ["max", "max_q", "max_t", "zdropped", "cells"]
            ok = all(int(ref[n]) == int(g[n]) for n in names)
            gc = cig[g["cigar_off"]: g["cigar_off"] + g["n_cigar"]]
            ok = ok and len(gc) == len(ref["cigar"]) and (gc == ref["cigar"]).all()
            if not ok:
                bad += 1; byflag[flag]+=1; small.setdefault(flag,(len(q),len(t)))
                if byflag[flag] <= 2: print("MISMATCH flag=%#x ql=%d tl=%d w=%d:" % (flag, len(q), len(t), w), {n: (int(ref[n]), int(g[n])) for n in names}, "cig", [(int(c)>>4, "MID"[int(c)&3]) for c in ref["cigar"]], [(int(c)>>4, "MID"[int(c)&3]) for c in gc], "q", q.tolist()[:20], "t", t.tolist()[:20], {n:(int(ref[n]),int(g[n])) for n in ("mqe","mqe_t","score","reach_end")})
print("bad", bad, dict(byflag), small)
