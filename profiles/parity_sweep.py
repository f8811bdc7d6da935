"""Wider GPU-vs-oracle parity sweep than the test suite runs: slices of every configuration at several offsets, every alignment
field, CIGAR, per-base depth, coverage integers and AF compared (tests/util.assert_same_results)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from telr_b200 import lib, synth
from tests import orc, util

ctx = lib.Context(0)
tot = 0
for cfg, n, offs, kw in (("ont_3k_50x", 40, (300, 1100, 1900, 2700), {}), ("clr_3k_40x", 24, (200, 900, 1700, 2500), {}),
                         ("hifi_3k_40x", 24, (150, 1300, 2100), {}), ("ont_30k_30x", 60, (5000, 12000, 21000, 29000), {}),
                         ("poly_10k_200x", 4, (100, 4000, 9000), {}), ("ont_3k_50x", 12, (50, 650), dict(p_n=0.003))):
    for first in offs:
        t0 = time.time()
        b = synth.generate(cfg, first, n, **kw)
        r = ctx.run(b, want_depth=True, want_aln=True)
        ro = orc.af_run(b, threads=0)
        util.assert_same_results(r, ro)
        assert r.c.dp_cells == ro.c.dp_cells and r.c.n_anchors == ro.c.n_anchors
        tot += b.n_loci
        print(f"{cfg} loci [{first}, {first + n}) {kw or ''}: {b.n_reads} reads, {len(r.alns)} records, {int(r.c.dp_cells) / 1e9:.2f} G cells equal ({time.time() - t0:.1f} s)", flush=True)
print("parity sweep ok:", tot, "loci")
