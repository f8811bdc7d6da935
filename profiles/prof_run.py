"""Small driver for ncu: two passes of the whole stage-4 body over N synthetic loci (first pass warms allocations)."""
import sys
sys.path.insert(0, ".")  # run from the repo root
from telr_b200 import lib, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
cfg = sys.argv[2] if len(sys.argv) > 2 else "ont_3k_50x"
b = synth.generate(cfg, 0, n)
ctx = lib.Context(0)
for _ in range(2):
    r = ctx.run(b)
print(r.stats())
