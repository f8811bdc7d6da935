#!/bin/bash
# A/B harness: runs the 296-locus bench once per prebuilt library variant in telr_b200/_variants/ (restores nothing: run on a gpurun snapshot)
for v in "$@"; do
  cp telr_b200/_variants/$v.so telr_b200/_telr_af.so
  printf "%s " $v
  python bench.py --loci 296 --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['gcups'],1), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()})"
done
