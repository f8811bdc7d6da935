#!/bin/bash
# A/B over prebuilt library variants (telr_b200/_variants/<name>.so) and environment settings:
#   profiles/ab_var.sh "base TELR_AL_QUEUE=0" "fc12 TELR_AL_QUEUE=1" ...      (config 2, 296 loci unless CFG / LOCI are set; run on a gpurun snapshot)
for spec in "$@"; do
  v=${spec%% *}; e=""; [ "$spec" != "$v" ] && e=${spec#* }
  cp telr_b200/_variants/$v.so telr_b200/_telr_af.so
  printf "%s : " "$spec"
  env $e TELR_BENCH_CONFIG=${CFG:-ont_3k_50x} TELR_BENCH_BUDGET_S=60 timeout 600 python bench.py --loci ${LOCI:-296} --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'gcups', round(d['gcups'],1), 'frac', round(d['roofline']['frac'],4), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()}, d['outputs_sha1'][:10])"
done
