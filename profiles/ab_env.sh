#!/bin/bash
# A/B over environment settings: profiles/ab_env.sh "TELR_AL_QUEUE=0" "TELR_AL_QUEUE=1 TELR_AL_EXT8=3" ...   (config 2, 296 loci unless LOCI is set)
for v in "$@"; do
  printf "%s : " "$v"
  env $v TELR_BENCH_CONFIG=${CFG:-ont_3k_50x} TELR_BENCH_BUDGET_S=60 timeout 600 python bench.py --loci ${LOCI:-296} --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'gcups', round(d['gcups'],1), 'frac', round(d['roofline']['frac'],4), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()}, d['outputs_sha1'][:10])"
done
