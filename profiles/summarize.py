"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (second pass only:
prof_run.py runs the workload twice, the first pass warms allocations)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
body = rows[1:]
body = body[len(body) // 2:]
agg, cnt = collections.OrderedDict(), collections.Counter()
for r in body:
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
    k = r[ki].split("(")[0].replace("telr::", "").replace("void ", "")
    agg[k] = agg.get(k, 0) + v
    cnt[k] += 1
tot = sum(agg.values())
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for k, v in agg.items():
    print(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / tot:.1f} % |")
print(f"| total | {sum(cnt.values())} | {tot:.3f} | 100 % |")
