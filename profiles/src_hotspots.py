"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by source file / line range:
warp-instructions executed and stall samples per source line, grouped by file and by the function regions given below."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur_file = None
hdr = None
by_line = collections.defaultdict(lambda: [0, 0, 0])   # (file, line) -> [inst, samples, thread_inst]
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        by_line[(cur_file, int(r[0]))][0] += int(r[ii]); by_line[(cur_file, int(r[0]))][1] += int(r[si]); by_line[(cur_file, int(r[0]))][2] += int(r[ti])
    except ValueError:
        pass
tot_i = sum(v[0] for v in by_line.values()); tot_s = sum(v[1] for v in by_line.values())
by_file = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), v in by_line.items():
    for k in range(3): by_file[f][k] += v[k]
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}")
print("| file | warp-inst | share | samples share | avg active threads |\n|---|---:|---:|---:|---:|")
for f, v in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f"| {f} | {v[0]:,} | {100*v[0]/tot_i:.1f} % | {100*v[1]/max(tot_s,1):.1f} % | {v[2]/max(v[0],1):.1f} |")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"\ntop {n} lines")
for (f, l), v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{f}:{l}\t{v[0]:,}\t{100*v[0]/tot_i:.2f} %\tsamples {100*v[1]/max(tot_s,1):.2f} %\tthr {v[2]/max(v[0],1):.1f}")
