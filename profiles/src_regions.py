"""Static code size / dynamic instruction share / stall mix of source regions, from an
`ncu --page source --csv --print-source sass,cuda` export.  usage: src_regions.py export.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None; cur = None; curline = None
static = collections.Counter(); dyn = collections.Counter(); stalls = collections.defaultdict(collections.Counter)
seen = set()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed")
        sc = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] == '-':
        try: curline = int(r[0])
        except ValueError: curline = None
        continue
    if r[2].startswith('0x') and curline is not None:
        if r[2] in seen: continue          # an instruction inlined from several lines is listed once per line
        seen.add(r[2])
        static[cur] += 1
        try: dyn[cur] += int(r[ii])
        except ValueError: pass
        for i, h in sc:
            try: stalls[cur][h] += int(r[i])
            except ValueError: pass
tot = sum(dyn.values()); ts = collections.Counter()
for f in stalls: ts.update(stalls[f])
print(f"instructions: static {sum(static.values())}  dynamic {tot:,}")
print("stall mix (all):", {k: round(100 * v / sum(ts.values()), 1) for k, v in ts.most_common(7)})
for f, d in dyn.most_common(6):
    t = sum(stalls[f].values())
    print(f"{f:16s} static {static[f]:6d} ({static[f]*16/1024:6.1f} KB) dynamic {100*d/tot:5.1f} %  stalls:", {k: round(100 * v / max(t, 1), 1) for k, v in stalls[f].most_common(5)})
