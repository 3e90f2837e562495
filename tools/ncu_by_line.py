"""Aggregate an ncu source-page CSV (SASS rows) by CUDA source line using nvdisasm line info.
usage: ncu_by_line.py <ncu_source.csv> <nvdisasm_all.sass> <mangled kernel name> [top]"""
import csv
import re
import sys
from collections import defaultdict

ncu_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]
ci = hdr.index("Instructions Executed")
cs = hdr.index("# Samples")
csrc = hdr.index("Source")
dyn = [(r[csrc].strip(), int(r[ci]), int(r[cs])) for r in rows[2:] if len(r) > ci]
# nvdisasm: find .text.<kname> section
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("//--------------------- .text." + kname))
cur = None
stat = []
inl = None
for l in lines[start + 1:]:
    if l.startswith("//--------------------- "):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        stat.append((cur, m.group(2)))
print("static instrs", len(stat), "dynamic rows", len(dyn))
agg = defaultdict(lambda: [0, 0, 0])
for (loc, txt), (s, n, smp) in zip(stat, dyn):
    a = agg[loc]
    a[0] += 1; a[1] += n; a[2] += smp
tot = sum(a[1] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print("total dyn warp-instr %d, samples %d" % (tot, tots))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-22s static %5d  dyn %6.2f%%  samples %6.2f%%" % ("%s:%d" % loc if loc else "?", a[0], 100.0 * a[1] / tot, 100.0 * a[2] / max(tots, 1)))
