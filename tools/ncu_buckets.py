"""Executed-count buckets of one kernel from an ncu source-page CSV: instructions that run once per pair, once per mate, once
per chunk of a slow read ... fall into distinct `executions / pair` buckets, which is how DESIGN.md section 9 splits the
warp-instructions per pair of generate_slots_kernel without relying on (inlined) line attribution.
usage: ncu -i x.ncu-rep --page source --csv --kernel-name regex:generate_slots > src.csv ; ncu_buckets.py src.csv <pairs per launch>"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else 2097152.0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
end = starts[1] if len(starts) > 1 else len(rows)          # first captured launch only
hdr = rows[starts[0] + 1]
ci = hdr.index("Instructions Executed")
dyn, stat = Counter(), Counter()
for r in rows[starts[0] + 2:end]:
    if len(r) > ci and r[ci].isdigit():
        k = round(int(r[ci]) / pairs, 2)
        dyn[k] += int(r[ci])
        stat[k] += 1
tot = sum(dyn.values())
print("warp-instructions per pair: %.1f" % (tot / pairs))
for k, v in sorted(dyn.items(), key=lambda kv: -kv[1])[:24]:
    print("executions/pair %5.2f   static %4d   instructions/pair %7.1f   %5.1f%%" % (k, stat[k], v / pairs, 100.0 * v / tot))
