"""Summaries of ncu captures for profiles/: `ncu_summary.py full <a.ncu-rep> [<b.ncu-rep> ...]` prints one column per captured
kernel launch (raw page, the metrics DESIGN.md quotes); `ncu_summary.py launches <launch list csv>` prints count / total / share
per kernel of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import subprocess
import sys
from collections import defaultdict

METRICS = """dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed gpu__time_duration.sum
launch__block_size launch__grid_size launch__registers_per_thread launch__shared_mem_per_block_dynamic sm__inst_executed.avg.per_cycle_elapsed
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
smsp__sass_average_data_bytes_per_sector_mem_global_op_st.ratio l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum lts__t_sector_hit_rate.pct""".split()
STALLS = "barrier branch_resolving dispatch_stall drain lg_throttle long_scoreboard math_pipe_throttle membar mio_throttle misc no_instruction not_selected selected short_scoreboard sleeping tex_throttle wait".split()
METRICS += ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s for s in STALLS]


def full(reps):
    cols = []
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            cols.append((dict(zip(hdr, r)), dict(zip(hdr, units))))
    print("%-110s %s" % ("Kernel Name", " | ".join(c["Kernel Name"].replace("ssc::", "")[:60] for c, _ in cols)))
    for m in METRICS:
        if m not in cols[0][0]:
            continue
        print("%-95s %-14s %s" % (m, cols[0][1][m], " | ".join(c[m] for c, _ in cols)))


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[h]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > mv:
            agg[r[kn].split("(")[0].replace("void ", "").replace("ssc::", "")].append(float(r[mv].replace(",", "")) / 1e6)
    tot = sum(sum(v) for v in agg.values())
    print("%-50s %6s %12s %7s" % ("kernel", "count", "total_ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-50s %6d %12.3f %6.1f%%" % (k, len(v), sum(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    (full if sys.argv[1] == "full" else launches)(sys.argv[2:] if sys.argv[1] == "full" else sys.argv[2])
