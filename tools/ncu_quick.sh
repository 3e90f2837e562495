#!/bin/bash
# usage: tools/ncu_quick.sh <tag> [lib]   -- issue/stall/bank-conflict counters of generate_slots_kernel on the quick bench
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__cycles_elapsed.max
for s in long_scoreboard short_scoreboard math_pipe_throttle lg_throttle wait not_selected no_instruction dispatch_stall mio_throttle branch_resolving; do M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio; done
if [ -n "$2" ]; then export SIMUSCOP_CUDA_LIB=$2; fi
ncu --metrics $M --clock-control none -k regex:generate_slots -c 1 --csv --log-file gpurun_out/ncu_$1.csv python tools/quick_bench.py 256 1 ${3:-XTen} > gpurun_out/ncu_$1.log 2>&1
