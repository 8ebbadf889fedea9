#!/bin/bash
# One training step of the bench workload under ncu with a small metric set (a few replays per launch):
#   per-launch duration, DRAM bytes, tensor-pipe / issue / occupancy / L2 figures -> gpurun_out/ncu_step.csv
# Usage (under gpurun): bash tools/ncu_step.sh [batch] [size]
set -u
B=${1:-64}; S=${2:-256}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,sm__warps_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,l1tex__data_pipe_tc_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size
python tools/one_step.py 2 $B $S > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics $M --clock-control none --launch-skip 520 --launch-count 620 --csv --log-file gpurun_out/ncu_step.csv \
    python tools/one_step.py 3 $B $S > gpurun_out/ncu_run.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_run.log
