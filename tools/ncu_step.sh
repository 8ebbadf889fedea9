#!/bin/bash
# One training step of the bench workload under ncu with a small metric set: per-launch duration, DRAM bytes,
# tensor-pipe activity -> gpurun_out/ncu_step.csv (summarise with tools/ncu_step_summary.py).
# Usage (under gpurun): bash tools/ncu_step.sh [batch] [size]
set -u
B=${1:-64}; S=${2:-256}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,sm__warps_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size
C=1150
if [ -n "${NCU_LITE:-}" ]; then   # five metrics, 700 launches: a third of the profiler time (the full set costs ~15 GPU-minutes)
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
  M=$M,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
  C=700
fi
# steps 1-2 warm up (the first optimizer step also zero-fills 346 state tensors); the window below covers step 3 entirely
python tools/one_step.py 2 $B $S > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics $M --clock-control none --launch-skip 1300 --launch-count $C --csv --log-file gpurun_out/ncu_step.csv \
    python tools/one_step.py 5 $B $S > gpurun_out/ncu_run.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_run.log
