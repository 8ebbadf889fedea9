#!/bin/bash
# A handful of ncu metrics (three replay passes instead of the ~40 of --set full) for every launch matching a kernel regex.
# Usage (under gpurun): bash tools/ncu_metrics.sh <kernel regex> <count> <out name> <python args...>   -> gpurun_out/<out>.csv
set -u
K=$1; C=$2; O=$3; shift 3
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,\
l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,\
l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,\
lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,\
dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
python "$@" > gpurun_out/ncu_plain_$O.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:$K -c $C --csv --log-file gpurun_out/$O.csv python "$@" > gpurun_out/ncu_run_$O.log 2>&1
echo "ncu rc $?"
