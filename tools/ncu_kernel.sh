#!/bin/bash
# `ncu --set full` capture of a few launches of one kernel from a tools/ micro-benchmark.
# Usage (under gpurun): bash tools/ncu_kernel.sh <kernel regex> <count> <out name> <python args...>
set -u
K=$1; C=$2; O=$3; shift 3
python "$@" > gpurun_out/ncu_plain_$O.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -c $C -o gpurun_out/$O python "$@" > gpurun_out/ncu_run_$O.log 2>&1
echo "ncu rc $?"
