#!/bin/bash
# `ncu --set full` of the bandwidth-bound kernels of one training step (third step; the first 30 matching launches:
# sa_reduce at every level, then the level-0 / level-1 backward passes).  Usage (under gpurun): bash tools/ncu_stream_kernels.sh [count]
set -u
C=${1:-30}
K='regex:sa_reduce_kernel|rb_bwd1_kernel|rb_bwd2_kernel|rb_bwd3_kernel|bn_bwd_apply_kernel|bn_bwd_reduce_kernel|ag_bwd1_kernel|ag_bwd2_kernel|ag_bwd3_kernel'
RBU_NO_OVERLAP=1 python tools/one_step.py 2 > gpurun_out/ncu_plain_stream.log 2>&1 &&
RBU_NO_OVERLAP=1 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 136 -c $C -f -o gpurun_out/r02_i_stream_kernels \
    python tools/one_step.py 3 > gpurun_out/ncu_run_stream.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_run_stream.log
