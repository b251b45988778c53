#!/bin/bash
# usage: gpu_ncu_ops.sh <ops> <kernel-regex> <outname> [levels]
mkdir -p gpurun_out
LV=${4:-0}
timeout 300 python scripts/opbench.py --ops $1 --levels $LV --iters 3 > gpurun_out/plain_$3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/prof_$3 python scripts/opbench.py --ops $1 --levels $LV --iters 3 > gpurun_out/ncu_$3.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$3.log
