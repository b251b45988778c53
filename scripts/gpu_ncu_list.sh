#!/bin/bash
# per-launch durations of an opbench run: gpu_ncu_list.sh <ops> <levels> <outname>
mkdir -p gpurun_out
timeout 300 python scripts/opbench.py --ops $1 --levels $2 --iters 3 > gpurun_out/plain_$3.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 300 --csv --log-file gpurun_out/list_$3.csv python scripts/opbench.py --ops $1 --levels $2 --iters 3 > gpurun_out/ncu_$3.log 2>&1
echo "ncu rc=$?"
