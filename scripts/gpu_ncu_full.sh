#!/bin/bash
# `ncu --set full` captures of the kernels scripts/ncu_targets.py launches (one GPU; run only after the plain command
# exited 0).  usage: gpurun --timeout 1800 -- 'bash scripts/gpu_ncu_full.sh <outname> [groups...]'
mkdir -p gpurun_out
NAME=${1:-r2_full}; shift
timeout 300 python scripts/ncu_targets.py "$@" > gpurun_out/plain_$NAME.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$NAME.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/$NAME \
  python scripts/ncu_targets.py "$@" > gpurun_out/ncu_$NAME.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$NAME.log; ls -la gpurun_out/$NAME.ncu-rep
