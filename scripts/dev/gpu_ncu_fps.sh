#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/opbench.py --ops fps --levels 0 --iters 3 > gpurun_out/plain_fps.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_onchip -s 2 -c 1 -f -o gpurun_out/prof_fps python scripts/opbench.py --ops fps --levels 0 --iters 3 > gpurun_out/ncu_fps.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fps.log
