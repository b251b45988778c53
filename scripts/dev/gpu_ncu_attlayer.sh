#!/bin/bash
mkdir -p gpurun_out
timeout 200 bash scripts/dev/gpu_attlayer.sh > gpurun_out/plain_attlayer.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_layer_c64 -s 2 -c 1 -f -o gpurun_out/prof_attlayer bash scripts/dev/gpu_attlayer.sh > gpurun_out/ncu_attlayer.log 2>&1
echo "ncu rc=$?"
