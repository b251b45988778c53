#!/bin/bash
# ncu launch list (per-launch device time) of a short bench run: forward, e2e, training step, whole-layer forward, two
# scans of config 4.  usage: gpurun -- 'bash scripts/gpu_ncu_bench_list.sh <outname>'
mkdir -p gpurun_out
NAME=${1:-launches}
CMD="python bench.py --steps 4 --warmup 3 --depth 2 --train-depth 2 --skip-probe --skip-cpu --scenes 2 --config5 0 --min-seconds 0"
timeout 300 $CMD > gpurun_out/plain_$NAME.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/$NAME.csv $CMD > gpurun_out/ncu_$NAME.log 2>&1
echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/$NAME.csv > gpurun_out/${NAME}_summary.txt; head -60 gpurun_out/${NAME}_summary.txt
