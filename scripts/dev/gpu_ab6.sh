#!/bin/bash
mkdir -p gpurun_out
run() { env $2 timeout 300 python bench.py --steps 96 --warmup 5 --depth 8 --skip-cpu --skip-probe $3 > gpurun_out/ab6_$1.json 2> gpurun_out/ab6_$1.err
  python -c "
import json;t=open('gpurun_out/ab6_$1.json').read().strip().splitlines()[-1].replace('Infinity','null').replace('NaN','null');d=json.loads(t);print('$1','value',round(d['value']),'ms/step',round(d['ms_per_step'],4))" || tail -3 gpurun_out/ab6_$1.err; }
run sideonly_noatt "PCOPS_PIPE_PARTS=side" "--attention 0"
run sideonly_noatt_allpairs "PCOPS_PIPE_PARTS=side" "--attention 0 --grid 0"
run sideonly "PCOPS_PIPE_PARTS=side" ""
run noatt_allpairs "A=1" "--attention 0 --grid 0"
