#!/bin/bash
mkdir -p gpurun_out
run() { # name, env, args
  env $2 timeout 300 python bench.py --steps 96 --warmup 5 --depth 8 --skip-cpu --skip-probe $3 > gpurun_out/ab3_$1.json 2> gpurun_out/ab3_$1.err; 
  python -c "
import json;d=json.loads(open('gpurun_out/ab3_$1.json').read().strip().splitlines()[-1]);print('$1','value',round(d['value']),'ms/step',round(d['ms_per_step'],4))"
}
run full "A=1" ""
run noatt "A=1" "--attention 0"
run fpsonly "PCOPS_PIPE_PARTS=fps" ""
run sideonly "PCOPS_PIPE_PARTS=side" ""
run sideonly_noatt "PCOPS_PIPE_PARTS=side" "--attention 0"
run sideonly_allpairs "PCOPS_PIPE_PARTS=side" "--grid 0"
