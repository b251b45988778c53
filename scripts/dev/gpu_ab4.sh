#!/bin/bash
mkdir -p gpurun_out
for s in 256 1 2 3 4 512; do
PCOPS_FPS_SHAPE=$s timeout 300 python bench.py --steps 96 --warmup 5 --depth 8 --skip-cpu --skip-probe > gpurun_out/ab4_$s.json 2> gpurun_out/ab4_$s.err
python -c "
import json;d=json.loads(open('gpurun_out/ab4_$s.json').read().strip().splitlines()[-1]);print('fps variant $s','value',round(d['value']),'ms/step',round(d['ms_per_step'],4), 'fps ms', round(d['roofline']['ms'],4))"
done
