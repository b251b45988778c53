#!/bin/bash
mkdir -p gpurun_out
for f in 0 1; do
timeout 300 python bench.py --steps 96 --warmup 5 --depth 8 --fuse-layers $f --skip-cpu --skip-probe > gpurun_out/ab5_$f.json 2> gpurun_out/ab5_$f.err
python -c "
import json;d=json.loads(open('gpurun_out/ab5_$f.json').read().strip().splitlines()[-1]);print('fuse_layers $f','value',round(d['value']),'ms/step',round(d['ms_per_step'],4), 'launches/step', d['gpu_launches']/96)"
done
