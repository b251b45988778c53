#!/bin/bash
mkdir -p gpurun_out
for b in 16 32 64; do
timeout 300 python bench.py --steps 96 --warmup 5 --depth 8 --batch $b --skip-cpu --skip-probe > gpurun_out/ab2_b$b.json 2> gpurun_out/ab2_b$b.err; echo "batch $b rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/ab2_b$b.json').read().strip().splitlines()[-1]);print('batch',$b,'value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'host enqueue ms/step',round(d['host_enqueue_ms_per_step'],4),'e2e',round(d['e2e']['value']))"
done
