#!/bin/bash
mkdir -p gpurun_out
for g in 0 1; do for d in 8 12; do
timeout 300 python bench.py --steps 128 --warmup 5 --depth $d --grid $g --skip-cpu --skip-probe > gpurun_out/ab7_g${g}_d$d.json 2> gpurun_out/ab7_g${g}_d$d.err
python -c "
import json;d=json.loads(open('gpurun_out/ab7_g${g}_d$d.json').read().strip().splitlines()[-1]);print('grid',$g,'depth',$d,'value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))"
done; done
