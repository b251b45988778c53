#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for s in 256 512 1024; do echo "FPS shape $s"; PCOPS_FPS_SHAPE=$s timeout 300 python scripts/opbench.py --ops fps --levels 0 2>&1 | grep -E "fps_sa1 |fps_sa1_B148"; done
timeout 600 python scripts/opbench.py --ops fps,ball,group,attention,three_nn,interp > gpurun_out/opbench.txt 2>&1; echo "opbench rc=$?"; grep -v "^{" gpurun_out/opbench.txt | grep -v "fps_sa1_B" | head -80
for d in 4 8; do
timeout 300 python bench.py --steps 64 --warmup 5 --depth $d --skip-cpu --skip-probe > gpurun_out/bench_d$d.json 2> gpurun_out/bench_d$d.err; echo "bench depth $d rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_d$d.json').read().strip().splitlines()[-1]);print('depth',$d,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'])"
done
