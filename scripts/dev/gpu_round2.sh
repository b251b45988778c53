#!/bin/bash
mkdir -p gpurun_out
./scripts/ubench/fp32_rate > gpurun_out/fp32_rate.txt 2>&1; cat gpurun_out/fp32_rate.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/opbench.py --ref 1 > gpurun_out/opbench.txt 2>&1; echo "opbench rc=$?"; cat gpurun_out/opbench.txt | head -80
for d in 1 2 4 8; do
timeout 300 python bench.py --steps 64 --warmup 5 --depth $d --skip-cpu --skip-probe > gpurun_out/bench_d$d.json 2> gpurun_out/bench_d$d.err; echo "bench depth $d rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_d$d.json').read().strip().splitlines()[-1]);print('depth',$d,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'])"
done
timeout 300 python bench.py --steps 64 --warmup 5 --depth 4 --graph 0 --skip-cpu --skip-probe > gpurun_out/bench_d4_eager.json 2> gpurun_out/bench_d4_eager.err; echo "bench eager rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_d4_eager.json').read().strip().splitlines()[-1]);print('eager depth 4 value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'])"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
