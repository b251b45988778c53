#!/bin/bash
# Full GPU check: parity tests, bench (own + reference arm), ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --depth ${DEPTH:-8} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','cpu_baseline','fps_us_per_scene')})
print(d['roofline'])
for k,v in d['rooflines'].items(): print("  %-24s %8.1f us %-5s frac %.3f"%(k,v['ms']*1e3,v['bound'],v['frac']))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-300
timeout 300 python bench.py --steps 8 --warmup 3 --depth ${DEPTH:-8} --skip-cpu --skip-probe > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --depth ${DEPTH:-8} --skip-cpu --skip-probe > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
