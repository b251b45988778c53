#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','config3_training_step','config4_whole_scene')})
for k,v in d['rooflines'].items():
    if 'grad' in k or 'bwd' in k: print("  %-28s %8.1f us %-5s frac %.3f"%(k,v['ms']*1e3,v['bound'],v['frac']))
PY
timeout 300 python scripts/opbench.py --ops grads --iters 10 2>&1 | grep -v "^{"
