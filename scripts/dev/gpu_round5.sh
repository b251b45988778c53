#!/bin/bash
# round 5: parity, bench with config 3 + steady-state gathers, launch list (graph and eager), ncu --set full of the gathers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','cpu_baseline','fps_us_per_scene','config3_training_step','gathers_steady_state')})
for k,v in d['rooflines'].items(): print("  %-28s %8.1f us %-5s frac %.3f"%(k,v['ms']*1e3,v['bound'],v['frac']))
PY
timeout 300 python scripts/opbench.py --ops group,interp,grads --iters 10 2>&1 | grep -v "^{"
timeout 300 python bench.py --steps 8 --warmup 3 --skip-cpu --skip-probe --train 0 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_graph.csv python bench.py --steps 8 --warmup 3 --skip-cpu --skip-probe --train 0 > gpurun_out/ncu_graph.log 2>&1
echo "ncu graph rc=$?"; tail -3 gpurun_out/ncu_graph.log
timeout 300 python bench.py --steps 8 --warmup 3 --skip-cpu --skip-probe --train 0 --graph 0 > gpurun_out/plain0.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --skip-cpu --skip-probe --train 0 --graph 0 > gpurun_out/ncu.log 2>&1
echo "ncu eager rc=$?"
bash scripts/gpu_ncu_ops.sh group group_vec4 group_sa2 1
bash scripts/gpu_ncu_ops.sh interp interp_vec4 interp_fp4 0
bash scripts/gpu_ncu_ops.sh grads "csr_" grads_fp4 0
