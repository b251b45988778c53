#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','with_attention_layers','config3_training_step','config4_whole_scene','attention_layer_tcgen05')})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
# launch list of the bench command (graph replay), then full captures of the new kernels
timeout 300 python bench.py --steps 8 --warmup 3 --skip-cpu --skip-probe --train 0 --scenes 0 --attention-layers 0 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --skip-cpu --skip-probe --train 0 --scenes 0 --attention-layers 0 > gpurun_out/ncu.log 2>&1
echo "ncu list rc=$?"
timeout 300 python scripts/dev/time_attlayer.py > gpurun_out/attlayer.txt 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_layer_wide_kernel -c 3 -f -o gpurun_out/prof_attwide python scripts/dev/time_attlayer.py > gpurun_out/ncu_attwide.log 2>&1
echo "ncu attwide rc=$?"; cat gpurun_out/attlayer.txt
