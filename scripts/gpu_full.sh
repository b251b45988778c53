#!/bin/bash
# Full GPU check: parity tests, smoke, bench (own + reference arm), ncu launch list of the bench command.
# usage: gpurun --timeout 2400 -- 'bash scripts/gpu_full.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv,noheader > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
