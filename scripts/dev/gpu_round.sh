#!/bin/bash
# One GPU-box session: parity tests, bench, then (only if the bench exited 0) the ncu launch list of the same command.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
timeout 300 python bench.py --steps 50 --warmup 5 --graph 1 --skip-cpu > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench graph rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-probe > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu --skip-probe > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
