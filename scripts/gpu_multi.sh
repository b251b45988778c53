#!/bin/bash
# N-GPU scaling check with the driver's launch line for N > 1 (every optional region on).  usage:
#   gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_multi.sh N'
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -3 gpurun_out/bench_n$N.err
python scripts/show_bench.py gpurun_out/bench_n$N.json | grep -E "^value|summary|config4|config3|full_model" | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N rc=$?"
cut -c1-300 gpurun_out/bench_ref_n$N.json
