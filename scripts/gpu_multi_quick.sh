#!/bin/bash
# Headline regions only on N GPUs, several configurations (rank spread, concurrency hint, batches in flight)
N=${1:-8}
F="--skip-cpu --config5 0 --full-model 0 --train 0 --attention-layers 0 --scenes 0 --skip-probe --steps 20 --warmup 5"
i=0
for extra in "" "--depth 8" "--hint 2 --depth 8" "--hint 1 --depth 8" ""; do
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+i)) bench.py --gpus $N $F $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('[$extra] value %.0f ms/step %.4f e2e %.0f spread' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['fastest_over_slowest_rank'])"
done
nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit,temperature.gpu --format=csv,noheader
