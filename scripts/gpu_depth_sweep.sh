#!/bin/bash
# value of the headline region against the number of batches in flight (and without the attention contraction)
F="--skip-cpu --config5 0 --full-model 0 --train 0 --attention-layers 0 --skip-probe --scenes 0"
for d in 6 8 9 10 12; do
  timeout 300 python bench.py $F --depth $d 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth $d value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
timeout 300 python bench.py $F --depth 8 --attention 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth 8 no attention value', round(d['value']))"
timeout 300 python bench.py $F --depth 8 --fuse-layers 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth 8 fused layers value', round(d['value']), 'e2e', round(d['e2e']['value']))"
