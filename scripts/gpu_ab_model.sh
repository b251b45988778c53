#!/bin/bash
# A/B of Dense-engine variants: dense + model parity tests, whole-model launch list (ncu, serialised) and throughput
PKG=pointcloud-segmentation-attention_b200
for lib in "$@"; do
  name=$(basename $lib .so)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/model_$name.csv python scripts/model_launches.py 16 $PKG/$lib > /dev/null 2>&1
  python scripts/print_launches.py gpurun_out/model_$name.csv | grep -E "dense|total" | awk '{print $NF, $(NF-1)}' | tr '\n' ' ' | sed "s/^/$name dense us: /"; echo
  timeout 300 python bench.py --lib $PKG/$lib --skip-cpu --skip-probe --scenes 0 --config5 0 --steps 100 --train 0 --attention-layers 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name full model %.0f scenes/s' % d['full_model_inference']['value'])"
done
