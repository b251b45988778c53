#!/bin/bash
# A/B of three_nn variants: step value and lone three_nn launches (cell-grid path, B = 16)
PKG=pointcloud-segmentation-attention_b200
for lib in "$@"; do
  name=$(basename $lib .so)
  timeout 300 python bench.py --lib $PKG/$lib --skip-cpu --scenes 0 --config5 0 --steps 200 --full-model 0 --train 0 --attention-layers 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
g=d['grid_variants_ms']
print('$name value %.0f e2e %.0f lone three_nn us:' % (d['value'], d['e2e']['value']), [round(g['three_nn_fp%d'%i]*1e3,1) for i in (1,2,3,4)])"
done
