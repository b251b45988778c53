#!/bin/bash
# A/B of attention-kernel variants: value of the pipelined step AND the lone-launch times of the four attention levels
PKG=pointcloud-segmentation-attention_b200
for lib in "$@"; do
  name=$(basename $lib .so)
  timeout 300 python bench.py --lib $PKG/$lib --skip-cpu --scenes 0 --config5 0 --steps 200 --full-model 0 --train 0 --attention-layers 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['rooflines']
print('$name value %.0f e2e %.0f lone attention us:' % (d['value'], d['e2e']['value']), [round(r['attention_sa%d'%i]['ms']*1e3,1) for i in (1,2,3,4)])"
done
