#!/bin/bash
# A/B of three_interpolate variants: step value, lone-launch times at B = 16, steady-state fractions at B = 64
PKG=pointcloud-segmentation-attention_b200
for lib in "$@"; do
  name=$(basename $lib .so)
  timeout 300 python bench.py --lib $PKG/$lib --skip-cpu --scenes 0 --config5 0 --steps 200 --full-model 0 --train 0 --attention-layers 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['rooflines']; g=d['gathers_steady_state']
print('$name value %.0f e2e %.0f lone interp us:' % (d['value'], d['e2e']['value']), [round(r['three_interpolate_fp%d'%i]['ms']*1e3,1) for i in (1,2,3,4)], 'group_feat', [round(r['group_feat_sa%d'%i]['ms']*1e3,1) for i in (1,2,3,4)], 'steady', {k[:22]: round(v['frac'],3) for k,v in g.items()})"
done
