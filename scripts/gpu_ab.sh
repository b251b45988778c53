#!/bin/bash
# A/B of library variants on the headline regions only (no probes, no CPU arm, no config 4/5):
# usage: gpurun -- 'bash scripts/gpu_ab.sh libpcops.so libpcops_fps88.so ...'   (paths relative to the package)
mkdir -p gpurun_out
PKG=pointcloud-segmentation-attention_b200
for lib in "$@"; do
  name=$(basename $lib .so)
  timeout 300 python bench.py --lib $PKG/$lib --skip-probe --skip-cpu --scenes 0 --config5 0 --steps 200 ${AB_FLAGS} \
    > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$?"
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/ab_%s.json' % n).read().strip().splitlines()[-1])
    print(n, "value %.0f  e2e %.0f  train %.0f  layers %.0f  fps_sa1 %.1f us" % (
        d['value'], d['e2e']['value'], (d.get('config3_training_step') or {}).get('value', 0),
        (d.get('with_attention_layers') or {}).get('value', 0), d['fps_us_per_scene']['sa1_batch_latency_us']))
except Exception as e:
    print(n, "failed:", e)
PY
done
