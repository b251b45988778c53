#!/bin/bash
# config 4 only, several times (host-side variance of the threaded chunker)
for i in 1 2 3 4 5; do
timeout 300 python bench.py --skip-cpu --config5 0 --full-model 0 --train 0 --attention-layers 0 --skip-probe --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config4_whole_scene']
print('config4 %.1f scans/s  %.2f ms per scan' % (c['value'], c['ms_per_scan']))"
done
