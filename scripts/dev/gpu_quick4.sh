#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "ball or grid or pipeline or sample_and_group or full_size" 2>&1 | tail -3
timeout 300 python scripts/opbench.py --ops ball --levels 0,1 2>&1 | grep -v "^{"
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import pcops_b200 as ops
from pcops_b200 import synth, tf_grouping
x,_=synth.scannet_batch(0,16,8192); x=torch.from_numpy(x).cuda()
_,nx=ops.farthest_point_sample_and_gather(1024,x)
def t(fn,it=10):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(it):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    return sorted(ts)[len(ts)//2]
for g in (True, False):
    tf_grouping.USE_GRID=g
    print("ball SA1 grid=%s: %.1f us"%(g, t(lambda: ops.query_ball_point(0.1,32,x,nx))))
PY
bash scripts/dev/gpu_ab7.sh
