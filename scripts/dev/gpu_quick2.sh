#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q -k "fps" 2>&1 | tail -6
python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
import pcops_b200 as ops
from pcops_b200 import synth
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(it):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
for n,m,b in [(16384,1024,64),(16384,4096,64),(65536,1024,64),(65536,4096,64),(131072,4096,16),(262144,1024,8)]:
    x=torch.rand(b,n,3,device='cuda')
    ms=t(lambda: ops.farthest_point_sample(m,x), 3)
    print("fps n=%d m=%d B=%d: %.2f ms  (%.1f us/scene, %.0f cycles/round-equivalent)"%(n,m,b,ms,ms*1e3/b, ms*1e-3*1.965e9/(m-1)/max(1,(b*((n+8191)//8192 if n<=131072 else 1)+147)//148)))
PY
