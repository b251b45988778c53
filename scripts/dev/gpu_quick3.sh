#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k pipeline 2>&1 | tail -3
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import pcops_b200 as ops
from pcops_b200 import synth, pointnet_util
def t(fn, it=10):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(it):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    return sorted(ts)[len(ts)//2]
x,f=synth.scannet_batch(0,16,8192); x=torch.from_numpy(x).cuda(); f=torch.from_numpy(f).cuda()
g=torch.Generator(device='cuda').manual_seed(0)
cases=[("SA1 c=6",x,f,1024,0.1)]
_,nx=ops.farthest_point_sample_and_gather(1024,x)
cases.append(("SA2 c=64",nx,torch.randn(16,1024,64,generator=g,device='cuda'),256,0.2))
_,nx2=ops.farthest_point_sample_and_gather(256,nx)
cases.append(("SA3 c=128",nx2,torch.randn(16,256,128,generator=g,device='cuda'),64,0.4))
for name,xyz,feat,m,r in cases:
    _,new_xyz=ops.farthest_point_sample_and_gather(m,xyz); idx,_=ops.query_ball_point(r,32,xyz,new_xyz)
    def fused(): return pointnet_util._SAGroup.apply(xyz,feat,idx,new_xyz)
    def comp():
        gx=ops.group_point(xyz,idx)-new_xyz.unsqueeze(2); gp=ops.group_point(feat,idx); return torch.cat([gx,gp],-1)
    print("%s grouping half: fused %.1f us, composition %.1f us"%(name,t(fused),t(comp)))
d,i3=ops.three_nn(x,ops.gather_point(x,ops.farthest_point_sample(1024,x)))
p2=torch.randn(16,1024,128,generator=g,device='cuda'); p1=torch.randn(16,8192,6,generator=g,device='cuda')
def fusedfp(): return pointnet_util._FPInterpolate.apply(d,i3,p2,p1)
def compfp(): w=ops.three_weights(d); return torch.cat([ops.three_interpolate(p2,i3,w),p1],2)
print("FP4 interpolation half (c2=128,c1=6): fused %.1f us, composition %.1f us"%(t(fusedfp),t(compfp)))
PY
