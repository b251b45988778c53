#!/bin/bash
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import pcops_b200 as ops
from pcops_b200.attention_layer import attention_layer_fused
def t(fn,it=10):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(it):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    return sorted(ts)[len(ts)//2]
G,S,C=16*1024,32,64
g=torch.Generator(device='cuda').manual_seed(0)
x=torch.randn(G,S,C,generator=g,device='cuda'); xq=x[:,0,:].contiguous()
W=[torch.randn(C,C,generator=g,device='cuda')/8 for _ in range(3)]; b=[torch.randn(C,generator=g,device='cuda')*0.1 for _ in range(3)]
torch.backends.cuda.matmul.allow_tf32=False
def comp():
    Q=xq@W[0]+b[0]; K=x@W[1]+b[1]; V=x@W[2]+b[2]
    return ops.attention_contract(Q,K,V,16,4)
def fused(): return attention_layer_fused(xq,x,W[0],b[0],W[1],b[1],W[2],b[2])
a=fused(); c=comp()
print("max rel err fused vs fp32 composition: %.2e"%((a-c).abs().max()/c.abs().max()).item())
print("SA1 attention layer (G=16384,S=32,C=64): fused tcgen05 %.1f us ; fp32 cuBLAS Dense x3 + contraction %.1f us"%(t(fused),t(comp)))
torch.backends.cuda.matmul.allow_tf32=True
print("   (composition with TF32 cuBLAS: %.1f us, max rel err %.2e)"%(t(comp), ((comp()-c).abs().max()/c.abs().max()).item()))
print("   contraction only on precomputed K,V: %.1f us"%t(lambda: ops.attention_contract(xq, x, x, 16, 4)))
flops=2*G*S*C*2*C*3
print("   tf32 MMA flops per call (3 splits): %.1f GFLOP"%(flops/1e9))
PY
