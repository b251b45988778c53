#!/bin/bash
# GPU-side time of the fused attention layer (3 launches captured in a CUDA graph) vs the fp32 composition
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import pcops_b200 as ops
from pcops_b200.attention_layer import attention_layer_fused
G,S,C=16*1024,32,64
g=torch.Generator(device='cuda').manual_seed(0)
x=torch.randn(G,S,C,generator=g,device='cuda'); xq=x[:,0,:].contiguous()
W=[torch.randn(C,C,generator=g,device='cuda')/8 for _ in range(3)]; b=[torch.randn(C,generator=g,device='cuda')*0.1 for _ in range(3)]
torch.backends.cuda.matmul.allow_tf32=False
def comp():
    Q=xq@W[0]+b[0]; K=x@W[1]+b[1]; V=x@W[2]+b[2]
    return ops.attention_contract(Q,K,V,16,4)
def fused(): return attention_layer_fused(xq,x,W[0],b[0],W[1],b[1],W[2],b[2])
def graph_time(fn, it=20):
    fn(); torch.cuda.synchronize()
    gr=torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): fn()
    gr.replay(); torch.cuda.synchronize(); ts=[]
    for _ in range(it):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    return sorted(ts)[len(ts)//2]
a=fused(); c=comp()
print("max rel err fused vs fp32 composition: %.2e"%((a-c).abs().max()/c.abs().max()).item())
tf=graph_time(fused); tc=graph_time(comp)
print("SA1 attention layer, GPU time (CUDA graph replay): fused tcgen05 %.1f us ; fp32 cuBLAS Dense x3 + contraction %.1f us ; %.1fx"%(tf,tc,tc/tf))
PY
