"""Fused tcgen05 AttentionLayer vs its fp32 cuBLAS composition at the four ScanNet attention levels (B=16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pcops_b200  # noqa
from pcops_b200.attention_layer import attention_contract, attention_layer_fused

torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(5)
for G, C in ((16384, 64), (4096, 128), (1024, 256), (256, 512)):
    S = 32
    x = torch.randn(G, S, C, generator=g, device="cuda")
    xq = x[:, 0, :].contiguous()
    W = [torch.randn(C, C, generator=g, device="cuda") / C ** 0.5 for _ in range(3)]
    b = [torch.randn(C, generator=g, device="cuda") * 0.1 for _ in range(3)]
    comp = lambda: attention_contract(xq @ W[0] + b[0], x @ W[1] + b[1], x @ W[2] + b[2], C // 4, 4)
    fused = lambda: attention_layer_fused(xq, x, W[0], b[0], W[1], b[1], W[2], b[2])

    def ms(fn):
        fn(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]
    a, c = fused(), comp()
    err = float(((a - c).abs().max() / c.abs().max()).item())
    tf, tc = ms(fused), ms(comp)
    flops = 3 * 2.0 * G * S * C * 2 * C
    print("C=%d G=%d fused %.1f us  composition %.1f us  speedup %.1fx  rel err %.2e  issued tf32 %.0f TFLOP/s"
          % (C, G, tf * 1e3, tc * 1e3, tc / tf, err, flops / (tf * 1e-3) / 1e12), flush=True)
