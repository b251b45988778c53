#!/usr/bin/env python
"""Read-only HBM bandwidth reference: torch.sum over a 1 GiB fp32 tensor (read once, nothing written) next to the
attention contraction's K / V stream at the SA1 shape.  python scripts/ubench/read_bw.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pcops_b200 as ops  # noqa: E402


def ms(fn, n=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


x = torch.randn(256 * 1024 * 1024, device="cuda")
t = ms(lambda: x.sum())
print("torch.sum over 1 GiB: %.3f ms = %.0f GB/s read-only" % (t, x.numel() * 4 / t / 1e6))
y = torch.empty_like(x)
t = ms(lambda: y.copy_(x))
print("copy 1 GiB -> 1 GiB: %.3f ms = %.0f GB/s (read + write)" % (t, 2 * x.numel() * 4 / t / 1e6))
del x, y
G, S, C = 16384, 32, 64
Q = torch.randn(G, C, device="cuda")
sets = [(torch.randn(G, S, C, device="cuda"), torch.randn(G, S, C, device="cuda")) for _ in range(4)]
i = [0]


def att():
    K, V = sets[i[0] % 4]
    i[0] += 1
    return ops.attention_contract(Q, K, V, C // 4, 4)


t = ms(att, 20)
print("attention contraction SA1 (268 MB of K / V, rotating buffers): %.3f ms = %.0f GB/s" % (t, 2 * G * S * C * 4 / t / 1e6))
