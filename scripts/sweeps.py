#!/usr/bin/env python
"""SURVEY.md 8(d) tables that bench.py's one line does not carry: FPS latency / throughput against batch size and the
config-5 geometry-op scaling sweep (FPS / ball query / kNN at N = 16k .. 256k points).  Writes markdown to stdout.

    python scripts/sweeps.py > profiles/<name>.md        (one GPU)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pcops_b200 as ops  # noqa: E402
from pcops_b200 import synth  # noqa: E402

FLUSH = None


def timeit(fn, iters=5):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        FLUSH.fill_(0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda"
    print("# Sweeps on %s\n" % torch.cuda.get_device_name(0))
    xyz_np, _ = synth.scannet_batch(0, 16, 8192)
    base = torch.from_numpy(xyz_np).to(dev)
    print("## FPS, 8192 -> 1024 (SA1), against batch size (synthetic ScanNet-shaped chunks)\n")
    print("| B | ms per launch | us per scene | SMs busy |")
    print("|---|---|---|---|")
    for B in (1, 4, 16, 64, 148, 296, 592):
        x = base.repeat((B + 15) // 16, 1, 1)[:B].contiguous()
        ms = timeit(lambda: ops.farthest_point_sample(1024, x))
        print("| %d | %.3f | %.2f | %d |" % (B, ms, ms * 1e3 / B, min(B, 148)))
    print("\n## Config 5: geometry-op scaling (uniform clouds, r chosen for ~32 expected neighbours)\n")
    print("| op | N | npoint | B | ms | us per scene | path |")
    print("|---|---|---|---|---|---|---|")
    g = torch.Generator(device=dev).manual_seed(1)
    for N, B in ((16384, 64), (65536, 64), (262144, 8), (300000, 8)):
        x = torch.rand((B, N, 3), generator=g, device=dev)
        for m in (1024, 4096):
            if N >= 262144 and m > 1024:
                continue
            path = "cluster of %d CTAs (DSMEM)" % min(16, max(2, 1 << ((N + 8191) // 8192 - 1).bit_length())) \
                if N <= 262144 else "streamed min-distances (L2)"
            ms = timeit(lambda: ops.farthest_point_sample(m, x), 3)
            print("| FPS | %d | %d | %d | %.2f | %.1f | %s |" % (N, m, B, ms, ms * 1e3 / B, path))
            if N > 65536:
                continue
            idx = ops.farthest_point_sample(m, x)
            q = ops.gather_point(x, idx)
            r = (32.0 / N * 3.0 / (4.0 * 3.14159265)) ** (1.0 / 3.0)
            ms = timeit(lambda: ops.query_ball_point(r, 32, x, q), 3)
            print("| ball query r=%.3f | %d | %d | %d | %.2f | %.1f | %s |"
                  % (r, N, m, B, ms, ms * 1e3 / B, "cell grid" if N <= 15872 else "all pairs"))
            if m <= 1024:
                ms = timeit(lambda: ops.knn_point(32, x, q), 3)
                print("| kNN k=32 (fused) | %d | %d | %d | %.2f | %.1f | all pairs, no (b,m,n) matrix |"
                      % (N, m, B, ms, ms * 1e3 / B))
    print("\n## Backward ops (config 3 shapes, B=16)\n")
    print("| op | shape | ms |")
    print("|---|---|---|")
    feats = torch.rand((16, 1024, 64), generator=g, device=dev)
    x1 = base
    idx1 = ops.farthest_point_sample(1024, x1)
    nx1 = ops.gather_point(x1, idx1)
    idx2 = ops.farthest_point_sample(256, nx1)
    nx2 = ops.gather_point(nx1, idx2)
    bi, _ = ops.query_ball_point(0.2, 32, nx1, nx2)
    go = torch.rand((16, 256, 32, 64), generator=g, device=dev)
    ms = timeit(lambda: ops.group_point_grad(feats, bi, go))
    print("| GroupPointGrad SA2 | grad_out (16,256,32,64) -> (16,1024,64) | %.3f |" % ms)
    d3, i3 = ops.three_nn(x1, nx1)
    w = ops.three_weights(d3)
    p2 = torch.rand((16, 1024, 128), generator=g, device=dev)
    g2 = torch.rand((16, 8192, 128), generator=g, device=dev)
    ms = timeit(lambda: ops.three_interpolate_grad(p2, i3, w, g2))
    print("| ThreeInterpolateGrad FP4 | grad_out (16,8192,128) -> (16,1024,128) | %.3f |" % ms)
    Q = torch.rand((16 * 1024, 64), generator=g, device=dev)
    K = torch.rand((16 * 1024, 32, 64), generator=g, device=dev)
    V = torch.rand((16 * 1024, 32, 64), generator=g, device=dev)
    do = torch.rand((16 * 1024, 64), generator=g, device=dev)
    L = ops._lib.lib()
    dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
    p = ops._lib.ptr
    ms = timeit(lambda: L.pc_attention_bwd(16 * 1024, 32, 16, 4, p(Q), p(K), p(V), p(do), p(dQ), p(dK), p(dV),
                                           ops._lib.stream()))
    print("| attention contraction backward SA1 | K,V (16384,32,64) | %.3f |" % ms)


if __name__ == "__main__":
    main()
