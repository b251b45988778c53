#!/usr/bin/env python
"""Launches every kernel the committed `ncu --set full` captures cover, once, between cudaProfilerStart/Stop (after an
untimed warm-up of the same call), at the bench shapes (B = 16 ScanNet-shaped chunks).

    ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_full \
        python scripts/ncu_targets.py [group ...]          groups: fps geom grads knn scene big (default: all)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import pcops_b200 as ops  # noqa: E402
from pcops_b200 import complete_scene_loader as csl  # noqa: E402
from pcops_b200 import synth  # noqa: E402

groups = set(sys.argv[1:]) or {"fps", "geom", "grads", "knn", "scene", "big", "dense", "att"}
dev = torch.device("cuda")
B = 16
xyz_np, feat_np = synth.scannet_batch(0, B, 8192)
xyz, feat = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(feat_np).to(dev)
g = torch.Generator(device=dev).manual_seed(0)


def profiled(fn):
    fn()                      # warm-up: attributes, code load, allocator
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    return out


fi, nx1 = ops.farthest_point_sample_and_gather(1024, xyz)
if "fps" in groups:          # fps_pruned_kernel<16,512> (SA1, the kernel `roofline` names), fps_onchip (SA2)
    profiled(lambda: ops.farthest_point_sample(1024, xyz))
    profiled(lambda: ops.farthest_point_sample(256, nx1))
if "geom" in groups:         # grid_build_kernel x4, ball_query_tile_kernel, three_nn_tile_kernel, gathers
    profiled(lambda: ops.query_ball_point(0.1, 32, xyz, nx1))
    profiled(lambda: ops.three_nn(xyz, nx1))
    idx, _ = ops.query_ball_point(0.1, 32, xyz, nx1)
    profiled(lambda: ops.group_point(feat, idx))
    d3, i3 = ops.three_nn(xyz, nx1)
    w = ops.three_weights(d3)
    p2 = torch.randn((B, 1024, 128), generator=g, device=dev)
    profiled(lambda: ops.three_interpolate(p2, i3, w))
    profiled(lambda: ops.sample_and_group(1024, 0.1, 32, xyz, feat))
if "grads" in groups:        # csr_build_kernel + csr_reduce_stream_kernel: FP4 interpolation gradient, SA2 group gradient
    d3, i3 = ops.three_nn(xyz, nx1)
    w = ops.three_weights(d3)
    p2 = torch.randn((B, 1024, 128), generator=g, device=dev)
    g2 = torch.randn((B, 8192, 128), generator=g, device=dev)
    profiled(lambda: ops.three_interpolate_grad(p2, i3, w, g2))
    _, nx2 = ops.farthest_point_sample_and_gather(256, nx1)
    i2, _ = ops.query_ball_point(0.2, 32, nx1, nx2)
    f2 = torch.randn((B, 1024, 64), generator=g, device=dev)
    gg = torch.randn((B, 256, 32, 64), generator=g, device=dev)
    profiled(lambda: ops.group_point_grad(f2, i2, gg))
if "knn" in groups:          # knn kernels at the SA1 shape
    profiled(lambda: ops.knn_point(32, xyz, nx1))
if "scene" in groups:        # the whole-scene chunker and map_back
    p, l, c, n = synth.whole_scene(1000)
    pts = torch.from_numpy(p).to(dev)
    f6 = torch.from_numpy(np.concatenate([c.astype(np.float32) / 255.0, n], 1)).to(dev)

    def scan():
        np.random.seed(5)
        ch = csl.chunk_scene(pts)
        ch.gather(f6)
        return csl.map_back(ch.point_sets.reshape(-1, 3), ch.orig_idx.reshape(-1), ch.masks.reshape(-1), (p.shape[0], 3))
    profiled(scan)
if "big" in groups:          # fps_cluster_kernel (8 CTAs per scene) and the cooperative grid beyond 262144 points
    xb = torch.rand((16, 65536, 3), generator=g, device=dev)
    profiled(lambda: ops.farthest_point_sample(256, xb))
    xc = torch.rand((2, 1 << 20, 3), generator=g, device=dev)
    profiled(lambda: ops.farthest_point_sample(64, xc))
if "dense" in groups:        # the tcgen05 Dense engine at a wide-row (SA1 layer 2) and a square (FP4) shape, and the weight gradient
    from pcops_b200 import sa_modules as sam
    for rows, K, N in ((B * 1024 * 32, 32, 32), (B * 8192, 128, 128), (B * 64 * 32, 259, 256)):
        xx = torch.randn((rows, K), generator=g, device=dev)
        img = sam.DenseImage(torch.randn((K, N), generator=g, device=dev), torch.randn(N, generator=g, device=dev))
        profiled(lambda: sam.dense(xx, img, True))
    xx = torch.randn((B * 1024 * 32, 64), generator=g, device=dev)
    dy = torch.randn((B * 1024 * 32, 64), generator=g, device=dev)
    profiled(lambda: sam.dense_weight_grad(xx, dy))
if "att" in groups:          # attention contraction (lane = head, bulk-copy staging) and its backward at SA1
    Q = torch.randn((B * 1024, 64), generator=g, device=dev)
    K_ = torch.randn((B * 1024, 32, 64), generator=g, device=dev)
    V_ = torch.randn((B * 1024, 32, 64), generator=g, device=dev)
    profiled(lambda: ops.attention_contract(Q, K_, V_, 16, 4))
print("done", sorted(groups))
