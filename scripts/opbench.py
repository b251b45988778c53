#!/usr/bin/env python
"""Per-op timings of libpcops.so on one GPU (development aid; bench.py is the contract).

    python scripts/opbench.py [--ops fps,ball,...] [--batch 16] [--iters 20] [--ref 1]

Each op is timed alone with CUDA events on the current stream, median of --iters, with a 256 MB write between
iterations to flush L2.  --ref 1 also times the reference's own CUDA kernels (oracle/_ref/libref_gpu.so, compiled
unmodified with the reference's flags) where they exist: the GPU "kernel to beat".
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pcops_b200 as ops  # noqa: E402
from pcops_b200 import synth  # noqa: E402
from pcops_b200.pipeline import SA_LEVELS  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="fps,gather,ball,group,attention,three_nn,interp,grads,knn")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ref", type=int, default=0)
    ap.add_argument("--levels", default="0,1,2,3")
    args = ap.parse_args()
    want = set(args.ops.split(","))
    B = args.batch
    dev = torch.device("cuda")
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    xyz_np, feat_np = synth.scannet_batch(0, B, 8192)
    xyz0, feat0 = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(feat_np).to(dev)
    refgpu = None
    if args.ref:
        from oracle import ref
        if ref.available_gpu(nofma=False):
            refgpu = ref.Gpu(nofma=False)
    out = {}

    def rec(name, fn, ref_fn=None):
        med, mn = timeit(fn, args.iters, flush)
        out[name] = {"us_median": round(med, 2), "us_min": round(mn, 2)}
        line = "%-28s %9.1f us (min %9.1f)" % (name, med, mn)
        if ref_fn is not None and refgpu is not None:
            rmed, rmn = timeit(ref_fn, max(3, args.iters // 4), flush)
            out[name]["ref_us_median"] = round(rmed, 2)
            line += "   reference kernel %10.1f us  -> %.1fx" % (rmed, rmed / med)
        print(line, flush=True)

    levels = [int(x) for x in args.levels.split(",")]
    xyz, feat = xyz0, feat0
    g = torch.Generator(device=dev).manual_seed(0)
    for li, (m, r, ns, cout) in enumerate(SA_LEVELS):
        n, cin = xyz.shape[1], feat.shape[2]
        tag = "_sa%d" % (li + 1)
        fi = ops.farthest_point_sample(m, xyz)
        new_xyz = ops.gather_point(xyz, fi)
        idx, cnt = ops.query_ball_point(r, ns, xyz, new_xyz)
        if li in levels:
            if "fps" in want:
                temp = torch.empty((32, n), dtype=torch.float32, device=dev)
                o = torch.empty((B, m), dtype=torch.int32, device=dev)
                rec("fps" + tag, lambda: ops.farthest_point_sample(m, xyz),
                    (lambda: refgpu.fps_launch(m, xyz, temp, o)) if refgpu else None)
                if li == 0:
                    for bb in (1, 4, 64, 148, 296):
                        xb = xyz0[:1].expand(bb, -1, -1).contiguous() if bb > B else xyz0[:bb].contiguous()
                        rec("fps_sa1_B%d" % bb, lambda: ops.farthest_point_sample(m, xb))
            if "gather" in want:
                rec("gather" + tag, lambda: ops.gather_point(xyz, fi))
            if "ball" in want:
                i2, c2 = torch.empty_like(idx), torch.empty_like(cnt)
                rec("query_ball" + tag, lambda: ops.query_ball_point(r, ns, xyz, new_xyz),
                    (lambda: refgpu.ball_launch(r, ns, xyz, new_xyz, i2, c2)) if refgpu else None)
            if "group" in want:
                go = torch.empty((B, m, ns, cin), dtype=torch.float32, device=dev)
                rec("group_xyz" + tag, lambda: ops.group_point(xyz, idx))
                rec("group_feat" + tag + "_c%d" % cin, lambda: ops.group_point(feat, idx),
                    (lambda: refgpu.group_launch(feat, idx, go)) if refgpu else None)
            if "grads" in want:
                gg = torch.randn((B, m, ns, cin), generator=g, device=dev)
                rec("group_grad" + tag + "_c%d" % cin, lambda: ops.group_point_grad(feat, idx, gg))
            if "attention" in want:
                Q = torch.randn((B * m, cout), generator=g, device=dev)
                K = torch.randn((B * m, ns, cout), generator=g, device=dev)
                V = torch.randn((B * m, ns, cout), generator=g, device=dev)
                rec("attention" + tag, lambda: ops.attention_contract(Q, K, V, cout // 4, 4))
            fpc = {0: 128, 1: 256, 2: 256, 3: 512}[li]
            p2 = torch.randn((B, m, fpc), generator=g, device=dev)
            if "three_nn" in want:
                rec("three_nn_fp%d" % (4 - li), lambda: ops.three_nn(xyz, new_xyz))
            if "interp" in want:
                d3, i3 = ops.three_nn(xyz, new_xyz)
                w = ops.three_weights(d3)
                rec("three_interp_fp%d_c%d" % (4 - li, fpc), lambda: ops.three_interpolate(p2, i3, w))
                if "grads" in want:
                    g2 = torch.randn((B, n, fpc), generator=g, device=dev)
                    rec("three_interp_grad_fp%d" % (4 - li), lambda: ops.three_interpolate_grad(p2, i3, w, g2))
            if "knn" in want and li <= 1:
                rec("knn_k32" + tag, lambda: ops.knn_point(32, xyz, new_xyz))
        xyz = new_xyz
        feat = torch.randn((B, m, cout), generator=g, device=dev)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
