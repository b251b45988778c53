"""The whole inference forward of the reference's attention model on this library's kernels, pre-allocated.

``ScanNetAttentionModel`` runs attention_points/models/pointnet2_sem_seg_attention.py:28-62 for a batch of B chunks of N
points: four ``pointnet_sa_module_attention`` levels (sample_and_group -> shared MLP -> AttentionLayer -> batch norm),
four ``pointnet_fp_module`` levels (three_nn -> weights -> three_interpolate -> concat -> MLP), fc1, fc2 -- every
tensor produced by the layer before it (no stand-ins), every launch one of libpcops.so's kernels:

  geometry        pc_fps_gather, pc_query_ball_grid, pc_sa_group, pc_three_nn_grid, pc_fp_interpolate
  dense layers    pc_dense_fwd (tcgen05, 3xTF32; inference batch norm folded into kernel and bias)
  attention       pc_attention_layer_fwd_prepared (tcgen05; the batch norm after the layer is folded into the value
                  projection: softmax weights sum to 1, so s * sum(a V) + t = sum(a (s V + t)))

58 launches per forward, CUDA-graph capturable (no allocation, no host sync).  Weights are seeded random tensors in
TensorFlow's layouts ((cin, cout) kernels); ``reference_forward`` of tests/ restates the same graph in float64.
The FPS chain runs on the instance's main stream, everything else follows it on a side stream.
"""
import ctypes

import torch

from . import _lib

SA_SPEC = ((1024, 0.1, 32, (32, 32, 64)), (256, 0.2, 32, (64, 64, 128)), (64, 0.4, 32, (128, 128, 256)),
           (16, 0.8, 32, (256, 256, 512)))                       # pointnet2_sem_seg_attention.py:28-43
FP_SPEC = ((256, 256), (256, 256), (256, 128), (128, 128, 128))   # fa_layer1..4, :46-53
NUM_CLASS = 21


class _Dense:
    """One folded Dense / 1x1-conv layer: weights (cin, cout), bias (cout), tensor-core image prepared once."""

    def __init__(self, L, cin, cout, gen, dev, relu=True):
        self.cin, self.cout, self.relu = cin, cout, relu
        self.w = torch.randn(cin, cout, generator=gen, device=dev) / cin ** 0.5
        self.b = torch.randn(cout, generator=gen, device=dev) * 0.1
        self.image = torch.empty(L.pc_dense_image_bytes(cin, cout), dtype=torch.uint8, device=dev)
        _c(L.pc_dense_prepare(cin, cout, _lib.ptr(self.w), 0, _lib.ptr(self.image), _lib.stream()))


class ScanNetAttentionModel:
    def __init__(self, batch, npoints=8192, feat_channels=6, device="cuda", seed=0):
        self.B, self.N, self.CF = batch, npoints, feat_channels
        self.dev = dev = torch.device(device)
        self.L = L = _lib.lib()
        f32, i32 = torch.float32, torch.int32
        g = torch.Generator(device=dev).manual_seed(seed)
        B = batch

        def buf(*shape, dtype=f32):
            return torch.empty(shape, dtype=dtype, device=dev)
        with torch.cuda.device(dev):
            self.xyz0 = torch.zeros((B, npoints, 3), dtype=f32, device=dev)
            self.feat0 = torch.zeros((B, npoints, feat_channels), dtype=f32, device=dev) if feat_channels else None
            self.sa = []
            n, cin, xyz, feat = npoints, feat_channels, self.xyz0, self.feat0
            for (m, r, ns, mlp) in SA_SPEC:
                lv = dict(n=n, m=m, r=r, ns=ns, cin=cin, cout=mlp[-1], xyz=xyz, feat=feat)
                lv["fps_idx"], lv["new_xyz"] = buf(B, m, dtype=i32), buf(B, m, 3)
                lv["idx"], lv["cnt"] = buf(B, m, ns, dtype=i32), buf(B, m, dtype=i32)
                lv["new_points"], lv["gxyz"] = buf(B, m, ns, 3 + cin), buf(B, m, ns, 3)
                lv["fps_ws"] = _lib.workspace(L.pc_fps_workspace_bytes(B, n, m), dev)
                lv["ball_ws"] = _lib.workspace(L.pc_query_ball_grid_workspace_bytes(B, n, m), dev)
                lv["mlp"], c = [], 3 + cin
                for cout in mlp:
                    lv["mlp"].append(_Dense(L, c, cout, g, dev))
                    c = cout
                rows = B * m * ns
                lv["h"] = [buf(rows, d.cout) for d in lv["mlp"]]
                C = mlp[-1]
                lv["W"] = [torch.randn(C, C, generator=g, device=dev) / C ** 0.5 for _ in range(3)]   # Dense q, k, v (in, out)
                lv["bq"] = [torch.randn(C, generator=g, device=dev) * 0.1 for _ in range(3)]
                lv["att_ws"] = _lib.workspace(L.pc_attention_layer_workspace_bytes(B * m, ns, C), dev)
                _c(L.pc_attention_layer_prepare(ns, C, _lib.ptr(lv["W"][0]), _lib.ptr(lv["bq"][0]), _lib.ptr(lv["W"][1]),
                                                _lib.ptr(lv["bq"][1]), _lib.ptr(lv["W"][2]), _lib.ptr(lv["bq"][2]),
                                                _lib.ptr(lv["att_ws"]), _lib.stream()))
                lv["out"] = buf(B, m, C)
                self.sa.append(lv)
                n, cin, xyz, feat = m, C, lv["new_xyz"], lv["out"]
            self.fp = []
            points2 = self.sa[3]["out"]
            for k, mlp in enumerate(FP_SPEC):
                lv = self.sa[3 - k]                      # FP level k interpolates from lv's output cloud onto its input cloud
                points1 = lv["feat"]                     # skip features on the dense cloud (None for xyz-only level 0)
                c1 = 0 if points1 is None else points1.shape[2]
                c2 = points2.shape[2]
                fp = dict(n=lv["n"], m=lv["m"], c1=c1, c2=c2, xyz1=lv["xyz"], xyz2=lv["new_xyz"], points1=points1, points2=points2)
                fp["dist"], fp["idx"], fp["w"] = buf(B, lv["n"], 3), buf(B, lv["n"], 3, dtype=i32), buf(B, lv["n"], 3)
                fp["cat"] = buf(B, lv["n"], c2 + c1)
                fp["nn_ws"] = _lib.workspace(L.pc_three_nn_grid_workspace_bytes(B, lv["n"], lv["m"]), dev)
                fp["mlp"], c = [], c2 + c1
                for cout in mlp:
                    fp["mlp"].append(_Dense(L, c, cout, g, dev))
                    c = cout
                fp["h"] = [buf(B * lv["n"], d.cout) for d in fp["mlp"]]
                self.fp.append(fp)
                points2 = fp["h"][-1].view(B, lv["n"], c)
            self.fc1 = _Dense(L, 128, 128, g, dev)                       # conv1d + bn + relu (:56-57)
            self.fc2 = _Dense(L, 128, NUM_CLASS, g, dev, relu=False)     # conv1d, activation_fn=None (:60)
            self.net = buf(B * npoints, 128)
            self.logits = buf(B, npoints, NUM_CLASS)
            self.main = torch.cuda.Stream(device=dev, priority=-1)
            self.side = torch.cuda.Stream(device=dev)
            torch.cuda.synchronize(dev)
        self.launches_per_step = sum(4 + len(lv["mlp"]) + 2 for lv in self.sa) + \
            sum((2 if fp["m"] >= 64 else 1) + 1 + len(fp["mlp"]) for fp in self.fp) + 2
        self._graph = None

    def set_inputs(self, xyz, feats=None, non_blocking=True):
        with torch.cuda.stream(self.main):
            self.xyz0.copy_(xyz, non_blocking=non_blocking)
            if self.feat0 is not None:
                self.feat0.copy_(feats, non_blocking=non_blocking)

    def dense_flops(self):
        """2 * rows * cin * cout over every Dense / conv layer and attention projection of one forward."""
        f = 0
        for lv in self.sa:
            rows = self.B * lv["m"] * lv["ns"]
            f += sum(2 * rows * d.cin * d.cout for d in lv["mlp"])
            f += 2 * rows * lv["cout"] * lv["cout"] * 2 + 2 * self.B * lv["m"] * lv["cout"] * lv["cout"]
        for fp in self.fp:
            f += sum(2 * self.B * fp["n"] * d.cin * d.cout for d in fp["mlp"])
        f += 2 * self.B * self.N * 128 * (128 + NUM_CLASS)
        return f

    def _dense(self, d, x, rows, out, st):
        _c(self.L.pc_dense_fwd(rows, d.cin, d.cout, _lib.ptr(x), d.cin, _lib.ptr(d.image), _lib.ptr(d.b), 1 if d.relu else 0,
                               _lib.ptr(out), d.cout, st))

    def forward(self, main=None, side=None):
        L, B, p = self.L, self.B, _lib.ptr
        main = main or self.main
        side = side or self.side
        s_main, s_side = ctypes.c_void_p(main.cuda_stream), ctypes.c_void_p(side.cuda_stream)
        side.wait_stream(main)          # inputs (copied on main) before anything on the side stream reads them
        for lv in self.sa:
            n, m, ns, cin, C = lv["n"], lv["m"], lv["ns"], lv["cin"], lv["cout"]
            _c(L.pc_fps_gather(B, n, m, p(lv["xyz"]), p(lv["fps_ws"]), p(lv["fps_idx"]), p(lv["new_xyz"]), s_main))
            side.wait_stream(main)
            _c(L.pc_query_ball_grid(B, n, m, lv["r"], ns, p(lv["xyz"]), p(lv["new_xyz"]), p(lv["idx"]), p(lv["cnt"]),
                                    p(lv["ball_ws"]), s_side))
            _c(L.pc_sa_group(B, n, cin, m, ns, p(lv["xyz"]), p(lv["feat"]), p(lv["idx"]), p(lv["new_xyz"]),
                             p(lv["new_points"]), p(lv["gxyz"]), s_side))
            rows, x = B * m * ns, lv["new_points"]
            for d, h in zip(lv["mlp"], lv["h"]):
                self._dense(d, x, rows, h, s_side)
                x = h
            W, b = lv["W"], lv["bq"]
            _c(L.pc_attention_layer_fwd_prepared(B * m, ns, C, p(x), ns * C, p(x), p(W[0]), p(b[0]), p(W[1]), p(b[1]),
                                                 p(W[2]), p(b[2]), p(lv["out"]), p(lv["att_ws"]), s_side))
        for fp in self.fp:
            n, m = fp["n"], fp["m"]
            _c(L.pc_three_nn_grid(B, n, m, p(fp["xyz1"]), p(fp["xyz2"]), p(fp["dist"]), p(fp["idx"]), p(fp["nn_ws"]), s_side))
            _c(L.pc_fp_interpolate(B, n, m, fp["c2"], fp["c1"], p(fp["dist"]), p(fp["idx"]), p(fp["points2"]),
                                   p(fp["points1"]), p(fp["cat"]), p(fp["w"]), s_side))
            x = fp["cat"]
            for d, h in zip(fp["mlp"], fp["h"]):
                self._dense(d, x, B * n, h, s_side)
                x = h
        self._dense(self.fc1, x, B * self.N, self.net, s_side)
        self._dense(self.fc2, self.net, B * self.N, self.logits, s_side)
        main.wait_stream(side)

    def capture(self):
        self.forward()
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.main):
            self.forward(main=torch.cuda.current_stream(self.dev))
        self._graph = g
        return g

    def replay(self):
        with torch.cuda.stream(self.main):
            self._graph.replay()


def _c(rc):
    if rc != 0:
        _lib.check(rc, "pcops model call")
