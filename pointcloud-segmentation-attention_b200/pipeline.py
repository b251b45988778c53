"""The whole geometry hot path of one PointNet++ ScanNet forward, with every buffer pre-allocated.

``ScanNetGeometry`` runs, for a batch of B chunks of N points (xyz + 6 feature channels), every call the reference's
4 set-abstraction + 4 feature-propagation levels make into the custom-op libraries plus the
attention contraction (a13) of the four attention SA levels:

  SA level l (pointnet_util.py:34-52, attention_layer.py:225,255-261; hyper-parameters of
  attention_points/models/pointnet2_sem_seg_attention.py:28-43):
      FPS -> gather_point -> query_ball_point -> group_point(xyz) -> group_point(features) -> attention contraction
  FP level (pointnet_util.py:218-223):
      three_nn -> inverse-distance weights -> three_interpolate

The shared MLPs / Dense projections between those calls are stock dense layers and out of scope, so the tensors they
would produce (per-level point features, and Q/K/V of each attention level) are fixed synthetic stand-ins of the
right shape, resident in HBM.  Because of that, an FP level's ops are issued as soon as the xyz they depend on exists.

Streams: the four FPS calls form a strictly sequential chain (level l+1 samples level l's centroids) on the main
stream (high priority: FPS is latency-bound and occupies only B SMs); everything else of level l runs on side stream l
behind a per-level event, so grouping / interpolation / attention of all levels overlap each other and the FPS chain.
The sequence is CUDA-graph capturable (no allocation, no host sync inside).  Independent batches overlap further by
running several ``ScanNetGeometry`` instances, each on its own streams (bench.py --depth); tell the library how many
with ``_lib.set_concurrency_hint(instances)`` BEFORE ``capture()`` -- it then launches its streaming kernels as few
long-lived CTAs that co-reside with the other batches' kernels (grid sizes are baked into a captured graph).

All integer results a host consumer reads back (FPS, ball and three_nn indices, counts) are views into ONE contiguous
device buffer, so a step's result is a single device-to-host copy.
"""
import ctypes

import torch

from . import _lib

# (npoint, radius, nsample, mlp[-1]) per SA level -- pointnet2_sem_seg_attention.py:28-43
SA_LEVELS = ((1024, 0.1, 32, 64), (256, 0.2, 32, 128), (64, 0.4, 32, 256), (16, 0.8, 32, 512))
KEY_DIM = 4  # attention_layer.py:256-258: heads = C // 4, key_dim = output_dim = 4


class ScanNetGeometry:
    def __init__(self, batch, npoints=8192, feat_channels=6, device="cuda", attention=True, seed=0, own_streams=False,
                 grid=True, fuse_gather=True, fuse_layers=False, attention_layers=False, parts=("fps", "side")):
        self.B, self.N, self.CF = batch, npoints, feat_channels
        # diagnostics only: which parts of the forward to enqueue ("fps": the FPS chain, "side": everything else)
        self.parts = tuple(parts)
        self.skip = ()   # diagnostics only (scripts/ablate.py): op-name prefixes left out of the forward
        self.fuse_gather = fuse_gather  # pc_fps_gather instead of pc_fps + pc_gather_point
        # pc_sa_group (group xyz + centre + group features + concat) and pc_fp_interpolate (weights + interpolate)
        # instead of two GroupPoint calls / weights + ThreeInterpolate: what sample_and_group / pointnet_fp_module need.
        # Off by default here: it produces MORE than the reference-signature ops the benchmark counts (the concatenated
        # tensor on top of grouped_xyz); the wrappers in pointnet_util.py use it, where it replaces 4-5 launches.
        self.fuse_layers = fuse_layers
        self.fuse_fp = fuse_layers       # pc_fp_interpolate instead of three_weights + three_interpolate
        self.grid = grid  # cell-grid ball query / three_nn (same outputs as the all-pairs kernels, far fewer pair tests)
        self.dev = torch.device(device)
        self.attention = attention
        # attention_layers: run the WHOLE AttentionLayer (Dense Q/K/V + contraction, pc_attention_layer_fwd on tcgen05)
        # on stand-in grouped activations X instead of the contraction alone on stand-in K / V
        self.attention_layers = bool(attention and attention_layers)
        self.L = _lib.lib()
        # result arena: every tensor of result_tensors() is a view into it (4-byte elements, 16-byte aligned slots)
        sizes, n_ = [], npoints
        for (m_, _r, ns_, _c) in SA_LEVELS:
            sizes += [batch * m_, batch * m_ * ns_, batch * m_, batch * n_ * 3]
            n_ = m_
        self._arena = torch.empty(sum((x + 3) // 4 * 4 for x in sizes), dtype=torch.int32, device=self.dev)
        self._arena_off = 0

        def res(shape, dtype):
            numel = 1
            for d in shape:
                numel *= d
            v = self._arena[self._arena_off:self._arena_off + numel]
            self._arena_off += (numel + 3) // 4 * 4
            return v.view(dtype).view(*shape)
        f32, i32 = torch.float32, torch.int32
        dev = self.dev
        g = torch.Generator(device=dev).manual_seed(seed)

        def rnd(*shape):
            return torch.randn(*shape, generator=g, dtype=f32, device=dev)

        # Packed input arena (host batches, set_inputs_packed): [xyz f32 | normals f32 | colours u8], one H2D copy of
        # 27 bytes per point; xyz0 IS the arena's first section, feat0 is produced by pc_unpack_features
        # (colours / 255 | normals, train.py:95-98).  Device-resident callers keep using set_inputs().
        npts = batch * npoints
        self._in_arena = torch.zeros(npts * 27 + 16, dtype=torch.uint8, device=dev)
        self.xyz0 = self._in_arena[:npts * 12].view(f32).view(batch, npoints, 3)
        self._in_normals = self._in_arena[npts * 12:npts * 24].view(f32).view(batch, npoints, 3)
        self._in_colors = self._in_arena[npts * 24:npts * 27].view(batch, npoints, 3)
        self.feat0 = torch.zeros((batch, npoints, feat_channels), dtype=f32, device=dev)
        self.levels = []
        n, cin = npoints, feat_channels
        xyz = self.xyz0
        for (m, r, ns, cout) in SA_LEVELS:
            lv = dict(n=n, m=m, r=r, ns=ns, cin=cin, cout=cout, xyz=xyz)
            lv["feat"] = self.feat0 if n == npoints else rnd(batch, n, cin)  # stand-in for the previous level's MLP output
            lv["fps_idx"] = res((batch, m), i32)
            lv["new_xyz"] = torch.empty((batch, m, 3), dtype=f32, device=dev)
            lv["idx"] = res((batch, m, ns), i32)
            lv["cnt"] = res((batch, m), i32)
            lv["gxyz"] = torch.empty((batch, m, ns, 3), dtype=f32, device=dev)
            lv["gfeat"] = torch.empty((batch, m, ns, cin), dtype=f32, device=dev)
            if fuse_layers:
                lv["new_points"] = torch.empty((batch, m, ns, 3 + cin), dtype=f32, device=dev)
            if self.attention_layers:  # stand-in for the shared MLP's output (B,m,ns,cout) + the three Dense layers
                lv["X"] = rnd(batch * m, ns, cout)
                lv["XQ"] = lv["X"][:, 0, :].contiguous()          # query = sample 0 of the group (attention_layer.py:259)
                lv["W"] = [rnd(cout, cout) / cout ** 0.5 for _ in range(3)]
                lv["b"] = [rnd(cout) * 0.1 for _ in range(3)]
                lv["att"] = torch.empty((batch * m, cout), dtype=f32, device=dev)
                lv["att_ws"] = _lib.workspace(self.L.pc_attention_layer_workspace_bytes(batch * m, ns, cout), dev)
            elif attention:  # stand-ins for the Dense projections of the level's (B,m,ns,cout) activations
                lv["Q"] = rnd(batch * m, cout)
                lv["K"] = rnd(batch * m, ns, cout)
                lv["V"] = rnd(batch * m, ns, cout)
                lv["att"] = torch.empty((batch * m, cout), dtype=f32, device=dev)
            ws = self.L.pc_fps_workspace_bytes(batch, n, m)
            lv["fps_ws"] = _lib.workspace(ws, dev)
            lv["ball_ws"] = _lib.workspace(self.L.pc_query_ball_grid_workspace_bytes(batch, n, m), dev)
            self.levels.append(lv)
            xyz, n, cin = lv["new_xyz"], m, cout
        # FP levels: dense xyz = level input, sparse xyz = level output; channels of the sparse features
        # (pointnet2_sem_seg_attention.py:46-53): FP1 c=512, FP2 c=256, FP3 c=256, FP4 c=128
        self.fps = []
        for li, c in ((3, 512), (2, 256), (1, 256), (0, 128)):
            lv = self.levels[li]
            fp = dict(n=lv["n"], m=lv["m"], c=c, xyz1=lv["xyz"], xyz2=lv["new_xyz"], level=li)
            fp["points2"] = rnd(batch, lv["m"], c)  # stand-in for the deeper level's features
            fp["dist"] = torch.empty((batch, lv["n"], 3), dtype=f32, device=dev)
            fp["idx"] = res((batch, lv["n"], 3), i32)
            fp["w"] = torch.empty((batch, lv["n"], 3), dtype=f32, device=dev)
            fp["out"] = torch.empty((batch, lv["n"], c), dtype=f32, device=dev)
            fp["nn_ws"] = _lib.workspace(self.L.pc_three_nn_grid_workspace_bytes(batch, lv["n"], lv["m"]), dev)
            self.fps.append(fp)
        self.main = torch.cuda.Stream(device=dev, priority=-1) if own_streams else None
        self.sides = [torch.cuda.Stream(device=dev) for _ in self.levels]
        self.side = self.sides[0]
        # kernels per forward: FPS, gather, ball (+ ONE grid build binning both clouds), group x2, attention per SA level;
        # three_nn (+ one grid build when the known cloud has >= 64 points), weights, interpolate per FP level
        self.launches_per_step = len(self.levels) * ((6 if attention else 5) + (1 if grid else 0) -
                                                     (1 if fuse_gather else 0) - (1 if fuse_layers else 0)) + \
            sum(3 + (1 if grid and fp["m"] >= 64 else 0) - (1 if fuse_layers else 0) for fp in self.fps) + \
            (2 * len(self.levels) if self.attention_layers else 0)   # fused layer = operand prep + Q + main kernel
        self._graph = None
        self._arena16 = None
        self.training = False

    # ---- inputs / outputs ----------------------------------------------------------------------------------
    def set_inputs(self, xyz, feats, non_blocking=True):
        """Copy a batch (device or pinned host tensors) into the input buffers, on this pipeline's stream."""
        with torch.cuda.stream(self.stream()):
            self.xyz0.copy_(xyz, non_blocking=non_blocking)
            self.feat0.copy_(feats, non_blocking=non_blocking)

    def read_results(self, host_arena, non_blocking=True):
        """One device-to-host copy of the result arena into a pinned int32 host tensor, on this pipeline's stream."""
        with torch.cuda.stream(self.stream()):
            host_arena.copy_(self.result_arena(), non_blocking=non_blocking)

    def input_bytes(self):
        return self.xyz0.numel() * 4 + self.feat0.numel() * 4

    # ---- host batches: one packed copy in, one narrowed copy out -------------------------------------------------
    def packed_input_bytes(self):
        return self.B * self.N * 27

    @staticmethod
    def pack_host_batch(xyz, colors_u8, normals, pinned=True):
        """(B,N,3) f32 xyz, (B,N,3) u8 colours, (B,N,3) f32 normals -> ONE pinned uint8 host arena in the device
        arena's layout."""
        npts = xyz.shape[0] * xyz.shape[1]
        host = torch.empty(npts * 27, dtype=torch.uint8)
        if pinned:
            host = host.pin_memory()
        host[:npts * 12].view(torch.float32).copy_(torch.as_tensor(xyz).reshape(-1))
        host[npts * 12:npts * 24].view(torch.float32).copy_(torch.as_tensor(normals).reshape(-1))
        host[npts * 24:].copy_(torch.as_tensor(colors_u8).reshape(-1))
        return host

    def set_inputs_packed(self, host_arena, non_blocking=True):
        """ONE host-to-device copy of a pack_host_batch() arena, then the feature prologue on the device."""
        if self.CF != 6:
            raise ValueError("the packed input path carries colours + normals (6 feature channels)")
        st = self.stream()
        with torch.cuda.stream(st):
            self._in_arena[:host_arena.numel()].copy_(host_arena, non_blocking=non_blocking)
            _c(self.L.pc_unpack_features(self.B * self.N, _lib.ptr(self._in_colors), _lib.ptr(self._in_normals),
                                         _lib.ptr(self.feat0), ctypes.c_void_p(st.cuda_stream)))

    def result_bytes_u16(self):
        return self._arena_off * 2

    def read_results_u16(self, host_arena_u16, non_blocking=True):
        """The result arena narrowed to uint16 on the device (every index of this pipeline is < 8192, every count
        <= 32: lossless) and ONE device-to-host copy of half the bytes."""
        if self._arena16 is None:
            self._arena16 = torch.empty(self._arena_off, dtype=torch.int16, device=self.dev)
        st = self.stream()
        with torch.cuda.stream(st):
            _c(self.L.pc_narrow_indices_u16(self._arena_off, _lib.ptr(self._arena), _lib.ptr(self._arena16),
                                            ctypes.c_void_p(st.cuda_stream)))
            host_arena_u16.copy_(self._arena16, non_blocking=non_blocking)

    def capture_e2e(self, host_in, host_out_u16, overlap=True):
        """One CUDA graph for a whole host-to-host step: H2D of the packed batch, feature prologue, the forward,
        index narrowing, D2H -- a single graph launch per step instead of two copies, two launches and a replay."""
        if self.npoints_over_u16():
            raise ValueError("indices do not fit uint16")
        if self._arena16 is None:
            self._arena16 = torch.empty(self._arena_off, dtype=torch.int16, device=self.dev)
        self.set_inputs_packed(host_in)
        self.forward(overlap)
        self.read_results_u16(host_out_u16)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        saved, self.main = self.main, None
        try:
            with torch.cuda.graph(g, stream=saved):
                self.set_inputs_packed(host_in)
                self.forward(overlap)
                self.read_results_u16(host_out_u16)
        finally:
            self.main = saved
        return g

    def npoints_over_u16(self):
        return self.N > 65536

    def result_tensors(self):
        """The integer geometry decisions of a forward (what a host-side consumer reads back; together with the inputs
        they determine every gathered / interpolated tensor): FPS indices, ball indices and counts per SA level, three_nn
        indices per FP level."""
        out = []
        for lv in self.levels:
            out += [lv["fps_idx"], lv["idx"], lv["cnt"]]
        for fp in self.fps:
            out.append(fp["idx"])
        return out

    def result_arena(self):
        """The one contiguous int32 device buffer all result_tensors() live in."""
        return self._arena[:self._arena_off]

    def stream(self):
        return self.main if self.main is not None else torch.cuda.current_stream(self.dev)

    # ---- one forward -----------------------------------------------------------------------------------------
    def _sa_rest(self, lv, li, side, run):
        L, B, p = self.L, self.B, _lib.ptr
        st = ctypes.c_void_p(side.cuda_stream)
        n, m, ns, cin = lv["n"], lv["m"], lv["ns"], lv["cin"]
        if self.grid:
            run("query_ball_sa%d" % (li + 1), side, lambda: L.pc_query_ball_grid(
                B, n, m, lv["r"], ns, p(lv["xyz"]), p(lv["new_xyz"]), p(lv["idx"]), p(lv["cnt"]), p(lv["ball_ws"]), st))
        else:
            run("query_ball_sa%d" % (li + 1), side, lambda: L.pc_query_ball(
                B, n, m, lv["r"], ns, p(lv["xyz"]), p(lv["new_xyz"]), p(lv["idx"]), p(lv["cnt"]), st))
        if self.fuse_layers:
            run("sa_group_sa%d" % (li + 1), side, lambda: L.pc_sa_group(
                B, n, cin, m, ns, p(lv["xyz"]), p(lv["feat"]), p(lv["idx"]), p(lv["new_xyz"]), p(lv["new_points"]),
                p(lv["gxyz"]), st))
        else:
            run("group_xyz_sa%d" % (li + 1), side, lambda: L.pc_group_point(
                B, n, 3, m, ns, p(lv["xyz"]), p(lv["idx"]), p(lv["gxyz"]), st))
            run("group_feat_sa%d" % (li + 1), side, lambda: L.pc_group_point(
                B, n, cin, m, ns, p(lv["feat"]), p(lv["idx"]), p(lv["gfeat"]), st))
        if self.attention_layers:
            W, b = lv["W"], lv["b"]
            run("attention_sa%d" % (li + 1), side, lambda: L.pc_attention_layer_fwd(
                B * m, ns, lv["cout"], p(lv["XQ"]), p(lv["X"]), p(W[0]), p(b[0]), p(W[1]), p(b[1]), p(W[2]), p(b[2]),
                p(lv["att"]), p(lv["att_ws"]), st))
        elif self.attention:
            run("attention_sa%d" % (li + 1), side, lambda: L.pc_attention_fwd(
                B * m, ns, lv["cout"] // KEY_DIM, KEY_DIM, p(lv["Q"]), p(lv["K"]), p(lv["V"]), p(lv["att"]), st))

    def _fp(self, fp, side, run):
        L, B, p = self.L, self.B, _lib.ptr
        st = ctypes.c_void_p(side.cuda_stream)
        n, m, c = fp["n"], fp["m"], fp["c"]
        tag = "fp%d" % (4 - fp["level"])  # FP1 is the deepest level (pointnet2_sem_seg_attention.py:46-53)
        if self.grid:
            run("three_nn_" + tag, side, lambda: L.pc_three_nn_grid(
                B, n, m, p(fp["xyz1"]), p(fp["xyz2"]), p(fp["dist"]), p(fp["idx"]), p(fp["nn_ws"]), st))
        else:
            run("three_nn_" + tag, side, lambda: L.pc_three_nn(
                B, n, m, p(fp["xyz1"]), p(fp["xyz2"]), p(fp["dist"]), p(fp["idx"]), st))
        if self.fuse_fp:
            run("fp_interpolate_" + tag, side, lambda: L.pc_fp_interpolate(
                B, n, m, c, 0, p(fp["dist"]), p(fp["idx"]), p(fp["points2"]), None, p(fp["out"]), p(fp["w"]), st))
        else:
            run("three_weights_" + tag, side, lambda: L.pc_three_weights(B * n, p(fp["dist"]), p(fp["w"]), st))
            run("three_interpolate_" + tag, side, lambda: L.pc_three_interpolate(
                B, m, c, n, p(fp["points2"]), p(fp["idx"]), p(fp["w"]), p(fp["out"]), st))

    def forward(self, overlap=True, probes=None, train=False):
        """Enqueue one forward on the current stream (plus the side stream when overlap=True).  train=True (after
        allocate_backward()) also enqueues the level's gradient ops behind its forward ops: a training step.

        ``probes``: optional dict name -> list; for every op whose name is a key, a (start, end) pair of CUDA events
        recorded on the op's own stream around its launch is appended (bench.py reads kernel durations from them)."""
        L, B, p = self.L, self.B, _lib.ptr
        main = self.stream()
        s_main = ctypes.c_void_p(main.cuda_stream)
        fp_of = {fp["level"]: fp for fp in self.fps}

        def run(name, stream, call):
            if self.skip and name.startswith(self.skip):
                return
            if probes is not None and name in probes:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                _c(call())
                e1.record(stream)
                probes[name].append((e0, e1))
            else:
                _c(call())

        parts = self.parts
        for li, lv in enumerate(self.levels):
            side = self.sides[li] if overlap else main
            if "fps" in parts and self.fuse_gather:   # FPS writes the sampled coordinates itself: one launch, not two
                run("fps_sa%d" % (li + 1), main, lambda: L.pc_fps_gather(
                    B, lv["n"], lv["m"], p(lv["xyz"]), p(lv["fps_ws"]), p(lv["fps_idx"]), p(lv["new_xyz"]), s_main))
            elif "fps" in parts:
                run("fps_sa%d" % (li + 1), main, lambda: L.pc_fps(
                    B, lv["n"], lv["m"], p(lv["xyz"]), p(lv["fps_ws"]), p(lv["fps_idx"]), s_main))
                run("gather_sa%d" % (li + 1), main, lambda: L.pc_gather_point(
                    B, lv["n"], lv["m"], p(lv["xyz"]), p(lv["fps_idx"]), p(lv["new_xyz"]), s_main))
            if overlap:
                side.wait_stream(main)  # also orders this step's side work after the previous step's join
            if "side" in parts:
                self._sa_rest(lv, li, side, run)
                self._fp(fp_of[li], side, run)
                if train:
                    self._backward_level(lv, li, fp_of[li], side, run)
        if overlap:
            for side in self.sides:
                main.wait_stream(side)

    def op_names(self):
        names = []
        for li in range(len(self.levels)):
            names += ["fps_sa%d" % (li + 1)] + ([] if self.fuse_gather else ["gather_sa%d" % (li + 1)]) + \
                     ["query_ball_sa%d" % (li + 1)] + \
                     (["sa_group_sa%d" % (li + 1)] if self.fuse_layers else
                      ["group_xyz_sa%d" % (li + 1), "group_feat_sa%d" % (li + 1)])
            if self.attention:
                names.append("attention_sa%d" % (li + 1))
        for fp in self.fps:
            tag = "fp%d" % (4 - fp["level"])
            names += ["three_nn_" + tag] + (["fp_interpolate_" + tag] if self.fuse_layers else
                                              ["three_weights_" + tag, "three_interpolate_" + tag])
        return names

    def algorithmic_work(self):
        """Per-op algorithmic bytes (HBM-bound ops) or fp32 operations (compute-bound ops) of ONE forward over the
        whole batch, from the formulas of SURVEY.md 8(d) / DESIGN.md.  name -> dict(kind, amount)."""
        B, w = self.B, {}
        for li, lv in enumerate(self.levels):
            n, m, ns, cin, cout, t = lv["n"], lv["m"], lv["ns"], lv["cin"], lv["cout"], "_sa%d" % (li + 1)
            w["fps" + t] = dict(kind="fp32_ops", amount=B * (m - 1) * n * 10, bytes=B * (12 * n + 4 * m))
            w["gather" + t] = dict(kind="bytes", amount=B * (4 * m + 12 * m + 12 * m))
            w["query_ball" + t] = dict(kind="fp32_ops", amount=B * m * n * 11,  # upper bound: every pair tested
                                       bytes=B * (12 * n + 12 * m + 4 * m * ns + 4 * m))
            for nm, c in (("group_xyz", 3), ("group_feat", cin)):
                w[nm + t] = dict(kind="bytes", amount=B * (4 * m * ns + 4 * min(n, m * ns) * c + 4 * m * ns * c))
            w["sa_group" + t] = dict(kind="bytes", amount=B * (4 * m * ns + 12 * m + 4 * min(n, m * ns) * (3 + cin) +
                                                               4 * m * ns * (3 + cin) + 12 * m * ns))
            if self.attention:
                w["attention" + t] = dict(kind="bytes", amount=B * m * (2 * 4 * ns * cout + 2 * 4 * cout))
        for fp in self.fps:
            n, m, c, t = fp["n"], fp["m"], fp["c"], "_fp%d" % (4 - fp["level"])
            w["three_nn" + t] = dict(kind="fp32_ops", amount=B * n * m * 11, bytes=B * (12 * n + 12 * m + 24 * n))
            w["three_weights" + t] = dict(kind="bytes", amount=B * n * 24)
            w["three_interpolate" + t] = dict(kind="bytes", amount=B * (24 * n + 4 * m * c + 4 * n * c))
            w["fp_interpolate" + t] = dict(kind="bytes", amount=B * (24 * n + 4 * m * c + 4 * n * c + 12 * n))
        return w

    # ---- backward (config 3: a training step) ----------------------------------------------------------------
    def allocate_backward(self, seed=1234):
        """Buffers of the registered gradients a training step of the attention model with features runs
        (a6 GroupPointGrad at SA2-4 -- the SA1 features are network inputs and carry no gradient,
        a10 ThreeInterpolateGrad at FP1-4, the attention contraction's backward at the four attention levels).
        The upstream gradients the (out-of-scope) dense layers would deliver are fixed synthetic stand-ins."""
        if self.training:
            return
        B, dev, f32 = self.B, self.dev, torch.float32
        g = torch.Generator(device=dev).manual_seed(seed)

        def rnd(*shape):
            return torch.randn(*shape, generator=g, dtype=f32, device=dev)
        for li, lv in enumerate(self.levels):
            n, m, ns, cin, cout = lv["n"], lv["m"], lv["ns"], lv["cin"], lv["cout"]
            if self.attention:
                lv["d_att"] = rnd(B * m, cout)
                lv["dQ"] = torch.empty_like(lv["Q"])
                lv["dK"] = torch.empty_like(lv["K"])
                lv["dV"] = torch.empty_like(lv["V"])
            if li > 0:
                lv["d_gfeat"] = rnd(B, m, ns, cin)
                lv["d_feat"] = torch.empty((B, n, cin), dtype=f32, device=dev)
                lv["gg_ws"] = _lib.workspace(self.L.pc_group_point_grad_workspace_bytes(B, n, cin, m, ns), dev)
        for fp in self.fps:
            n, m, c = fp["n"], fp["m"], fp["c"]
            fp["d_out"] = rnd(B, n, c)
            fp["d_points2"] = torch.empty((B, m, c), dtype=f32, device=dev)
            fp["ig_ws"] = _lib.workspace(self.L.pc_three_interpolate_grad_workspace_bytes(B, n, c, m), dev)
        self.training = True
        self.launches_per_train_step = self.launches_per_step + (len(self.levels) if self.attention else 0) + \
            2 * (len(self.levels) - 1) + 2 * len(self.fps)   # each gradient op = CSR build + segmented reduce

    def _backward_level(self, lv, li, fp, side, run):
        L, B, p = self.L, self.B, _lib.ptr
        st = ctypes.c_void_p(side.cuda_stream)
        tag = "fp%d" % (4 - fp["level"])
        run("three_interpolate_grad_" + tag, side, lambda: L.pc_three_interpolate_grad(
            B, fp["n"], fp["c"], fp["m"], p(fp["d_out"]), p(fp["idx"]), p(fp["w"]), p(fp["d_points2"]), p(fp["ig_ws"]), st))
        if self.attention:
            run("attention_bwd_sa%d" % (li + 1), side, lambda: L.pc_attention_bwd(
                B * lv["m"], lv["ns"], lv["cout"] // KEY_DIM, KEY_DIM, p(lv["Q"]), p(lv["K"]), p(lv["V"]), p(lv["d_att"]),
                p(lv["dQ"]), p(lv["dK"]), p(lv["dV"]), st))
        if li > 0:
            run("group_point_grad_sa%d" % (li + 1), side, lambda: L.pc_group_point_grad(
                B, lv["n"], lv["cin"], lv["m"], lv["ns"], p(lv["d_gfeat"]), p(lv["idx"]), p(lv["d_feat"]), p(lv["gg_ws"]), st))

    def backward_op_names(self):
        names = []
        for li in range(len(self.levels)):
            names.append("three_interpolate_grad_fp%d" % (4 - li))
            if self.attention:
                names.append("attention_bwd_sa%d" % (li + 1))
            if li > 0:
                names.append("group_point_grad_sa%d" % (li + 1))
        return names

    def backward_work(self):
        """Algorithmic bytes of the gradient ops (same formulas as the forward ops with read / write swapped)."""
        B, w = self.B, {}
        for li, lv in enumerate(self.levels):
            n, m, ns, cin, cout = lv["n"], lv["m"], lv["ns"], lv["cin"], lv["cout"]
            w["group_point_grad_sa%d" % (li + 1)] = dict(kind="bytes", amount=B * (4 * m * ns + 4 * n * cin + 4 * m * ns * cin))
            # reads Q, K, V, dout; writes dQ, dK, dV
            w["attention_bwd_sa%d" % (li + 1)] = dict(kind="bytes", amount=B * m * (4 * 4 * ns * cout + 3 * 4 * cout))
        for fp in self.fps:
            n, m, c = fp["n"], fp["m"], fp["c"]
            w["three_interpolate_grad_fp%d" % (4 - fp["level"])] = dict(kind="bytes", amount=B * (24 * n + 4 * m * c + 4 * n * c))
        return w

    # ---- CUDA graph ------------------------------------------------------------------------------------------
    def capture(self, overlap=True, train=False):
        """Capture forward() (train=True: forward + backward) into a CUDA graph (after one eager warm-up); replay
        with .replay()."""
        self.forward(overlap, train=train)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        saved, self.main = self.main, None  # inside the capture the capture stream plays the role of main
        try:
            with torch.cuda.graph(g, stream=saved):
                self.forward(overlap, train=train)
        finally:
            self.main = saved
        self._graph = g
        return g

    def replay(self):
        """Launch the captured forward on this pipeline's stream."""
        with torch.cuda.stream(self.stream()):
            self._graph.replay()


def _c(rc):
    if rc != 0:
        _lib.check(rc, "pcops pipeline call")
