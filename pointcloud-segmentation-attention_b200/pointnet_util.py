"""The geometry front-ends of the PointNet++ layers, mirroring pointnet2_tensorflow/utils/pointnet_util.py.

Only the parts of the layer library that sit ON the accelerated path are here: ``sample_and_group`` (:16-58),
``sample_and_group_all`` (:61-87) and the interpolation front half of ``pointnet_fp_module`` (:218-226).  The shared
MLPs, batch-norm and pooling that follow them in the reference are stock dense layers and are out of scope
(SURVEY.md 2.1, 8f).

Same inputs, outputs and values as the reference functions.  Where the reference strings stock TF ops around the
custom ops (tile / subtract / concat after GroupPoint; max / reciprocal / normalise / concat around ThreeInterpolate)
the default path here is ONE fused kernel (csrc/fused.cu: pc_sa_group, pc_fp_interpolate) whose outputs are
bit-identical to the op-by-op composition; ``FUSED = False`` switches back to that composition.
"""
import torch

from . import _lib
from .tf_grouping import group_point, group_point_grad, knn_point, query_ball_point
from .tf_interpolate import three_interpolate, three_interpolate_grad, three_nn, three_weights
from .tf_sampling import farthest_point_sample, farthest_point_sample_and_gather, gather_point

FUSED = True


class _SAGroup(torch.autograd.Function):
    """new_points = concat(xyz[idx] - new_xyz, points[idx]); gradient flows to ``points`` only, through GroupPointGrad
    (tf_grouping.py:42-46) -- xyz is not differentiated on this path."""

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, xyz, points, idx, new_xyz):
        b, n, _ = xyz.shape
        _, m, ns = idx.shape
        c = 0 if points is None else points.shape[2]
        new_points = torch.empty((b, m, ns, 3 + c), dtype=torch.float32, device=xyz.device)
        grouped_xyz = torch.empty((b, m, ns, 3), dtype=torch.float32, device=xyz.device)
        rc = _lib.lib().pc_sa_group(b, n, c, m, ns, _lib.ptr(xyz), _lib.ptr(points), _lib.ptr(idx), _lib.ptr(new_xyz),
                                    _lib.ptr(new_points), _lib.ptr(grouped_xyz), _lib.stream())
        _lib.check(rc, "pc_sa_group")
        ctx.save_for_backward(idx)
        ctx.shape = (n, c)
        ctx.mark_non_differentiable(grouped_xyz)
        return new_points, grouped_xyz

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, g_new_points, _g_grouped_xyz):
        (idx,) = ctx.saved_tensors
        n, c = ctx.shape
        gp = None
        if c > 0 and ctx.needs_input_grad[1]:
            b = idx.shape[0]
            ref = torch.empty((b, n, c), dtype=torch.float32, device=idx.device)  # shape carrier, as in the reference op
            gp = group_point_grad(ref, idx, g_new_points[..., 3:].contiguous())
        return None, gp, None, None


class _FPInterpolate(torch.autograd.Function):
    """out = concat(three_interpolate(points2, idx, w(dist)), points1); gradients to points2 (ThreeInterpolateGrad,
    tf_interpolate.py:29-34) and points1 (slice)."""

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, dist, idx, points2, points1):
        b, n, _ = idx.shape
        m, c2 = points2.shape[1], points2.shape[2]
        c1 = 0 if points1 is None else points1.shape[2]
        out = torch.empty((b, n, c2 + c1), dtype=torch.float32, device=points2.device)
        weight = torch.empty((b, n, 3), dtype=torch.float32, device=points2.device)
        rc = _lib.lib().pc_fp_interpolate(b, n, m, c2, c1, _lib.ptr(dist), _lib.ptr(idx), _lib.ptr(points2),
                                          _lib.ptr(points1), _lib.ptr(out), _lib.ptr(weight), _lib.stream())
        _lib.check(rc, "pc_fp_interpolate")
        ctx.save_for_backward(idx, weight)
        ctx.dims = (m, c2, c1)
        return out

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, g_out):
        idx, weight = ctx.saved_tensors
        m, c2, c1 = ctx.dims
        g2 = g1 = None
        if ctx.needs_input_grad[2]:
            ref = torch.empty((idx.shape[0], m, c2), dtype=torch.float32, device=idx.device)
            g2 = three_interpolate_grad(ref, idx, weight, g_out[..., :c2].contiguous())
        if c1 > 0 and ctx.needs_input_grad[3]:
            g1 = g_out[..., c2:].contiguous()
        return None, None, g2, g1


@_lib.on_tensor_device
def sample_and_group(npoint, radius, nsample, xyz, points, knn=False, use_xyz=True):
    """Same inputs / outputs as pointnet_util.py:16-58:
    returns new_xyz (B,npoint,3), new_points (B,npoint,nsample,3+C), idx (B,npoint,nsample), grouped_xyz."""
    if xyz.requires_grad:  # keep GatherPoint's registered gradient in the graph (tf_sampling.py:44-48)
        new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
    else:                  # one launch for the pair
        _, new_xyz = farthest_point_sample_and_gather(npoint, xyz)
    if knn:
        _, idx = knn_point(nsample, xyz, new_xyz)
    else:
        idx, _pts_cnt = query_ball_point(radius, nsample, xyz, new_xyz)
    if points is not None and points.numel() == 0 and points.dim() == 1:
        points = None  # the reference's Keras-layer convention (:41-42)
    if FUSED and use_xyz and not xyz.requires_grad:
        new_points, grouped_xyz = _SAGroup.apply(_lib.cuda_f32(xyz, "xyz"),
                                                 None if points is None else _lib.cuda_f32(points, "points"),
                                                 idx, new_xyz)
        return new_xyz, new_points, idx, grouped_xyz
    grouped_xyz = group_point(xyz, idx)
    grouped_xyz = grouped_xyz - new_xyz.unsqueeze(2)  # translation normalisation (:40)
    if points is not None:
        grouped_points = group_point(points, idx)
        new_points = torch.cat([grouped_xyz, grouped_points], dim=-1) if use_xyz else grouped_points
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


@_lib.on_tensor_device
def sample_and_group_all(xyz, points, use_xyz=True):
    """pointnet_util.py:61-87: one group holding every point, centroid (0,0,0)."""
    b, n, _ = xyz.shape
    new_xyz = torch.zeros((b, 1, 3), dtype=torch.float32, device=xyz.device)
    idx = torch.arange(n, dtype=torch.int32, device=xyz.device).reshape(1, 1, n).repeat(b, 1, 1)
    grouped_xyz = xyz.reshape(b, 1, n, 3)
    if points is not None:
        new_points = torch.cat([xyz, points], dim=2) if use_xyz else points
        new_points = new_points.unsqueeze(1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


@_lib.on_tensor_device
def fp_interpolate(xyz1, xyz2, points1, points2):
    """Front half of pointnet_fp_module (pointnet_util.py:218-226): three_nn -> inverse-distance weights ->
    three_interpolate -> concat with the skip features.  Returns (B, n1, C2 [+ C1])."""
    dist, idx = three_nn(xyz1, xyz2)
    if FUSED:
        return _FPInterpolate.apply(dist, idx, _lib.cuda_f32(points2, "points2"),
                                    None if points1 is None else _lib.cuda_f32(points1, "points1"))
    weight = three_weights(dist)
    interpolated = three_interpolate(points2, idx, weight)
    if points1 is not None:
        return torch.cat([interpolated, points1], dim=2)
    return interpolated
