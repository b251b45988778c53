"""Grouping ops with the reference's names and argument order, on torch CUDA tensors.

Mirrors pointnet2_tensorflow/tf_ops/grouping/tf_grouping.py: ``query_ball_point(radius, nsample, xyz1, xyz2)`` (:8-21,
NoGradient), ``select_top_k(k, dist)`` (:22-32, NoGradient), ``group_point(points, idx)`` (:33-41) with its registered
gradient (:42-46 -> ``[GroupPointGrad, None]``) and ``knn_point(k, xyz1, xyz2)`` (:48-73), which here is one fused
kernel instead of a TF graph over a (b,m,n,c) tile.  Shape / attribute errors carry the reference OpKernel's
messages (tf_grouping.cpp:70-74,79-85,112-118,149-157,180-191).
"""
import torch

from . import _lib


# Module switch between the two implementations (identical outputs, tests run both).  ONE default everywhere -- these
# wrappers, pipeline.ScanNetGeometry and the TF shim all call the cell-grid entry points, which issue
# far fewer pair tests; set it to False to get the all-pairs kernels behind the reference launchers' exact signatures
# (slightly lower latency for one small call on an idle GPU).
USE_GRID = True


@_lib.on_tensor_device
def query_ball_point(radius, nsample, xyz1, xyz2):
    """xyz1 (b,n,3) dataset, xyz2 (b,m,3) queries -> idx (b,m,nsample) i32, pts_cnt (b,m) i32."""
    if not float(radius) > 0:
        raise ValueError("QueryBallPoint expects positive radius")
    if int(nsample) <= 0:
        raise ValueError("QueryBallPoint expects positive nsample")
    if xyz1.dim() != 3 or xyz1.shape[2] != 3:
        raise ValueError("QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.")
    if xyz2.dim() != 3 or xyz2.shape[2] != 3:
        raise ValueError("QueryBallPoint expects (batch_size, npoint, 3) xyz2 shape.")
    xyz1 = _lib.cuda_f32(xyz1.detach(), "xyz1")
    xyz2 = _lib.cuda_f32(xyz2.detach(), "xyz2")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = torch.empty((b, m, int(nsample)), dtype=torch.int32, device=xyz1.device)
    cnt = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
    L = _lib.lib()
    if USE_GRID:  # same outputs bit for bit; a query only meets its 3x3x3 cell neighbourhood (csrc/grid.cu)
        ws = _lib.workspace(L.pc_query_ball_grid_workspace_bytes(b, n, m), xyz1.device)
        rc = L.pc_query_ball_grid(b, n, m, float(radius), int(nsample), _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(idx),
                                  _lib.ptr(cnt), _lib.ptr(ws), _lib.stream())
    else:         # all-pairs kernel, the reference launcher's exact signature (csrc/ball_query.cu)
        rc = L.pc_query_ball(b, n, m, float(radius), int(nsample), _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(idx),
                             _lib.ptr(cnt), _lib.stream())
    _lib.check(rc, "pc_query_ball")
    return idx, cnt


@_lib.on_tensor_device
def select_top_k(k, dist):
    """dist (b,m,n) -> (outi (b,m,n) i32, out (b,m,n) f32); first k columns are the k smallest."""
    if int(k) <= 0:
        raise ValueError("SelectionSort expects positive k")
    if dist.dim() != 3:
        raise ValueError("SelectionSort expects (b,m,n) dist shape.")
    dist = _lib.cuda_f32(dist.detach(), "dist")
    b, m, n = dist.shape
    outi = torch.empty((b, m, n), dtype=torch.int32, device=dist.device)
    out = torch.empty((b, m, n), dtype=torch.float32, device=dist.device)
    rc = _lib.lib().pc_selection_sort(b, n, m, int(k), _lib.ptr(dist), _lib.ptr(outi), _lib.ptr(out), _lib.stream())
    _lib.check(rc, "pc_selection_sort")
    return outi, out


class _GroupPoint(torch.autograd.Function):
    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, points, idx):
        b, n, c = points.shape
        _, m, ns = idx.shape
        out = torch.empty((b, m, ns, c), dtype=torch.float32, device=points.device)
        rc = _lib.lib().pc_group_point(b, n, c, m, ns, _lib.ptr(points), _lib.ptr(idx), _lib.ptr(out), _lib.stream())
        _lib.check(rc, "pc_group_point")
        ctx.save_for_backward(idx)
        ctx.nc = (n, c)
        return out

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        n, c = ctx.nc
        return _group_point_grad(n, c, idx, grad_out), None


@_lib.on_tensor_device
def _group_point_grad(n, c, idx, grad_out):
    grad_out = _lib.cuda_f32(grad_out, "grad_out")
    b, m, ns = idx.shape
    gp = torch.empty((b, n, c), dtype=torch.float32, device=grad_out.device)
    L = _lib.lib()
    ws = _lib.workspace(L.pc_group_point_grad_workspace_bytes(b, n, c, m, ns), grad_out.device)
    rc = L.pc_group_point_grad(b, n, c, m, ns, _lib.ptr(grad_out), _lib.ptr(idx), _lib.ptr(gp), _lib.ptr(ws),
                               _lib.stream())
    _lib.check(rc, "pc_group_point_grad")
    return gp


@_lib.on_tensor_device
def group_point(points, idx):
    """points (b,n,c) f32, idx (b,m,nsample) i32 -> (b,m,nsample,c)."""
    if points.dim() != 3:
        raise ValueError("GroupPoint expects (batch_size, num_points, channel) points shape")
    if idx.dim() != 3 or idx.shape[0] != points.shape[0]:
        raise ValueError("GroupPoint expects (batch_size, npoints, nsample) idx shape")
    return _GroupPoint.apply(_lib.cuda_f32(points, "points"), _lib.cuda_i32(idx, "idx"))


@_lib.on_tensor_device
def group_point_grad(points, idx, grad_out):
    """The GroupPointGrad op itself (tf_grouping.cpp:55-63,174-208): `points` is used for its shape only."""
    if points.dim() != 3:
        raise ValueError("GroupPointGrad expects (batch_size, num_points, channel) points shape")
    if idx.dim() != 3 or idx.shape[0] != points.shape[0]:
        raise ValueError("GroupPointGrad expects (batch_size, npoints, nsample) idx shape")
    b, n, c = points.shape
    if grad_out.dim() != 4 or tuple(grad_out.shape) != (b, idx.shape[1], idx.shape[2], c):
        raise ValueError("GroupPointGrad expects (batch_size, npoints, nsample, channel) grad_out shape")
    return _group_point_grad(n, c, _lib.cuda_i32(idx, "idx"), grad_out)


@_lib.on_tensor_device
def knn_point(k, xyz1, xyz2):
    """xyz1 (b,n,c) dataset, xyz2 (b,m,c) queries -> (val (b,m,k) f32 squared L2, idx (b,m,k) i32)."""
    if int(k) <= 0:
        raise ValueError("SelectionSort expects positive k")
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[0] != xyz2.shape[0] or xyz1.shape[2] != xyz2.shape[2]:
        raise ValueError("knn_point expects (b,n,c) xyz1 and (b,m,c) xyz2")
    xyz1 = _lib.cuda_f32(xyz1.detach(), "xyz1")
    xyz2 = _lib.cuda_f32(xyz2.detach(), "xyz2")
    b, n, c = xyz1.shape
    m = xyz2.shape[1]
    val = torch.empty((b, m, int(k)), dtype=torch.float32, device=xyz1.device)
    idx = torch.empty((b, m, int(k)), dtype=torch.int32, device=xyz1.device)
    rc = _lib.lib().pc_knn(b, n, m, int(k), c, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(val), _lib.ptr(idx),
                           _lib.stream())
    _lib.check(rc, "pc_knn")
    return val, idx
