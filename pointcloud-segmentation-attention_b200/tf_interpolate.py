"""Interpolation ops with the reference's names and argument order, on torch CUDA tensors.

Mirrors pointnet2_tensorflow/tf_ops/interpolation_3d/tf_interpolate.py: ``three_nn(xyz1, xyz2)`` (:8-18, NoGradient),
``three_interpolate(points, idx, weight)`` (:19-28) with its registered gradient (:29-34 ->
``[ThreeInterpolateGrad, None, None]``).  The reference registers these ops for the CPU only
(tf_interpolate.cpp:187,222,262); here they are CUDA kernels.  Shape errors carry the reference OpKernel's messages
(tf_interpolate.cpp:163-168,197-206,231-243).  ``three_weights`` is pointnet_fp_module's inverse-distance weighting
(pointnet_util.py:219-222) as one kernel.
"""
import torch

from . import _lib


# Module switch between the two implementations (identical outputs, tests run both).  ONE default everywhere -- these
# wrappers, pipeline.ScanNetGeometry and the TF shim all call the cell-grid entry points, which issue
# far fewer pair tests; set it to False to get the all-pairs kernels behind the reference launchers' exact signatures
# (slightly lower latency for one small call on an idle GPU).
USE_GRID = True


@_lib.on_tensor_device
def three_nn(xyz1, xyz2):
    """xyz1 (b,n,3) unknown, xyz2 (b,m,3) known -> dist (b,n,3) f32 squared, idx (b,n,3) i32."""
    if xyz1.dim() != 3 or xyz1.shape[2] != 3:
        raise ValueError("ThreeNN expects (b,n,3) xyz1 shape.")
    if xyz2.dim() != 3 or xyz2.shape[2] != 3:
        raise ValueError("ThreeNN expects (b,m,3) xyz2 shape.")
    xyz1 = _lib.cuda_f32(xyz1.detach(), "xyz1")
    xyz2 = _lib.cuda_f32(xyz2.detach(), "xyz2")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
    idx = torch.empty((b, n, 3), dtype=torch.int32, device=xyz1.device)
    L = _lib.lib()
    if USE_GRID:  # same outputs bit for bit; cell rings around the query instead of all m known points (csrc/grid.cu)
        ws = _lib.workspace(L.pc_three_nn_grid_workspace_bytes(b, n, m), xyz1.device)
        rc = L.pc_three_nn_grid(b, n, m, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(dist), _lib.ptr(idx), _lib.ptr(ws),
                                _lib.stream())
    else:         # all-pairs kernel (csrc/interpolate.cu)
        rc = L.pc_three_nn(b, n, m, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(dist), _lib.ptr(idx), _lib.stream())
    _lib.check(rc, "pc_three_nn")
    return dist, idx


@_lib.on_tensor_device
def three_weights(dist):
    """dist (...,3) -> weight (...,3): max(dist,1e-10) -> normalised inverse distances (pointnet_util.py:219-222)."""
    if dist.shape[-1] != 3:
        raise ValueError("three_weights expects (...,3) dist")
    dist = _lib.cuda_f32(dist.detach(), "dist")
    w = torch.empty_like(dist)
    rc = _lib.lib().pc_three_weights(dist.numel() // 3, _lib.ptr(dist), _lib.ptr(w), _lib.stream())
    _lib.check(rc, "pc_three_weights")
    return w


class _ThreeInterpolate(torch.autograd.Function):
    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, points, idx, weight):
        b, m, c = points.shape
        n = idx.shape[1]
        out = torch.empty((b, n, c), dtype=torch.float32, device=points.device)
        rc = _lib.lib().pc_three_interpolate(b, m, c, n, _lib.ptr(points), _lib.ptr(idx), _lib.ptr(weight),
                                             _lib.ptr(out), _lib.stream())
        _lib.check(rc, "pc_three_interpolate")
        ctx.save_for_backward(idx, weight)
        ctx.mc = (m, c)
        return out

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        m, c = ctx.mc
        return _three_interpolate_grad(m, c, idx, weight, grad_out), None, None


@_lib.on_tensor_device
def _three_interpolate_grad(m, c, idx, weight, grad_out):
    grad_out = _lib.cuda_f32(grad_out, "grad_out")
    b, n, _ = idx.shape
    gp = torch.empty((b, m, c), dtype=torch.float32, device=grad_out.device)
    L = _lib.lib()
    ws = _lib.workspace(L.pc_three_interpolate_grad_workspace_bytes(b, n, c, m), grad_out.device)
    rc = L.pc_three_interpolate_grad(b, n, c, m, _lib.ptr(grad_out), _lib.ptr(idx), _lib.ptr(weight), _lib.ptr(gp),
                                     _lib.ptr(ws), _lib.stream())
    _lib.check(rc, "pc_three_interpolate_grad")
    return gp


@_lib.on_tensor_device
def three_interpolate(points, idx, weight):
    """points (b,m,c), idx (b,n,3) i32, weight (b,n,3) -> (b,n,c)."""
    if points.dim() != 3:
        raise ValueError("ThreeInterpolate expects (b,m,c) points shape")
    b = points.shape[0]
    if idx.dim() != 3 or idx.shape[0] != b or idx.shape[2] != 3:
        raise ValueError("ThreeInterpolate expects (b,n,3) idx shape")
    if weight.dim() != 3 or tuple(weight.shape) != (b, idx.shape[1], 3):
        raise ValueError("ThreeInterpolate expects (b,n,3) weight shape")
    return _ThreeInterpolate.apply(_lib.cuda_f32(points, "points"), _lib.cuda_i32(idx, "idx"),
                                   _lib.cuda_f32(weight.detach(), "weight"))


@_lib.on_tensor_device
def three_interpolate_grad(points, idx, weight, grad_out):
    """The ThreeInterpolateGrad op itself (tf_interpolate.cpp:37-46,225-262): `points` is used for its shape only."""
    if points.dim() != 3:
        raise ValueError("ThreeInterpolateGrad expects (b,m,c) points shape")
    b, m, c = points.shape
    if idx.dim() != 3 or idx.shape[0] != b:
        raise ValueError("ThreeInterpolateGrad expects (b,n,3) idx shape")
    n = idx.shape[1]
    if weight.dim() != 3 or tuple(weight.shape) != (b, n, 3):
        raise ValueError("ThreeInterpolateGrad expects (b,n,3) weight shape")
    if grad_out.dim() != 3 or tuple(grad_out.shape) != (b, n, c):
        raise ValueError("ThreeInterpolateGrad expects (b,n,c) grad_out shape")
    return _three_interpolate_grad(m, c, _lib.cuda_i32(idx, "idx"), _lib.cuda_f32(weight, "weight"), grad_out)
