"""Sampling ops with the reference's names and argument order, on torch CUDA tensors.

Mirrors pointnet2_tensorflow/tf_ops/sampling/tf_sampling.py: ``farthest_point_sample(npoint, inp)`` (:49-58,
NoGradient), ``gather_point(inp, idx)`` (:30-38) with its registered gradient (:44-48 -> ``[GatherPointGrad, None]``).
``prob_sample(inp, inpr)`` (:14-23, NoGradient) is on no model's path but part of the library's op
surface.  Shape errors carry the reference OpKernel's messages (tf_sampling.cpp:105,131,135).
"""
import torch

from . import _lib


@_lib.on_tensor_device
def farthest_point_sample(npoint, inp):
    """inp (b,n,3) f32 -> (b,npoint) i32.  FarthestPointSample, tf_sampling.cpp:28-40,95-123."""
    if int(npoint) <= 0:
        raise ValueError("FarthestPointSample expects positive npoint")
    if inp.dim() != 3 or inp.shape[2] != 3:
        raise ValueError("FarthestPointSample expects (batch_size,num_points,3) inp shape")
    inp = _lib.cuda_f32(inp.detach(), "inp")
    b, n, _ = inp.shape
    out = torch.empty((b, int(npoint)), dtype=torch.int32, device=inp.device)
    L = _lib.lib()
    ws = _lib.workspace(L.pc_fps_workspace_bytes(b, n, int(npoint)), inp.device)
    rc = L.pc_fps(b, n, int(npoint), _lib.ptr(inp), _lib.ptr(ws), _lib.ptr(out), _lib.stream())
    _lib.check(rc, "pc_fps", "FarthestPointSample expects (batch_size,num_points,3) inp shape")
    return out


@_lib.on_tensor_device
def farthest_point_sample_and_gather(npoint, inp):
    """farthest_point_sample followed by gather_point on the same cloud (pointnet_util.py:34) in ONE kernel launch:
    inp (b,n,3) -> (idx (b,npoint) i32, new_xyz (b,npoint,3) f32).  No gradient flows to inp (as for FPS)."""
    if int(npoint) <= 0:
        raise ValueError("FarthestPointSample expects positive npoint")
    if inp.dim() != 3 or inp.shape[2] != 3:
        raise ValueError("FarthestPointSample expects (batch_size,num_points,3) inp shape")
    inp = _lib.cuda_f32(inp.detach(), "inp")
    b, n, _ = inp.shape
    out = torch.empty((b, int(npoint)), dtype=torch.int32, device=inp.device)
    new_xyz = torch.empty((b, int(npoint), 3), dtype=torch.float32, device=inp.device)
    L = _lib.lib()
    ws = _lib.workspace(L.pc_fps_workspace_bytes(b, n, int(npoint)), inp.device)
    rc = L.pc_fps_gather(b, n, int(npoint), _lib.ptr(inp), _lib.ptr(ws), _lib.ptr(out), _lib.ptr(new_xyz), _lib.stream())
    _lib.check(rc, "pc_fps_gather", "FarthestPointSample expects (batch_size,num_points,3) inp shape")
    return out, new_xyz


class _GatherPoint(torch.autograd.Function):
    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, inp, idx):
        b, n, _ = inp.shape
        m = idx.shape[1]
        out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
        rc = _lib.lib().pc_gather_point(b, n, m, _lib.ptr(inp), _lib.ptr(idx), _lib.ptr(out), _lib.stream())
        _lib.check(rc, "pc_gather_point")
        ctx.save_for_backward(idx)
        ctx.n = n
        return out

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, out_g):
        (idx,) = ctx.saved_tensors
        return gather_point_grad_shape(ctx.n, idx, out_g), None


@_lib.on_tensor_device
def gather_point_grad_shape(n, idx, out_g):
    out_g = _lib.cuda_f32(out_g, "out_g")
    b, m = idx.shape
    inp_g = torch.empty((b, n, 3), dtype=torch.float32, device=out_g.device)
    L = _lib.lib()
    ws = _lib.workspace(L.pc_gather_point_grad_workspace_bytes(b, n, m), out_g.device)
    rc = L.pc_gather_point_grad(b, n, m, _lib.ptr(out_g), _lib.ptr(idx), _lib.ptr(inp_g), _lib.ptr(ws), _lib.stream())
    _lib.check(rc, "pc_gather_point_grad")
    return inp_g


@_lib.on_tensor_device
def gather_point(inp, idx):
    """inp (b,n,3) f32, idx (b,m) i32 -> (b,m,3).  GatherPoint, tf_sampling.cpp:41-54,126-148."""
    if inp.dim() != 3 or inp.shape[2] != 3:
        raise ValueError("GatherPoint expects (batch_size,num_points,3) inp shape")
    if idx.dim() != 2 or idx.shape[0] != inp.shape[0]:
        raise ValueError("GatherPoint expects (batch_size,num_result) idx shape")
    return _GatherPoint.apply(_lib.cuda_f32(inp, "inp"), _lib.cuda_i32(idx, "idx"))


@_lib.on_tensor_device
def gather_point_grad(inp, idx, out_g):
    """The GatherPointGrad op itself (tf_sampling.cpp:55-63,151-178): (inp, idx, out_g) -> inp_g (b,n,3)."""
    if inp.dim() != 3 or inp.shape[2] != 3:
        raise ValueError("GatherPointGradGpuOp expects (batch_size,num_points,3) inp")
    if idx.dim() != 2 or idx.shape[0] != inp.shape[0]:
        raise ValueError("GatherPointGradGpuOp expects (batch_size,num_result) idx shape")
    if out_g.dim() != 3 or tuple(out_g.shape) != (inp.shape[0], idx.shape[1], 3):
        raise ValueError("GatherPointGradGpuOp expects (batch_size,num_result,3) out_g shape")
    return gather_point_grad_shape(inp.shape[1], _lib.cuda_i32(idx, "idx"), out_g)


@_lib.on_tensor_device
def prob_sample(inp, inpr):
    """inp (b,ncategory) f32 weights, inpr (b,npoints) f32 uniforms -> (b,npoints) i32.  ProbSample,
    tf_sampling.py:14-23 / tf_sampling.cpp:14-27,66-92; NoGradient (tf_sampling.py:23)."""
    if inp.dim() != 2:
        raise ValueError("ProbSample expects (batch_size,num_choices) inp shape")
    if inpr.dim() != 2 or inpr.shape[0] != inp.shape[0]:
        raise ValueError("ProbSample expects (batch_size,num_points) inpr shape")
    inp = _lib.cuda_f32(inp.detach(), "inp")
    inpr = _lib.cuda_f32(inpr.detach(), "inpr")
    b, n = inp.shape
    m = inpr.shape[1]
    temp = torch.empty((b, n), dtype=torch.float32, device=inp.device)
    out = torch.empty((b, m), dtype=torch.int32, device=inp.device)
    rc = _lib.lib().pc_prob_sample(b, n, m, _lib.ptr(inp), _lib.ptr(inpr), _lib.ptr(temp), _lib.ptr(out), _lib.stream())
    _lib.check(rc, "pc_prob_sample")
    return out


@_lib.on_tensor_device
def cumsum(inp):
    """(b,n) f32 -> (b,n) cumulative sums in the reference's summation order (cumsumLauncher, tf_sampling_g.cu:193-195)."""
    if inp.dim() != 2:
        raise ValueError("cumsum expects (batch_size,n) inp shape")
    inp = _lib.cuda_f32(inp.detach(), "inp")
    out = torch.empty_like(inp)
    rc = _lib.lib().pc_cumsum(inp.shape[0], inp.shape[1], _lib.ptr(inp), _lib.ptr(out), _lib.stream())
    _lib.check(rc, "pc_cumsum")
    return out
