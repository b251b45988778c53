"""ctypes binding of libpcops.so (include/pcops.h) -- the only doorway from Python to the CUDA kernels.

There is NO fallback: if the shared library is missing, or an op is handed a non-CUDA tensor, the call raises.
torch is used for device memory and the current stream only.
"""
import ctypes
import functools
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcops.so")
_lib = None

PC_OK = 0
PC_ERR_INVALID_ARGUMENT = -1
PC_ERR_UNSUPPORTED = -2
PC_ERR_WORKSPACE = -3

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/pcops.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "pc_version": (_i, []),
    "pc_error_string": (ctypes.c_char_p, [_i]),
    "pc_num_sms": (_i, []),
    "pc_set_concurrency_hint": (_i, [_i]),
    "pc_host_legacy_shuffle": (_i, [_vp, _vp, _i, _vp]),
    "pc_host_legacy_randint": (_i, [_vp, _vp, _i, _i, _vp]),
    "pc_get_concurrency_hint": (_i, []),
    "pc_fps_workspace_bytes": (_sz, [_i, _i, _i]),
    "pc_fps": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "pc_fps_gather": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_cumsum": (_i, [_i, _i, _vp, _vp, _vp]),
    "pc_prob_sample": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_gather_point": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "pc_gather_point_grad_workspace_bytes": (_sz, [_i, _i, _i]),
    "pc_gather_point_grad": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_query_ball": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_query_ball_grid_workspace_bytes": (_sz, [_i, _i, _i]),
    "pc_query_ball_grid": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pc_group_point": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pc_group_point_grad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "pc_group_point_grad": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_selection_sort": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pc_knn": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_three_nn": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_three_nn_grid_workspace_bytes": (_sz, [_i, _i, _i]),
    "pc_three_nn_grid": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pc_three_weights": (_i, [_sz, _vp, _vp, _vp]),
    "pc_three_interpolate": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_three_interpolate_grad_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "pc_three_interpolate_grad": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pc_sa_group": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pc_fp_interpolate": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pc_attention_layer_workspace_bytes": (_sz, [_i, _i, _i]),
    "pc_attention_layer_fwd": (_i, [_i, _i, _i] + [_vp] * 11),
    "pc_attention_layer_prepare": (_i, [_i, _i] + [_vp] * 8),
    "pc_attention_layer_fwd_prepared": (_i, [_i, _i, _i, _vp, _sz] + [_vp] * 10),
    "pc_attention_fwd": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_inner_attention_fwd": (_i, [_sz, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_attention_bwd": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pc_scene_cells_workspace_bytes": (_sz, [_i, _i]),
    "pc_scene_bbox": (_i, [_i, _vp, _vp, _vp]),
    "pc_scene_cells": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pc_scene_chunk_masksum": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_scene_chunk_assemble": (_i, [_i, _i] + [_vp] * 11),
    "pc_gather_rows_bytes": (_i, [_sz, _i, _vp, _vp, _vp, _vp]),
    "pc_scene_sample_weights": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pc_map_back_winner": (_i, [_sz, _i, _vp, _vp, _vp, _vp]),
    "pc_dense_image_bytes": (_sz, [_i, _i]),
    "pc_dense_prepare": (_i, [_i, _i, _vp, _i, _vp, _vp]),
    "pc_dense_fwd": (_i, [_sz, _i, _i, _vp, _sz, _vp, _vp, _i, _vp, _sz, _vp]),
    "pc_dense_pool_fwd": (_i, [_sz, _i, _i, _i, _vp, _sz, _vp, _vp, _i, _vp, _sz, _vp, _sz, _vp]),
    "pc_dense_bwd_weight_workspace_bytes": (_sz, [_sz, _i, _i]),
    "pc_dense_bwd_weight": (_i, [_sz, _i, _i, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp]),
    "pc_unpack_features": (_i, [_sz, _vp, _vp, _vp, _vp]),
    "pc_narrow_indices_u16": (_i, [_sz, _vp, _vp, _vp]),
}


class PcopsError(RuntimeError):
    pass


def lib():
    """Load libpcops.so once.  Raises if it has not been built (python __graft_entry__.py / build.sh)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PcopsError("libpcops.so not found at %s -- build it with `sh %s` (there is no CPU fallback)"
                             % (LIB_PATH, os.path.join(_HERE, "build.sh")))
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what, invalid_message=None):
    """Turn a pcops return code into the exception the reference op would raise."""
    if rc == PC_OK:
        return
    if rc == PC_ERR_INVALID_ARGUMENT:
        raise ValueError(invalid_message or ("%s: invalid argument" % what))
    name = lib().pc_error_string(rc).decode()
    if rc == PC_ERR_UNSUPPORTED:
        raise NotImplementedError("%s: %s" % (what, name))
    raise PcopsError("%s failed: %s (%d)" % (what, name, rc))


def stream():
    """The current stream of the CURRENT device.  Every wrapper that launches runs under on_tensor_device, so the
    current device is the one its tensor arguments live on."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda_devices(args):
    for a in args:
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                yield a.device
        elif isinstance(a, (list, tuple)):
            for d in _cuda_devices(a):
                yield d


def on_tensor_device(fn):
    """Decorator of every op wrapper: all CUDA tensor arguments must share ONE device, and the call (stream lookup,
    workspace allocation, kernel launch, per-device kernel attributes inside libpcops.so) runs with that device
    current -- not whatever device happens to be current in the calling thread."""
    @functools.wraps(fn)
    def wrapped(*args, **kw):
        dev = None
        for d in _cuda_devices(list(args) + list(kw.values())):
            if dev is None:
                dev = d
            elif d != dev:
                raise PcopsError("%s: all tensor arguments must live on one CUDA device, got %s and %s"
                                 % (fn.__name__, dev, d))
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)
    return wrapped


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def cuda_f32(t, what):
    return _prep(t, torch.float32, what)


def cuda_i32(t, what):
    return _prep(t, torch.int32, what)


def _prep(t, dtype, what):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % what)
    if not t.is_cuda:
        raise PcopsError("%s must be a CUDA tensor: these ops have no CPU implementation" % what)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (what, dtype, t.dtype))
    return t.contiguous()


def workspace(nbytes, device):
    if nbytes == 0:
        return None
    return torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=device)


def set_concurrency_hint(n):
    """Tell the library how many independent launches the caller keeps in flight (pc_set_concurrency_hint): 1 (default)
    sizes the streaming kernels for a lone launch, larger values as few long-lived CTAs that co-reside with other
    kernels.  Results never depend on it.  Returns the previous value."""
    L = lib()
    old = L.pc_get_concurrency_hint()
    check(L.pc_set_concurrency_hint(int(n)), "pc_set_concurrency_hint")
    return old
