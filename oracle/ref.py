"""The reference's own code, compiled unmodified into oracle/_ref/*.so (see oracle/Makefile).

TEST INFRASTRUCTURE ONLY.  ``cpu_*`` functions call the reference's CPU functions on numpy
arrays; ``Gpu`` calls the reference's CUDA launchers on torch CUDA tensors (legacy default
stream; the wrapper synchronises).  ``available_*`` say whether the prebuilt library exists --
/root/reference is never read at run time.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)
_cache = {}


def _path(name):
    return os.path.join(_REF, name)


def available_cpu():
    return all(os.path.exists(_path(n)) for n in
               ("libref_grouping_cpu.so", "libref_selsort_cpu.so", "libref_interpolate_cpu.so"))


def available_gpu(nofma=True):
    return os.path.exists(_path("libref_gpu_nofma.so" if nofma else "libref_gpu.so"))


def _load(name):
    if name not in _cache:
        _cache[name] = ctypes.CDLL(_path(name))
    return _cache[name]


def _sym(libname, needle):
    """Resolve a C++-mangled export by its plain function name."""
    key = (libname, needle)
    if key not in _cache:
        out = subprocess.run(["nm", "-D", "--defined-only", _path(libname)], check=True, capture_output=True,
                             text=True).stdout
        hits = [ln.split()[-1] for ln in out.splitlines() if needle in ln.split()[-1] and " T " in ln]
        if len(hits) != 1:
            raise RuntimeError("cannot resolve %s in %s: %s" % (needle, libname, hits))
        fn = getattr(_load(libname), hits[0])
        fn.restype = None
        _cache[key] = fn
    return _cache[key]


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---- grouping/test/query_ball_point.cpp:19-84 ------------------------------------------------
def cpu_query_ball_point(radius, nsample, xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.zeros((b, m, nsample), np.int32)  # the program memsets idx first (:94)
    _sym("libref_grouping_cpu.so", "query_ball_point_cpu")(b, n, m, ctypes.c_float(radius), nsample, _fp(xyz1),
                                                          _fp(xyz2), _ip(idx))
    return idx


def cpu_group_point(points, idx):
    points, idx = _f32(points), _i32(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.zeros((b, m, ns, c), np.float32)
    _sym("libref_grouping_cpu.so", "group_point_cpu")(b, n, c, m, ns, _fp(points), _ip(idx), _fp(out))
    return out


def cpu_group_point_grad(points, idx, grad_out):
    points, idx, grad_out = _f32(points), _i32(idx), _f32(grad_out)
    b, n, c = points.shape
    _, m, ns = idx.shape
    gp = np.zeros((b, n, c), np.float32)
    _sym("libref_grouping_cpu.so", "group_point_grad_cpu")(b, n, c, m, ns, _fp(grad_out), _ip(idx), _fp(gp))
    return gp


# ---- grouping/test/selection_sort.cpp:20-63 (prints while sorting: tiny inputs only) --------
def cpu_selection_sort(k, dist):
    dist = _f32(dist)
    b, m, n = dist.shape
    idx = np.zeros((b, m, n), np.int32)
    val = np.zeros((b, m, n), np.float32)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    try:
        os.dup2(devnull, 1)
        _sym("libref_selsort_cpu.so", "selection_sort_cpu")(b, n, m, k, _fp(dist), _ip(idx), _fp(val))
        ctypes.CDLL(None).fflush(None)
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return idx, val


# ---- interpolation_3d/tf_interpolate.cpp:60-153 ------------------------------------------------
def cpu_three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.zeros((b, n, 3), np.float32)
    idx = np.zeros((b, n, 3), np.int32)
    _sym("libref_interpolate_cpu.so", "threenn_cpu")(b, n, m, _fp(xyz1), _fp(xyz2), _fp(dist), _ip(idx))
    return dist, idx


def cpu_three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.zeros((b, n, c), np.float32)
    _sym("libref_interpolate_cpu.so", "threeinterpolate_cpu")(b, m, c, n, _fp(points), _ip(idx), _fp(weight), _fp(out))
    return out


def cpu_three_interpolate_grad(points, idx, weight, grad_out):
    points, idx, weight, grad_out = _f32(points), _i32(idx), _f32(weight), _f32(grad_out)
    b, m, c = points.shape
    n = idx.shape[1]
    gp = np.zeros((b, m, c), np.float32)  # the op memsets first (tf_interpolate.cpp:258)
    _sym("libref_interpolate_cpu.so", "threeinterpolate_grad_cpu")(b, n, c, m, _fp(grad_out), _ip(idx), _fp(weight),
                                                                  _fp(gp))
    return gp


# ---- the reference CUDA launchers (tf_sampling_g.cu:194-211, tf_grouping_g.cu:125-141) -------
class Gpu:
    """Reference GPU kernels on torch CUDA tensors.  nofma=True -> built with --fmad=false."""

    def __init__(self, nofma=True):
        import torch
        self.torch = torch
        self.lib = ctypes.CDLL(_path("libref_gpu_nofma.so" if nofma else "libref_gpu.so"))

    def _p(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def _sync(self):
        self.torch.cuda.synchronize()
        rc = self.lib.ref_sync()
        if rc:
            raise RuntimeError("reference kernel failed: cuda error %d" % rc)

    def prob_sample(self, inp, inpr):
        t = self.torch
        b, n = inp.shape
        m = inpr.shape[1]
        out = t.zeros((b, m), dtype=t.int32, device=inp.device)
        temp = t.empty((b, n), dtype=t.float32, device=inp.device)  # tf_sampling.cpp:85-88
        self._sync()
        self.lib.ref_prob_sample(b, n, m, self._p(inp), self._p(inpr), self._p(temp), self._p(out))
        self._sync()
        return out, temp

    def farthest_point_sample(self, npoint, inp):
        t = self.torch
        b, n, _ = inp.shape
        out = t.zeros((b, npoint), dtype=t.int32, device=inp.device)
        temp = t.empty((32, n), dtype=t.float32, device=inp.device)  # tf_sampling.cpp:115
        self._sync()
        self.lib.ref_fps(b, n, npoint, self._p(inp), self._p(temp), self._p(out))
        self._sync()
        return out

    def fps_launch(self, npoint, inp, temp, out):
        b, n, _ = inp.shape
        return self.lib.ref_fps(b, n, npoint, self._p(inp), self._p(temp), self._p(out))

    def gather_point(self, inp, idx):
        t = self.torch
        b, n, _ = inp.shape
        m = idx.shape[1]
        out = t.zeros((b, m, 3), dtype=t.float32, device=inp.device)
        self._sync()
        self.lib.ref_gather_point(b, n, m, self._p(inp), self._p(idx), self._p(out))
        self._sync()
        return out

    def query_ball_point(self, radius, nsample, xyz1, xyz2):
        t = self.torch
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        idx = t.zeros((b, m, nsample), dtype=t.int32, device=xyz1.device)
        cnt = t.zeros((b, m), dtype=t.int32, device=xyz1.device)
        self._sync()
        self.lib.ref_query_ball(b, n, m, ctypes.c_float(radius), nsample, self._p(xyz1), self._p(xyz2), self._p(idx),
                                self._p(cnt))
        self._sync()
        return idx, cnt

    def ball_launch(self, radius, nsample, xyz1, xyz2, idx, cnt):
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        return self.lib.ref_query_ball(b, n, m, ctypes.c_float(radius), nsample, self._p(xyz1), self._p(xyz2),
                                       self._p(idx), self._p(cnt))

    def group_point(self, points, idx):
        t = self.torch
        b, n, c = points.shape
        _, m, ns = idx.shape
        out = t.zeros((b, m, ns, c), dtype=t.float32, device=points.device)
        self._sync()
        self.lib.ref_group_point(b, n, c, m, ns, self._p(points), self._p(idx), self._p(out))
        self._sync()
        return out

    def group_launch(self, points, idx, out):
        b, n, c = points.shape
        _, m, ns = idx.shape
        return self.lib.ref_group_point(b, n, c, m, ns, self._p(points), self._p(idx), self._p(out))

    def group_point_grad(self, points, idx, grad_out):
        t = self.torch
        b, n, c = points.shape
        _, m, ns = idx.shape
        gp = t.zeros((b, n, c), dtype=t.float32, device=points.device)
        self._sync()
        self.lib.ref_group_point_grad(b, n, c, m, ns, self._p(grad_out), self._p(idx), self._p(gp))
        self._sync()
        return gp

    def select_top_k(self, k, dist):
        t = self.torch
        b, m, n = dist.shape
        outi = t.zeros((b, m, n), dtype=t.int32, device=dist.device)
        out = t.zeros((b, m, n), dtype=t.float32, device=dist.device)
        self._sync()
        self.lib.ref_selection_sort(b, n, m, k, self._p(dist), self._p(outi), self._p(out))
        self._sync()
        return outi, out
