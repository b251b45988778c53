"""numpy wrappers over liboracle.so (oracle/pcops_oracle.c).  TEST INFRASTRUCTURE ONLY.

Function names and argument order follow the reference's Python op wrappers
(tf_ops/sampling/tf_sampling.py, tf_ops/grouping/tf_grouping.py,
tf_ops/interpolation_3d/tf_interpolate.py) so parity tests read like the reference's own.
``omp=True`` selects the OpenMP-over-batch loops used only for cpu_baseline timing.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is mounted)."""
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return int(lib().orc_num_threads())


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def farthest_point_sample(npoint, inp, omp=False):
    inp = _f32(inp)
    b, n, _ = inp.shape
    out = np.zeros((b, npoint), np.int32)
    fn = lib().orc_fps_omp if omp else lib().orc_fps
    fn(b, n, npoint, _fp(inp), _ip(out))
    return out


def cumsum(inp):
    inp = _f32(inp)
    b, n = inp.shape
    out = np.zeros((b, n), np.float32)
    lib().orc_cumsum(b, n, _fp(inp), _fp(out))
    return out


def prob_sample(inp, inpr):
    """inp (b,n) weights, inpr (b,m) uniforms in [0,1) -> (b,m) i32 category indices (tf_sampling.py:14-23)."""
    inp, inpr = _f32(inp), _f32(inpr)
    b, n = inp.shape
    m = inpr.shape[1]
    temp = np.zeros((b, n), np.float32)
    out = np.zeros((b, m), np.int32)
    lib().orc_prob_sample(b, n, m, _fp(inp), _fp(inpr), _fp(temp), _ip(out))
    return out


def gather_point(inp, idx):
    inp, idx = _f32(inp), _i32(idx)
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = np.zeros((b, m, 3), np.float32)
    lib().orc_gather_point(b, n, m, _fp(inp), _ip(idx), _fp(out))
    return out


def gather_point_grad(inp, idx, out_g):
    inp, idx, out_g = _f32(inp), _i32(idx), _f32(out_g)
    b, n, _ = inp.shape
    m = idx.shape[1]
    inp_g = np.zeros((b, n, 3), np.float32)
    lib().orc_gather_point_grad(b, n, m, _fp(out_g), _ip(idx), _fp(inp_g))
    return inp_g


def query_ball_point(radius, nsample, xyz1, xyz2, omp=False):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.zeros((b, m, nsample), np.int32)
    cnt = np.zeros((b, m), np.int32)
    fn = lib().orc_query_ball_omp if omp else lib().orc_query_ball
    fn(b, n, m, ctypes.c_float(radius), nsample, _fp(xyz1), _fp(xyz2), _ip(idx), _ip(cnt))
    return idx, cnt


def group_point(points, idx, omp=False):
    points, idx = _f32(points), _i32(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.zeros((b, m, ns, c), np.float32)
    fn = lib().orc_group_point_omp if omp else lib().orc_group_point
    fn(b, n, c, m, ns, _fp(points), _ip(idx), _fp(out))
    return out


def group_point_grad(points, idx, grad_out, omp=False):
    points, idx, grad_out = _f32(points), _i32(idx), _f32(grad_out)
    b, n, c = points.shape
    _, m, ns = idx.shape
    gp = np.zeros((b, n, c), np.float32)
    fn = lib().orc_group_point_grad_omp if omp else lib().orc_group_point_grad
    fn(b, n, c, m, ns, _fp(grad_out), _ip(idx), _fp(gp))
    return gp


def select_top_k(k, dist):
    dist = _f32(dist)
    b, m, n = dist.shape
    outi = np.zeros((b, m, n), np.int32)
    out = np.zeros((b, m, n), np.float32)
    lib().orc_selection_sort(b, n, m, k, _fp(dist), _ip(outi), _fp(out))
    return outi, out


def knn_dist(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, c = xyz1.shape
    m = xyz2.shape[1]
    dist = np.zeros((b, m, n), np.float32)
    lib().orc_knn_dist(b, n, m, c, _fp(xyz1), _fp(xyz2), _fp(dist))
    return dist


def knn_point(k, xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, c = xyz1.shape
    m = xyz2.shape[1]
    val = np.zeros((b, m, k), np.float32)
    idx = np.zeros((b, m, k), np.int32)
    lib().orc_knn(b, n, m, k, c, _fp(xyz1), _fp(xyz2), _fp(val), _ip(idx))
    return val, idx


def three_nn(xyz1, xyz2, omp=False):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.zeros((b, n, 3), np.float32)
    idx = np.zeros((b, n, 3), np.int32)
    fn = lib().orc_three_nn_omp if omp else lib().orc_three_nn
    fn(b, n, m, _fp(xyz1), _fp(xyz2), _fp(dist), _ip(idx))
    return dist, idx


def three_weights(dist):
    dist = _f32(dist)
    w = np.zeros_like(dist)
    lib().orc_three_weights(ctypes.c_size_t(dist.size // 3), _fp(dist), _fp(w))
    return w


def three_interpolate(points, idx, weight, omp=False):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.zeros((b, n, c), np.float32)
    fn = lib().orc_three_interpolate_omp if omp else lib().orc_three_interpolate
    fn(b, m, c, n, _fp(points), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(points, idx, weight, grad_out, omp=False):
    points, idx, weight, grad_out = _f32(points), _i32(idx), _f32(weight), _f32(grad_out)
    b, m, c = points.shape
    n = idx.shape[1]
    gp = np.zeros((b, m, c), np.float32)
    fn = lib().orc_three_interpolate_grad_omp if omp else lib().orc_three_interpolate_grad
    fn(b, n, c, m, _fp(grad_out), _ip(idx), _fp(weight), _fp(gp))
    return gp


def attention_fwd(Q, K, V, heads, key_dim, return_attn=False):
    """Q (G,HD), K (G,S,HD), V (G,S,HD) -> out (G,HD) [, attn (G,H,S)]."""
    Q, K, V = _f32(Q), _f32(K), _f32(V)
    G, S, HD = K.shape
    assert HD == heads * key_dim and Q.shape == (G, HD) and V.shape == K.shape
    out = np.zeros((G, HD), np.float32)
    attn = np.zeros((G, heads, S), np.float32) if return_attn else None
    lib().orc_attention_fwd(G, S, heads, key_dim, _fp(Q), _fp(K), _fp(V), _fp(out),
                            _fp(attn) if return_attn else None)
    return (out, attn) if return_attn else out


def attention_bwd(Q, K, V, dout, heads, key_dim):
    Q, K, V, dout = _f32(Q), _f32(K), _f32(V), _f32(dout)
    G, S, HD = K.shape
    dQ, dK, dV = np.zeros_like(Q), np.zeros_like(K), np.zeros_like(V)
    lib().orc_attention_bwd(G, S, heads, key_dim, _fp(Q), _fp(K), _fp(V), _fp(dout), _fp(dQ), _fp(dK), _fp(dV))
    return dQ, dK, dV


def dense(x, W, bias):
    x, W = _f32(x), _f32(W)
    cin, cout = W.shape
    rows = x.size // cin
    y = np.zeros(x.shape[:-1] + (cout,), np.float32)
    bias = _f32(bias) if bias is not None else None
    lib().orc_dense(ctypes.c_size_t(rows), cin, cout, _fp(x), _fp(W), _fp(bias) if bias is not None else None, _fp(y))
    return y


def attention_layer(x, xq, Wq, bq, Wk, bk, Wv, bv, heads, key_dim):
    """AttentionLayer.call (attention_layer.py:29-45): x (G,S,Cin), xq (G,Cin) -> (G, heads*key_dim)."""
    Q = dense(xq, Wq, bq)
    K = dense(x, Wk, bk)
    V = dense(x, Wv, bv)
    return attention_fwd(Q, K, V, heads, key_dim)


# ---- the shared-MLP block of a set-abstraction level (float64 numpy restatement; TF absent -> source-pinned) ----------
def batch_norm_inference(x, gamma, beta, moving_mean, moving_var, eps=1e-3):
    """tf.contrib.layers.batch_norm(center=True, scale=True, is_training=False) as tf_util.batch_norm_template calls it
    (utils/tf_util.py:512-530; the layer's default epsilon is 0.001), float64."""
    x = np.asarray(x, np.float64)
    return (x - np.asarray(moving_mean, np.float64)) / np.sqrt(np.asarray(moving_var, np.float64) + eps) * \
        np.asarray(gamma, np.float64) + np.asarray(beta, np.float64)


def conv2d_1x1(x, W, b, bn=None, relu=True):
    """tf_util.conv2d with kernel [1,1], stride 1, VALID (utils/tf_util.py:120-186): tf.nn.conv2d -> bias_add -> batch norm
    (inference) -> ReLU, over the last axis.  x (..., cin), W (cin, cout) = the conv kernel [1,1,cin,cout] squeezed;
    bn = (gamma, beta, moving_mean, moving_var) or None.  float64 accumulation, float64 result."""
    y = np.asarray(x, np.float64) @ np.asarray(W, np.float64)
    if b is not None:
        y = y + np.asarray(b, np.float64)
    if bn is not None:
        y = batch_norm_inference(y, *bn)
    return np.maximum(y, 0.0) if relu else y


def shared_mlp(new_points, layers):
    """The 'Point Feature Embedding' loop of pointnet_sa_module (utils/pointnet_util.py:119-131): layers = [(W, b, bn)]."""
    y = np.asarray(new_points, np.float64)
    for W, b, bn in layers:
        y = conv2d_1x1(y, W, b, bn, True)
    return y


def sa_module_tail(new_points, layers, pooling="max"):
    """MLP + 'Pooling in Local Regions' (pointnet_util.py:119-135): (B,m,ns,C) -> (B,m,mlp[-1])."""
    y = shared_mlp(new_points, layers)
    assert pooling == "max"
    return y.max(axis=2)


def sa_attention_tail(new_points, layers, Wq, bq, Wk, bk, Wv, bv, bn_out=None, and_pooling=False):
    """pointnet_sa_module_attention (attention_layer.py:227-263) after sample_and_group: MLP -> AttentionLayer with
    query = sample 0 (:259), heads = mlp[-1] // 4, key_dim = output_dim = 4 (:255-258) -> batch norm (:263); the
    _and_pooling variant (:306-323) adds the max over nsample of the MLP output AFTER that batch norm (:323).
    float64 throughout (attention_layer_f64)."""
    y = shared_mlp(new_points, layers)                       # (B,m,ns,C)
    B, m, ns, C = y.shape
    x = y.reshape(B * m, ns, C)
    att = attention_layer_f64(x, x[:, 0, :], Wq, bq, Wk, bk, Wv, bv, C // 4, 4).reshape(B, m, C)
    if bn_out is not None:
        att = batch_norm_inference(att, *bn_out)
    if and_pooling:
        att = att + y.max(axis=2)
    return att


def attention_layer_f64(x, xq, Wq, bq, Wk, bk, Wv, bv, heads, key_dim):
    """AttentionLayer.call (attention_layer.py:29-45) in float64 numpy with the reference's RAW reshape (:35)."""
    f = np.float64
    x, xq = np.asarray(x, f), np.asarray(xq, f)
    G, S, _ = x.shape
    Q = xq @ np.asarray(Wq, f) + (0 if bq is None else np.asarray(bq, f))
    K = x @ np.asarray(Wk, f) + (0 if bk is None else np.asarray(bk, f))
    V = x @ np.asarray(Wv, f) + (0 if bv is None else np.asarray(bv, f))
    Qh = Q.reshape(G, heads, 1, key_dim)
    Kh = K.reshape(G, heads, S, key_dim)                     # raw reshape of the (S, heads*key_dim) buffer
    Vh = V.reshape(G, heads, S, key_dim)
    w = Qh @ Kh.transpose(0, 1, 3, 2) / np.sqrt(float(key_dim))
    w = np.exp(w - w.max(-1, keepdims=True))
    w = w / w.sum(-1, keepdims=True)
    return (w @ Vh).reshape(G, heads * key_dim)


# ---- the experimental layers (attention_layer.py:48-210), float64 numpy restatements ---------------------------------
def dense_f64(x, W, b, relu=False):
    y = np.asarray(x, np.float64) @ np.asarray(W, np.float64) + (0 if b is None else np.asarray(b, np.float64))
    return np.maximum(y, 0.0) if relu else y


def inner_attention_f64(Q, K, V, key_dim):
    """InnerAttentionLayer.call (:61-75) without its Dense layers: Q, K, V (..., 5 * key_dim)."""
    f = np.float64
    shp = Q.shape
    q, k, v = (np.asarray(t, f).reshape(shp[:-1] + (5, key_dim)) for t in (Q, K, V))
    w = q @ np.swapaxes(k, -1, -2) / np.sqrt(float(key_dim))
    w = np.exp(w - w.max(-1, keepdims=True))
    w = w / w.sum(-1, keepdims=True)
    return (w @ v).reshape(shp)


def attention_contract_f64(Q, K, V, heads, key_dim):
    """AttentionLayer.call :35-42 on given projections: Q (G, H*D), K and V (G, S, H*D), raw reshape to heads."""
    f = np.float64
    G, S, HD = K.shape
    Qh = np.asarray(Q, f).reshape(G, heads, 1, key_dim)
    Kh = np.asarray(K, f).reshape(G, heads, S, key_dim)
    Vh = np.asarray(V, f).reshape(G, heads, S, key_dim)
    w = Qh @ Kh.transpose(0, 1, 3, 2) / np.sqrt(float(key_dim))
    w = np.exp(w - w.max(-1, keepdims=True))
    w = w / w.sum(-1, keepdims=True)
    return (w @ Vh).reshape(G, HD)
