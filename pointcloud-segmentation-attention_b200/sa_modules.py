"""The set-abstraction layer tails on the tensor cores: shared MLP -> max-pool / attention.

Mirrors the part of the reference's layer library that follows ``sample_and_group``:

* ``pointnet_sa_module``                        pointnet2_tensorflow/utils/pointnet_util.py:88-162
* ``pointnet_sa_module_attention``              attention_points/attention_scannet/attention_layer.py:213-281
* ``pointnet_sa_module_attention_and_pooling``  attention_layer.py:284-345

In the reference every 1x1 ``tf_util.conv2d`` (conv -> bias -> batch norm -> ReLU, utils/tf_util.py:120-186) is a
cuDNN call with its activation tensor in HBM, the max over nsample another op.  Here a layer is one tcgen05 kernel
(csrc/gemm_tf32.cu, 3xTF32 split precision): inference-mode batch norm is folded into the layer's kernel and bias, ReLU
and -- for the last layer -- the max over the 32 samples of a neighbourhood run in the TMEM epilogue, so the widest
activation of a level is never written.  Parameters keep TensorFlow's layouts (conv kernel [1,1,cin,cout] squeezed to
(cin, cout); Dense kernel (cin, cout)), so a checkpoint of the reference maps name by name.

Inference only (``is_training=False``): batch norm with batch statistics needs a reduction over the whole batch between
layers and is out of scope here; training-time Dense gradients are in ``attention_layer.py``.
"""
import torch

from . import _lib
from .attention_layer import PreparedAttentionLayer
from .pointnet_util import sample_and_group, sample_and_group_all

BN_EPSILON = 1e-3   # tf.contrib.layers.batch_norm default, as used by tf_util.batch_norm_template (tf_util.py:512-530)


class DenseImage:
    """The tensor-core weight image of one (K, N) layer (TF32 hi / lo parts, 128-byte-swizzled K blocks), built ONCE per
    set of weights and reused by every call -- the image depends on the weights only."""

    def __init__(self, weight, bias=None, transpose=False):
        w = _lib.cuda_f32(weight.detach(), "weight")
        if w.dim() != 2:
            raise ValueError("a Dense / 1x1 conv kernel is (cin, cout)")
        self.K, self.N = (w.shape[1], w.shape[0]) if transpose else (w.shape[0], w.shape[1])
        L = _lib.lib()
        nbytes = L.pc_dense_image_bytes(self.K, self.N)
        self.image = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        self.bias = None if bias is None else _lib.cuda_f32(bias.detach(), "bias").clone()
        with torch.cuda.device(w.device):
            _lib.check(L.pc_dense_prepare(self.K, self.N, _lib.ptr(w), 1 if transpose else 0, _lib.ptr(self.image),
                                          _lib.stream()), "pc_dense_prepare")


@_lib.on_tensor_device
def dense(x, image, relu=False):
    """x (..., K) -> act(x . W + b) (..., N) on the tensor cores.  ``image`` is a DenseImage."""
    x = _lib.cuda_f32(x.detach(), "x")
    if x.shape[-1] != image.K:
        raise ValueError("dense expects (..., %d) input, got %s" % (image.K, tuple(x.shape)))
    rows = x.numel() // image.K
    y = torch.empty(x.shape[:-1] + (image.N,), dtype=torch.float32, device=x.device)
    rc = _lib.lib().pc_dense_fwd(rows, image.K, image.N, _lib.ptr(x), image.K, _lib.ptr(image.image), _lib.ptr(image.bias),
                                 1 if relu else 0, _lib.ptr(y), image.N, _lib.stream())
    _lib.check(rc, "pc_dense_fwd")
    return y


@_lib.on_tensor_device
def dense_max_pool(x, image, relu=True, keep_full=False):
    """x (..., 32, K) -> max over the 32 samples of act(x . W + b): (..., N); with keep_full also the un-pooled
    (..., 32, N) activation."""
    x = _lib.cuda_f32(x.detach(), "x")
    if x.dim() < 2 or x.shape[-1] != image.K or x.shape[-2] != 32:
        raise ValueError("dense_max_pool expects (..., 32, %d) input, got %s" % (image.K, tuple(x.shape)))
    groups = x.numel() // (32 * image.K)
    pooled = torch.empty(x.shape[:-2] + (image.N,), dtype=torch.float32, device=x.device)
    full = torch.empty(x.shape[:-1] + (image.N,), dtype=torch.float32, device=x.device) if keep_full else None
    rc = _lib.lib().pc_dense_pool_fwd(groups, 32, image.K, image.N, _lib.ptr(x), image.K, _lib.ptr(image.image),
                                      _lib.ptr(image.bias), 1 if relu else 0, _lib.ptr(full), image.N, _lib.ptr(pooled),
                                      image.N, _lib.stream())
    _lib.check(rc, "pc_dense_pool_fwd")
    return (pooled, full) if keep_full else pooled


@_lib.on_tensor_device
def dense_weight_grad(x, dy, want_bias=True):
    """dW (K, N) = x^T . dy and db (N) = column sums of dy for x (..., K), dy (..., N) -- pc_dense_bwd_weight."""
    x = _lib.cuda_f32(x.detach(), "x")
    dy = _lib.cuda_f32(dy.detach(), "dy")
    K, N = x.shape[-1], dy.shape[-1]
    rows = x.numel() // K
    if dy.numel() // N != rows:
        raise ValueError("dense_weight_grad expects x (..., K) and dy (..., N) over the same rows")
    L = _lib.lib()
    dw = torch.empty((K, N), dtype=torch.float32, device=x.device)
    db = torch.empty(N, dtype=torch.float32, device=x.device) if want_bias else None
    ws = _lib.workspace(L.pc_dense_bwd_weight_workspace_bytes(rows, K, N), x.device)
    rc = L.pc_dense_bwd_weight(rows, K, N, _lib.ptr(x), K, _lib.ptr(dy), N, _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws),
                               _lib.stream())
    _lib.check(rc, "pc_dense_bwd_weight")
    return dw, db


class _DenseFn(torch.autograd.Function):
    """y = act(x . W + b) with every product of the forward and the backward on the tcgen05 engine:
    dx = (dy * act') . W^T (transposed weight image), dW = x^T . (dy * act'), db = column sums.
    ``linear_layout``: W is stored (cout, cin) as torch.nn.Linear keeps it (the image is built with the matching strides,
    nothing is copied); otherwise (cin, cout) as TensorFlow's Dense / conv kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, linear_layout):
        image = DenseImage(weight, bias, transpose=linear_layout)
        y = dense(x, image, relu)
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.cfg = (relu, linear_layout, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        relu, linear_layout, has_bias = ctx.cfg
        dy = _lib.cuda_f32(dy, "dy")
        if relu:
            dy = dy * (y > 0).to(dy.dtype)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # forward image reads W as (K=cin, N=cout); the input gradient needs (K=cout, N=cin) of the same memory
            dx = dense(dy, DenseImage(weight, None, transpose=not linear_layout))
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            dw, db = dense_weight_grad(x, dy, want_bias=has_bias)       # (cin, cout)
            if linear_layout:
                dw = dw.t()
        return dx, dw, (db if has_bias else None), None, None


def dense_layer(x, weight, bias=None, relu=False, linear_layout=False):
    """Differentiable Dense / 1x1-conv layer on the tensor cores (forward and backward)."""
    return _DenseFn.apply(_lib.cuda_f32(x, "x"), _lib.cuda_f32(weight, "weight"),
                          None if bias is None else _lib.cuda_f32(bias, "bias"), bool(relu), bool(linear_layout))


def fold_batch_norm(weight, bias, gamma, beta, moving_mean, moving_var, eps=BN_EPSILON):
    """conv -> bias_add -> batch norm (inference) as ONE affine layer: W' = W * s, b' = (b - mean) * s + beta with
    s = gamma / sqrt(var + eps) (tf_util.py:171-181 with is_training False)."""
    s = gamma / torch.sqrt(moving_var + eps)
    b = torch.zeros_like(beta) if bias is None else bias
    return weight * s.unsqueeze(0), (b - moving_mean) * s + beta


class Conv2d1x1(torch.nn.Module):
    """tf_util.conv2d(inputs, cout, [1,1], padding='VALID', stride=[1,1], bn=bn) -- utils/tf_util.py:120-186.
    Parameters in TF layout: ``weights`` (cin, cout), ``biases`` (cout), and the batch-norm variables gamma / beta /
    moving_mean / moving_variance."""

    def __init__(self, cin, cout, bn=True, relu=True):
        super().__init__()
        self.weights = torch.nn.Parameter(torch.empty(cin, cout))
        torch.nn.init.xavier_uniform_(self.weights)            # use_xavier=True (tf_util.py:127)
        self.biases = torch.nn.Parameter(torch.zeros(cout))      # constant_initializer(0.0) (:171)
        self.bn, self.relu = bn, relu
        if bn:
            self.gamma = torch.nn.Parameter(torch.ones(cout))
            self.beta = torch.nn.Parameter(torch.zeros(cout))
            self.register_buffer("moving_mean", torch.zeros(cout))
            self.register_buffer("moving_variance", torch.ones(cout))
        self._image, self._key = None, None

    def image(self):
        """The folded tensor-core image, rebuilt only when a parameter changed (torch's version counters)."""
        ts = [self.weights, self.biases] + ([self.gamma, self.beta, self.moving_mean, self.moving_variance] if self.bn else [])
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if self._image is None or key != self._key:
            with torch.no_grad():
                if self.bn:
                    w, b = fold_batch_norm(self.weights, self.biases, self.gamma, self.beta, self.moving_mean,
                                           self.moving_variance)
                else:
                    w, b = self.weights, self.biases
                self._image, self._key = DenseImage(w.contiguous(), b.contiguous()), key
        return self._image

    def forward(self, x):
        return dense(x, self.image(), self.relu)


class SharedMLP(torch.nn.Module):
    """The 'Point Feature Embedding' loop: conv0, conv1, ... over (B, npoint, nsample, C) (pointnet_util.py:119-131)."""

    def __init__(self, cin, mlp, bn=True):
        super().__init__()
        layers, c = [], cin
        for cout in mlp:
            layers.append(Conv2d1x1(c, cout, bn=bn, relu=True))
            c = cout
        self.layers = torch.nn.ModuleList(layers)
        self.out_channels = c

    def forward(self, new_points, pooling=None):
        """pooling None -> (B,m,ns,C_out); 'max' -> (B,m,C_out) with the last layer's activation never written;
        'both' -> (pooled, full)."""
        x = new_points
        for layer in self.layers[:-1]:
            x = layer(x)
        last = self.layers[-1]
        if pooling is None:
            return last(x)
        if x.shape[-2] != 32:   # the pooled epilogue covers nsample = 32 (every level of the ScanNet models)
            full = last(x)
            pooled = full.max(dim=-2).values
            return pooled if pooling == "max" else (pooled, full)
        if pooling == "max":
            return dense_max_pool(x, last.image(), last.relu)
        if pooling == "both":
            return dense_max_pool(x, last.image(), last.relu, keep_full=True)
        raise ValueError("pooling must be None, 'max' or 'both'")


def _group(xyz, points, npoint, radius, nsample, group_all, knn, use_xyz):
    if group_all:
        return sample_and_group_all(xyz, points, use_xyz)
    return sample_and_group(npoint, radius, nsample, xyz, points, knn, use_xyz)


class PointnetSAModule(torch.nn.Module):
    """pointnet_sa_module (pointnet_util.py:88-162), pooling='max', inference: returns (new_xyz, new_points, idx)."""

    def __init__(self, npoint, radius, nsample, in_channels, mlp, mlp2=None, group_all=False, bn=True, knn=False,
                 use_xyz=True):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.group_all, self.knn, self.use_xyz = group_all, knn, use_xyz
        self.mlp = SharedMLP(in_channels + (3 if use_xyz else 0), mlp, bn)
        self.mlp2 = SharedMLP(self.mlp.out_channels, mlp2, bn) if mlp2 else None

    def forward(self, xyz, points):
        with torch.no_grad():
            new_xyz, new_points, idx, _ = _group(xyz, points, self.npoint, self.radius, self.nsample, self.group_all,
                                                 self.knn, self.use_xyz)
            out = self.mlp(new_points, pooling="max")                      # (B, npoint, mlp[-1])
            if self.mlp2 is not None:
                out = self.mlp2(out)
        return new_xyz, out, idx


class PointnetSAModuleAttention(torch.nn.Module):
    """pointnet_sa_module_attention (attention_layer.py:213-281) and, with and_pooling=True,
    pointnet_sa_module_attention_and_pooling (:284-345): shared MLP -> AttentionLayer(output_dim=4, key_dim=4,
    heads=mlp[-1]//4) with the group's first sample as the query (:259) -> batch norm (:263) [+ max over nsample of the
    MLP output, added AFTER the batch norm (:323)] -> optional mlp2.  Inference."""

    def __init__(self, npoint, radius, nsample, in_channels, mlp, mlp2=None, group_all=False, bn=True, knn=False,
                 use_xyz=True, and_pooling=False):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.group_all, self.knn, self.use_xyz, self.and_pooling = group_all, knn, use_xyz, and_pooling
        self.mlp = SharedMLP(in_channels + (3 if use_xyz else 0), mlp, bn)
        C = self.mlp.out_channels
        if C % 4:
            raise ValueError("heads = mlp[-1] // 4 (attention_layer.py:255-256): mlp[-1] must be a multiple of 4")
        self.query_net = torch.nn.Linear(C, C)     # tf.layers.Dense(key_dim * heads); weight kept as (out, in) by torch
        self.key_net = torch.nn.Linear(C, C)
        self.value_net = torch.nn.Linear(C, C)
        self.gamma = torch.nn.Parameter(torch.ones(C))          # batch_norm_for_conv2d after the attention (:263)
        self.beta = torch.nn.Parameter(torch.zeros(C))
        self.register_buffer("moving_mean", torch.zeros(C))
        self.register_buffer("moving_variance", torch.ones(C))
        self.mlp2 = SharedMLP(C, mlp2, bn) if mlp2 else None
        self._attention = PreparedAttentionLayer()

    def forward(self, xyz, points):
        with torch.no_grad():
            new_xyz, new_points, idx, _ = _group(xyz, points, self.npoint, self.radius, self.nsample, self.group_all,
                                                 self.knn, self.use_xyz)
            if self.and_pooling:
                pooled, feats = self.mlp(new_points, pooling="both")
            else:
                pooled, feats = None, self.mlp(new_points)                 # (B, npoint, nsample, C)
            B, m, ns, C = feats.shape
            x = feats.reshape(B * m, ns, C)
            att = self._attention(x[:, 0, :].contiguous(), x, self.query_net, self.key_net, self.value_net).reshape(B, m, C)
            s = self.gamma / torch.sqrt(self.moving_variance + BN_EPSILON)
            out = (att - self.moving_mean) * s + self.beta
            if pooled is not None:
                out = out + pooled
            if self.mlp2 is not None:
                out = self.mlp2(out)
        return new_xyz, out, idx
