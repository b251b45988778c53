"""The reference's experimental attention layers, composed from this library's kernels.

Mirrors attention_points/attention_scannet/attention_layer.py:48-210 and pooling_attention_layer.py:6-46 -- layers no
shipped model reaches (train.py builds pointnet_sa_module_attention only), kept for completeness of the op surface:

* ``InnerAttentionLayer``      (:48-78)   attention ACROSS the 5 heads of one point + out_net
* ``FeedForwardLayer``         (:81-104)  Dense-ReLU x3 + Dense
* ``InnerAttentionBlock``      (:107-124)
* ``AttentionNetLayer``        (:127-168) sample_and_group -> inner blocks -> AttentionLayer(out_dim, key_dim=out_dim, 16 heads)
* ``AttentionNetMLPLayer``     (:171-210)
* ``PoolingAttentionNetLayer`` (pooling_attention_layer.py:6-46) query = new_xyz

Every Dense runs on the tcgen05 engine (``sa_modules.dense_layer``), the per-neighbourhood contraction on
``pc_attention_fwd`` (key_dim up to 64), the 5 x 5 inner contraction on ``pc_inner_attention_fwd`` (forward only).
``torch.nn.Linear`` is the parameter container.  Dense layers are created on first use from the input width, like Keras'.
"""
import torch

from . import _lib
from .attention_layer import AttentionLayer
from .pointnet_util import sample_and_group
from .sa_modules import SharedMLP, dense_layer


class _LazyDense(torch.nn.Module):
    """tf.layers.Dense(units): the kernel is built on first call from the input's last dimension."""

    def __init__(self, units):
        super().__init__()
        self.units, self.lin = units, None

    def forward(self, x, relu=False):
        if self.lin is None:
            self.lin = torch.nn.Linear(x.shape[-1], self.units).to(x.device)
        return dense_layer(x, self.lin.weight, self.lin.bias, relu=relu, linear_layout=True)


class _InnerContract(torch.autograd.Function):
    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, Q, K, V, key_dim):
        rows = Q.numel() // (5 * key_dim)
        out = torch.empty_like(Q)
        rc = _lib.lib().pc_inner_attention_fwd(rows, key_dim, _lib.ptr(Q), _lib.ptr(K), _lib.ptr(V), _lib.ptr(out),
                                               _lib.stream())
        _lib.check(rc, "pc_inner_attention_fwd")
        return out

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError("the inner 5 x 5 contraction of the experimental layers is forward-only here")


class InnerAttentionLayer(torch.nn.Module):
    """attention_layer.py:48-78 (num_heads fixed to 5, :53)."""

    def __init__(self, output_dim, key_dim):
        super().__init__()
        self.output_dim, self.key_dim, self.num_heads = output_dim, key_dim, 5
        self.query_net, self.key_net, self.value_net = (_LazyDense(key_dim * 5) for _ in range(3))
        self.out_net = _LazyDense(output_dim)

    def forward(self, x):
        Q, K, V = self.query_net(x), self.key_net(x), self.value_net(x)
        att = _InnerContract.apply(_lib.cuda_f32(Q, "Q"), _lib.cuda_f32(K, "K"), _lib.cuda_f32(V, "V"), self.key_dim)
        return self.out_net(att)


class FeedForwardLayer(torch.nn.Module):
    """attention_layer.py:81-104 (dropout is the identity at inference and rate 0 by default)."""

    def __init__(self, input_and_output_dim, inner_dim, dropout=0):
        super().__init__()
        self.layer_1, self.layer_2, self.layer_3 = (_LazyDense(inner_dim) for _ in range(3))
        self.layer_4 = _LazyDense(input_and_output_dim)

    def forward(self, x):
        x = self.layer_1(x, relu=True)
        x = self.layer_2(x, relu=True)
        x = self.layer_3(x, relu=True)
        return self.layer_4(x)


class InnerAttentionBlock(torch.nn.Module):
    """attention_layer.py:107-124."""

    def __init__(self, out_dim, key_dim):
        super().__init__()
        self.attention_layer = InnerAttentionLayer(out_dim, key_dim)
        self.feed_forward_layer = FeedForwardLayer(out_dim, out_dim)
        self.pre_feed_forward_layer = FeedForwardLayer(out_dim, out_dim)

    def forward(self, points):
        points = self.pre_feed_forward_layer(points)
        points = self.attention_layer(points)
        return self.feed_forward_layer(points) + points


class AttentionNetLayer(torch.nn.Module):
    """attention_layer.py:127-168: returns [new_xyz, new_points (B, npoint, 16 * out_dim), idx]."""

    def __init__(self, npoint, out_dim, inner_dimensions, radius=0.1, nsample=32):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.attention_layer = AttentionLayer(out_dim, out_dim)           # 16 heads (default, :11)
        self.inner_blocks = torch.nn.ModuleList([InnerAttentionBlock(i, out_dim) for i in inner_dimensions])

    def forward(self, inputs):
        xyz, points = inputs
        new_xyz, new_points, idx, _ = sample_and_group(self.npoint, self.radius, self.nsample, xyz, points, False, True)
        for block in self.inner_blocks:
            new_points = block(new_points)
        new_points = self.attention_layer([new_points, new_points[:, :, :1, :]])
        return [new_xyz, new_points, idx]


class AttentionNetMLPLayer(torch.nn.Module):
    """attention_layer.py:171-210."""

    def __init__(self, npoint, out_dim, inner_dimensions, radius=0.1, nsample=32):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.attention_layer = AttentionLayer(out_dim, out_dim)
        self.inner_blocks = torch.nn.ModuleList([FeedForwardLayer(i, i) for i in inner_dimensions])

    def forward(self, inputs):
        xyz, points = inputs
        new_xyz, new_points, idx, _ = sample_and_group(self.npoint, self.radius, self.nsample, xyz, points, False, True)
        for block in self.inner_blocks[:-1]:
            new_points = torch.relu(block(new_points))
        new_points = self.inner_blocks[-1](new_points)
        new_points = self.attention_layer([new_points, new_points[:, :, :1, :]])
        return [new_xyz, new_points, idx]


class PoolingAttentionNetLayer(torch.nn.Module):
    """pooling_attention_layer.py:6-46: shared MLP, then AttentionLayer with the centroid coordinates as the query."""

    def __init__(self, in_channels, mlp, npoint, out_dim, radius=0.1, nsample=32, bn=True):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlp = SharedMLP(in_channels + 3, mlp, bn)
        self.attention_layer = AttentionLayer(out_dim, out_dim)

    def forward(self, inputs):
        xyz, points = inputs
        with torch.no_grad():
            new_xyz, new_points, idx, _ = sample_and_group(self.npoint, self.radius, self.nsample, xyz, points, False, True)
            new_points = self.mlp(new_points)
            layer = self.attention_layer
            if layer.query_net is None:   # the query is 3-wide, the keys / values mlp[-1]-wide: separate input widths
                hd = layer.key_dim * layer.num_heads
                layer.query_net = torch.nn.Linear(3, hd).to(xyz.device)
                layer.key_net = torch.nn.Linear(new_points.shape[-1], hd).to(xyz.device)
                layer.value_net = torch.nn.Linear(new_points.shape[-1], hd).to(xyz.device)
            new_points = layer([new_points, new_xyz.unsqueeze(2)])
        return new_xyz, new_points, idx
