"""the reference's experimental attention layers (attention_layer.py:48-210,
pooling_attention_layer.py:6-46) on this library's kernels, against float64 numpy restatements on the same weights.
Not reachable from the reference's train.py; covered for completeness of the op surface."""
import numpy as np
import pytest
import torch

import pcops_b200 as ops
from oracle import cpu, synth
from pcops_b200 import experimental_layers as ex
from tests.test_gpu_dense import check

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


def lin(d):
    """(W (in, out), b) of a lazily built Dense."""
    return npy(d.lin.weight).T, npy(d.lin.bias)


@pytest.mark.parametrize("G,S,H,D", [(50, 32, 16, 32), (33, 32, 16, 64), (7, 20, 16, 8), (40, 32, 16, 16)])
def test_contraction_with_wide_heads(G, S, H, D):
    """AttentionLayer(out_dim, key_dim=out_dim, heads=16) of the experimental layers: key_dim up to 64."""
    rng = np.random.default_rng(G + D)
    Q = rng.standard_normal((G, H * D), dtype=np.float32)
    K = rng.standard_normal((G, S, H * D), dtype=np.float32)
    V = rng.standard_normal((G, S, H * D), dtype=np.float32)
    q, k, v = (cu(t).requires_grad_(True) for t in (Q, K, V))
    out = ops.attention_contract(q, k, v, H, D)
    np.testing.assert_allclose(npy(out), cpu.attention_fwd(Q, K, V, H, D), rtol=1e-5, atol=2e-6)
    g = rng.standard_normal((G, H * D), dtype=np.float32)
    out.backward(cu(g))
    dQ, dK, dV = cpu.attention_bwd(Q, K, V, g, H, D)
    for got, want in ((q.grad, dQ), (k.grad, dK), (v.grad, dV)):
        np.testing.assert_allclose(npy(got), want.reshape(got.shape), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kd", [4, 8, 16, 32, 64])
def test_inner_attention_contraction(kd):
    rng = np.random.default_rng(kd)
    Q, K, V = (rng.standard_normal((3, 40, 32, 5 * kd), dtype=np.float32) for _ in range(3))
    got = ex._InnerContract.apply(cu(Q), cu(K), cu(V), kd)
    check(got, cpu.inner_attention_f64(Q, K, V, kd), "inner attention kd=%d" % kd)


def ff64(x, layer):
    for d, relu in ((layer.layer_1, True), (layer.layer_2, True), (layer.layer_3, True), (layer.layer_4, False)):
        x = cpu.dense_f64(x, *lin(d), relu=relu)
    return x


def outer64(x, layer, query):
    C = layer.key_dim * layer.num_heads
    Wq, bq = npy(layer.query_net.weight).T, npy(layer.query_net.bias)
    Wk, bk = npy(layer.key_net.weight).T, npy(layer.key_net.bias)
    Wv, bv = npy(layer.value_net.weight).T, npy(layer.value_net.bias)
    B, m, S, _ = x.shape
    Q = cpu.dense_f64(query, Wq, bq).reshape(B * m, C)
    K = cpu.dense_f64(x, Wk, bk).reshape(B * m, S, C)
    V = cpu.dense_f64(x, Wv, bv).reshape(B * m, S, C)
    return cpu.attention_contract_f64(Q, K, V, layer.num_heads, layer.key_dim).reshape(B, m, C)


def grouped(xyz, feats, m, r):
    fi = cpu.farthest_point_sample(m, xyz)
    nx = cpu.gather_point(xyz, fi)
    idx, _ = cpu.query_ball_point(r, 32, xyz, nx)
    return nx, idx, np.concatenate([cpu.group_point(xyz, idx) - nx[:, :, None, :], cpu.group_point(feats, idx)], -1)


def test_attention_net_layer_and_mlp_layer():
    torch.manual_seed(0)
    xyz, feats = synth.scannet_batch(70, 2, 2048)
    nx, idx, new_points = grouped(xyz, feats, 64, 0.3)
    with torch.no_grad():
        layer = ex.AttentionNetLayer(64, 8, [16, 16], radius=0.3).to(DEV)
        new_xyz, out, got_idx = layer([cu(xyz), cu(feats)])
        assert np.array_equal(npy(new_xyz), nx) and np.array_equal(npy(got_idx), idx)
        x = new_points.astype(np.float64)
        for blk in layer.inner_blocks:
            p = ff64(x, blk.pre_feed_forward_layer)
            a = blk.attention_layer
            att = cpu.inner_attention_f64(cpu.dense_f64(p, *lin(a.query_net)), cpu.dense_f64(p, *lin(a.key_net)),
                                          cpu.dense_f64(p, *lin(a.value_net)), a.key_dim)
            p = cpu.dense_f64(att, *lin(a.out_net))
            x = ff64(p, blk.feed_forward_layer) + p
        check(out, outer64(x, layer.attention_layer, x[:, :, :1, :]), "AttentionNetLayer", scale_tol=5e-5, rel_tol=2e-4)

        mlp_layer = ex.AttentionNetMLPLayer(64, 16, [32, 32], radius=0.3).to(DEV)
        _, out2, _ = mlp_layer([cu(xyz), cu(feats)])
        x = new_points.astype(np.float64)
        for blk in mlp_layer.inner_blocks[:-1]:
            x = np.maximum(ff64(x, blk), 0.0)
        x = ff64(x, mlp_layer.inner_blocks[-1])
        check(out2, outer64(x, mlp_layer.attention_layer, x[:, :, :1, :]), "AttentionNetMLPLayer", scale_tol=5e-5, rel_tol=2e-4)


def test_pooling_attention_net_layer():
    torch.manual_seed(1)
    xyz, feats = synth.scannet_batch(80, 2, 2048)
    nx, idx, new_points = grouped(xyz, feats, 64, 0.3)
    layer = ex.PoolingAttentionNetLayer(6, [32, 64], 64, 4, radius=0.3).to(DEV)
    new_xyz, out, got_idx = layer([cu(xyz), cu(feats)])
    assert np.array_equal(npy(got_idx), idx)
    x = new_points.astype(np.float64)
    for conv in layer.mlp.layers:
        x = cpu.conv2d_1x1(x, npy(conv.weights), npy(conv.biases),
                           (npy(conv.gamma), npy(conv.beta), npy(conv.moving_mean), npy(conv.moving_variance)), True)
    check(out, outer64(x, layer.attention_layer, nx[:, :, None, :].astype(np.float64)), "PoolingAttentionNetLayer",
          scale_tol=2e-5, rel_tol=1e-4)
