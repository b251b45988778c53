"""Parity of the CUDA path (through the C ABI / the reference-named wrappers) against the CPU oracle, the committed
golden fixtures (reference outputs) and -- when oracle/_ref/libref_gpu_nofma.so travelled to the box -- the
reference's own CUDA kernels.  Bit-exact for every index / gathered value / gradient / interpolation; 1e-5 relative
for the attention contraction (tolerance from BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

import pcops_b200 as ops
from oracle import cpu, ref, synth
from tests.test_oracle_pins import radius_probe

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


def same(t, a):
    return np.array_equal(npy(t), a)


@pytest.fixture(autouse=True, params=["grid", "allpairs"])
def neighbour_search_mode(request):
    """Every test runs twice: with the cell-grid ball query / three_nn (the default path) and with the all-pairs
    kernels behind the reference launchers' exact signatures.  Both must match the oracle bit for bit."""
    from pcops_b200 import tf_grouping, tf_interpolate
    saved = tf_grouping.USE_GRID, tf_interpolate.USE_GRID
    tf_grouping.USE_GRID = tf_interpolate.USE_GRID = (request.param == "grid")
    yield request.param
    tf_grouping.USE_GRID, tf_interpolate.USE_GRID = saved


@pytest.fixture(scope="module")
def refgpu():
    if not ref.available_gpu(True):
        pytest.skip("oracle/_ref/libref_gpu_nofma.so not present")
    return ref.Gpu(nofma=True)


# ------------------------------------------------------------------------------------------------------ FPS (a1)
@pytest.mark.parametrize("n,m", [(64, 16), (256, 64), (511, 40), (512, 40), (513, 40), (1024, 256), (3072, 100),
                                 (3073, 100), (8192, 1024), (5000, 200), (1, 1), (2, 2), (33, 7)])
def test_fps_matches_oracle(n, m):
    xyz, _ = synth.scannet_batch(n + m, 3, n)
    assert same(ops.farthest_point_sample(m, cu(xyz)), cpu.farthest_point_sample(m, xyz))


def test_fps_gather_fused_equals_the_pair():
    for n, m in ((8192, 1024), (1024, 256), (200, 64), (20000, 50)):
        xyz, _ = synth.scannet_batch(n, 2, n)
        x = cu(xyz)
        idx, new_xyz = ops.farthest_point_sample_and_gather(m, x)
        assert torch.equal(idx, ops.farthest_point_sample(m, x))
        assert torch.equal(new_xyz, ops.gather_point(x, idx))
        assert same(idx, cpu.farthest_point_sample(m, xyz))


def test_fps_ties_duplicates_and_big_batches():
    rs = np.random.RandomState(11)
    xyz = (rs.randint(0, 5, size=(40, 700, 3)) / 4.0).astype(np.float32)   # b > 32 (reference grid-strides at 32)
    assert same(ops.farthest_point_sample(60, cu(xyz)), cpu.farthest_point_sample(60, xyz))
    ones = np.ones((2, 100, 3), np.float32)
    assert same(ops.farthest_point_sample(9, cu(ones)), np.zeros((2, 9), np.int32))
    small = synth.uniform_cube(3, 2, 10, 3)
    assert same(ops.farthest_point_sample(16, cu(small)), cpu.farthest_point_sample(16, small))   # m > n


# --------------------------------------------------------------------------------- prob_sample (a0, op surface only)
@pytest.mark.parametrize("n,m,b", [(1, 5, 2), (3, 7, 1), (4, 9, 2), (5, 16, 3), (100, 64, 2), (4095, 100, 2),
                                   (8192, 300, 2), (8193, 300, 2), (8197, 64, 1), (40000, 1000, 3)])
def test_prob_sample_matches_oracle(n, m, b):
    """Cumulative sums in the reference's blocked summation tree (bit-exact), then its power-of-two descent."""
    rs = np.random.RandomState(n + m)
    w = rs.random_sample((b, n)).astype(np.float32)
    w[:, rs.permutation(n)[: n // 5]] = 0.0              # zero-weight categories: plateaus in the cumulative sums
    r = rs.random_sample((b, m)).astype(np.float32)
    r[:, 0] = 0.0
    assert same(ops.cumsum(cu(w)), cpu.cumsum(w))
    assert same(ops.prob_sample(cu(w), cu(r)), cpu.prob_sample(w, r))


def test_prob_sample_matches_reference_cuda_kernel(refgpu):
    rs = np.random.RandomState(100)                        # the reference smoke scripts' seed (tf_sampling.py:63)
    for b, n, m in ((32, 512, 128), (3, 20000, 999), (2, 8195, 64)):
        w = cu(rs.random_sample((b, n)).astype(np.float32))
        r = cu(rs.random_sample((b, m)).astype(np.float32))
        out, temp = refgpu.prob_sample(w, r)
        assert torch.equal(ops.cumsum(w), temp)
        assert torch.equal(ops.prob_sample(w, r), out)


@pytest.mark.parametrize("n,m,b", [(8193, 64, 2), (16384, 80, 3), (16385, 80, 1), (20000, 96, 2), (40000, 64, 2),
                                   (70000, 48, 1), (131072, 40, 1), (140000, 24, 1), (262144, 20, 1), (270000, 12, 1)])
def test_fps_cluster_and_streaming_paths_above_8192_points(n, m, b):
    """8192 < n <= 262144: one thread-block cluster (2/4/8/16 CTAs, DSMEM exchange) per scene; beyond: streaming."""
    xyz, _ = synth.scannet_batch(n % 1000, b, n)
    x = cu(xyz)
    idx, new_xyz = ops.farthest_point_sample_and_gather(m, x)
    assert same(idx, cpu.farthest_point_sample(m, xyz))
    assert torch.equal(new_xyz, ops.gather_point(x, idx))


def test_fps_cluster_ties_across_ctas():
    """Exact duplicates placed in different 8192-point slices: the (k mod 512, k) tie-break must hold cluster-wide."""
    rng = np.random.default_rng(3)
    base = rng.random((1, 600, 3)).astype(np.float32)
    xyz = np.tile(base, (1, 30, 1))[:, :17000]          # every point repeated ~28 times across slices
    assert same(ops.farthest_point_sample(300, cu(xyz)), cpu.farthest_point_sample(300, xyz))
    xyz = np.tile(base, (1, 240, 1))[:, :140000]        # 16384-point slices, half of each read from shared memory
    assert same(ops.farthest_point_sample(200, cu(xyz)), cpu.farthest_point_sample(200, xyz))


def test_fps_matches_reference_cuda_kernel(refgpu):
    xyz, _ = synth.scannet_batch(5, 4, 8192)
    x = cu(xyz)
    assert torch.equal(ops.farthest_point_sample(1024, x), refgpu.farthest_point_sample(1024, x))
    x = cu(synth.uniform_cube(100, 33, 512, 3))
    assert torch.equal(ops.farthest_point_sample(128, x), refgpu.farthest_point_sample(128, x))


# --------------------------------------------------------------------------------------------- gather_point (a2,a3)
def test_gather_point_and_grad():
    xyz, _ = synth.scannet_batch(1, 3, 1000)
    idx = cpu.farthest_point_sample(77, xyz)
    idx[:, 5] = idx[:, 3]                                           # repeated index -> two contributions
    out = ops.gather_point(cu(xyz), cu(idx))
    assert same(out, cpu.gather_point(xyz, idx))
    og = synth.features(4, 3, 77, 3)
    assert same(ops.gather_point_grad(cu(xyz), cu(idx), cu(og)), cpu.gather_point_grad(xyz, idx, og))
    x = cu(xyz).requires_grad_(True)
    ops.gather_point(x, cu(idx)).backward(cu(og))
    assert same(x.grad, cpu.gather_point_grad(xyz, idx, og))


def test_gather_point_matches_reference_cuda_kernel(refgpu):
    xyz, _ = synth.scannet_batch(2, 5, 2048)
    idx = cu(cpu.farthest_point_sample(300, xyz))
    assert torch.equal(ops.gather_point(cu(xyz), idx), refgpu.gather_point(cu(xyz), idx))


# ------------------------------------------------------------------------------------------------ ball query (a4)
def test_ball_query_golden(golden):
    g = golden("grouping")
    for r, ns in ((0.1, 64), (0.2, 8), (0.4, 32)):
        idx, cnt = ops.query_ball_point(r, ns, cu(g["xyz1"]), cu(g["xyz2"]))
        assert same(idx, g["idx_r%g_ns%d" % (r, ns)])
        oi, oc = cpu.query_ball_point(r, ns, g["xyz1"], g["xyz2"])
        assert same(cnt, oc)


@pytest.mark.parametrize("n,m,r,ns", [(8192, 1024, 0.1, 32), (1024, 256, 0.2, 32), (256, 64, 0.4, 32),
                                      (64, 16, 0.8, 32), (2500, 333, 0.3, 5), (100, 7, 10.0, 64), (3000, 50, 1e-4, 8)])
def test_ball_query_matches_oracle(n, m, r, ns):
    xyz, _ = synth.scannet_batch(n, 2, n)
    new_xyz = cpu.gather_point(xyz, cpu.farthest_point_sample(m, xyz))
    idx, cnt = ops.query_ball_point(r, ns, cu(xyz), cu(new_xyz))
    oi, oc = cpu.query_ball_point(r, ns, xyz, new_xyz)
    assert same(idx, oi) and same(cnt, oc)


def test_ball_query_radius_boundary_and_empty_balls():
    for r in (0.1, 0.2, 0.4, 0.8, 0.3, 1.7):
        cand, expect, _ = radius_probe(r)
        idx, cnt = ops.query_ball_point(r, cand.shape[1], cu(cand), cu(np.zeros((1, 1, 3), np.float32)))
        assert int(cnt[0, 0]) == expect.sum()
        assert np.array_equal(npy(idx)[0, 0, :expect.sum()], np.flatnonzero(expect))
    xyz = synth.uniform_cube(1, 2, 50, 3)
    far = np.full((2, 3, 3), 50.0, np.float32)
    idx, cnt = ops.query_ball_point(0.5, 6, cu(xyz), cu(far))
    assert int(cnt.sum()) == 0 and int(idx.abs().sum()) == 0


def test_ball_query_matches_reference_cuda_kernel(refgpu):
    xyz, _ = synth.scannet_batch(9, 4, 8192)
    new_xyz = cu(cpu.gather_point(xyz, cpu.farthest_point_sample(1024, xyz)))
    idx, cnt = ops.query_ball_point(0.1, 32, cu(xyz), new_xyz)
    ridx, rcnt = refgpu.query_ball_point(0.1, 32, cu(xyz), new_xyz)   # rows pre-zeroed by the wrapper
    assert torch.equal(idx, ridx) and torch.equal(cnt, rcnt)


# ---------------------------------------------------------------------------------------- group_point (a5, a6)
@pytest.mark.parametrize("c", [1, 3, 6, 9, 32, 64, 96, 128, 192, 256, 512])
def test_group_point_and_grad_match_oracle(c):
    n, m, ns = 600, 130, 32
    xyz = synth.uniform_cube(c, 2, n, 3)
    idx, _ = cpu.query_ball_point(0.25, ns, xyz, xyz[:, :m])
    pts = synth.features(c, 2, n, c)
    assert same(ops.group_point(cu(pts), cu(idx)), cpu.group_point(pts, idx))
    go = synth.features(c + 1, 2, m, ns, c)
    want = cpu.group_point_grad(pts, idx, go)
    assert same(ops.group_point_grad(cu(pts), cu(idx), cu(go)), want)
    p = cu(pts).requires_grad_(True)
    ops.group_point(p, cu(idx)).backward(cu(go))
    assert same(p.grad, want)


def test_group_point_unaligned_buffers_and_golden(golden):
    g = golden("grouping")
    idx = g["idx_r0.2_ns8"]
    assert same(ops.group_point(cu(g["pts"]), cu(idx)), g["group"])
    assert same(ops.group_point_grad(cu(g["pts"]), cu(idx), cu(g["grad_out"])), g["group_grad"])
    # c % 4 == 0 but the base pointer is only 4-byte aligned: the library must pick the scalar path itself
    b, n, c = g["pts"].shape
    raw = torch.zeros(b * n * c + 1, device=DEV)
    raw[1:] = cu(g["pts"]).reshape(-1)
    pts_off = raw[1:].view(b, n, c)
    assert pts_off.data_ptr() % 16 != 0
    assert same(ops.group_point(pts_off, cu(idx)), g["group"])


def test_group_point_matches_reference_cuda_kernels(refgpu):
    n, m, ns, c = 1024, 256, 32, 64
    xyz, _ = synth.scannet_batch(3, 4, n)
    idx, _ = cpu.query_ball_point(0.2, ns, xyz, xyz[:, :m])
    pts, go = cu(synth.features(1, 4, n, c)), cu(synth.features(2, 4, m, ns, c))
    assert torch.equal(ops.group_point(pts, cu(idx)), refgpu.group_point(pts, cu(idx)))
    mine = ops.group_point_grad(pts, cu(idx), go)
    theirs = refgpu.group_point_grad(pts, cu(idx), go)              # float atomics: same sum, undefined order
    torch.testing.assert_close(mine, theirs, rtol=1e-5, atol=1e-5)
    assert torch.equal(mine, ops.group_point_grad(pts, cu(idx), go))   # ours is bit-reproducible


def test_group_point_grad_popular_point_and_big_key_range():
    # every slot points at one row (longest possible segment) and a key range beyond the shared-memory counters
    b, n, m, ns, c = 2, 70000, 40, 16, 8
    idx = np.zeros((b, m, ns), np.int32)
    idx[1] = np.random.RandomState(0).randint(0, n, size=(m, ns))
    go = synth.features(3, b, m, ns, c)
    pts = np.zeros((b, n, c), np.float32)
    assert same(ops.group_point_grad(cu(pts), cu(idx), cu(go)), cpu.group_point_grad(pts, idx, go))


def test_gradients_at_bench_shapes_many_keys_per_warp():
    """The entry-stream reduce gives a warp KW = 1..16 consecutive keys depending on the size of the launch; the
    small shapes above all run with KW = 1.  Bench shapes (B = 16): SA1-like 8192 keys (KW = 16, most rows empty or
    1-5 entries, first-hit keys ~30), SA2 (KW = 2), FP4 interpolation gradient (weighted form, KW = 2)."""
    b = 16
    xyz, _ = synth.scannet_batch(40, b, 8192)
    fi = cpu.farthest_point_sample(1024, xyz, omp=True)
    nx = cpu.gather_point(xyz, fi)
    for (cloud, q, r, c) in ((xyz, nx, 0.1, 64), (nx, nx[:, :256].copy(), 0.2, 64), (nx, nx[:, :256].copy(), 0.2, 128)):
        idx, _ = cpu.query_ball_point(r, 32, cloud, q, omp=True)
        go = synth.features(c, b, q.shape[1], 32, c)
        pts = np.zeros((b, cloud.shape[1], c), np.float32)
        assert same(ops.group_point_grad(cu(pts), cu(idx), cu(go)), cpu.group_point_grad(pts, idx, go, omp=True))
    od, oi = cpu.three_nn(xyz, nx, omp=True)
    ow = cpu.three_weights(od)
    for c in (128, 64, 32):
        go = synth.features(c + 3, b, 8192, c)
        pts = np.zeros((b, 1024, c), np.float32)
        assert same(ops.three_interpolate_grad(cu(pts), cu(oi), cu(ow), cu(go)),
                    cpu.three_interpolate_grad(pts, oi, ow, go, omp=True))


# ------------------------------------------------------------------------------------ selection sort / kNN (a7)
def test_selection_sort_golden_and_ties(golden):
    g = golden("selsort")
    for tag, k in (("0", 3), ("1", 9), ("2", 4)):
        outi, out = ops.select_top_k(k, cu(g["d" + tag]))
        assert same(outi, g["i" + tag]) and same(out, g["v" + tag])
    rs = np.random.RandomState(2)
    d = rs.randint(0, 9, size=(4, 33, 300)).astype(np.float32)
    outi, out = ops.select_top_k(64, cu(d))
    oi, oo = cpu.select_top_k(64, d)
    assert same(outi, oi) and same(out, oo)
    outi, out = ops.select_top_k(500, cu(d))                        # k > n behaves like k = n
    oi, oo = cpu.select_top_k(500, d)
    assert same(outi, oi) and same(out, oo)


def test_selection_sort_matches_reference_cuda_kernel(refgpu):
    d = cu(np.random.RandomState(4).randint(0, 50, size=(3, 70, 512)).astype(np.float32))
    a, b = ops.select_top_k(64, d)
    ra, rb = refgpu.select_top_k(64, d)
    assert torch.equal(a, ra) and torch.equal(b, rb)


@pytest.mark.parametrize("n,m,k,c,quant", [(512, 128, 64, 3, False), (512, 40, 32, 3, True), (300, 33, 7, 5, True),
                                           (2100, 20, 100, 3, False), (64, 16, 64, 3, True), (40, 9, 1, 2, True),
                                           (8192, 64, 32, 3, False), (3000, 50, 32, 3, True), (2500, 30, 128, 3, True),
                                           (1500, 20, 128, 4, False), (129, 10, 128, 3, False), (700, 12, 33, 1, True)])
def test_knn_matches_oracle(n, m, k, c, quant):
    rs = np.random.RandomState(n + k)
    if quant:   # quantised coordinates -> many exactly equal distances -> the swap-induced tie order matters
        xyz1 = (rs.randint(0, 6, size=(2, n, c)) / 4.0).astype(np.float32)
    else:
        xyz1 = rs.random_sample((2, n, c)).astype(np.float32)
    xyz2 = xyz1[:, rs.permutation(n)[:m]].copy()
    val, idx = ops.knn_point(k, cu(xyz1), cu(xyz2))
    oval, oidx = cpu.knn_point(k, xyz1, xyz2)
    assert same(idx, oidx) and same(val, oval)


def test_knn_all_distances_equal():
    """Every candidate ties: the buffer compaction must keep the EARLIEST positions, whatever the fill level."""
    xyz1 = np.zeros((1, 1000, 3), np.float32)
    xyz1[0, 500:] = 1.0                                  # two clusters of identical points
    xyz2 = np.array([[[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5]]], np.float32)
    for k in (5, 32, 100):
        val, idx = ops.knn_point(k, cu(xyz1), cu(xyz2))
        oval, oidx = cpu.knn_point(k, xyz1, xyz2)
        assert same(idx, oidx) and same(val, oval)


# --------------------------------------------------------------------------------- three_nn / interpolate (a8-a11)
def test_interpolation_golden(golden):
    g = golden("interpolate")
    dist, idx = ops.three_nn(cu(g["xyz1"]), cu(g["xyz2"]))
    assert same(idx, g["idx"]) and same(dist, g["dist"])
    w = ops.three_weights(dist)
    assert same(w, g["weight"])
    assert same(ops.three_interpolate(cu(g["pts"]), idx, w), g["out"])
    assert same(ops.three_interpolate_grad(cu(g["pts"]), idx, w, cu(g["grad_out"])), g["grad_points"])
    d2, i2 = ops.three_nn(cu(g["xyz1"][:, :16]), cu(g["xyz2"][:, :2]))
    assert same(i2, g["idx_m2"]) and same(d2, g["dist_m2"])


@pytest.mark.parametrize("n,m,c", [(64, 16, 512), (256, 64, 256), (1024, 256, 256), (8192, 1024, 128), (1000, 37, 6),
                                   (50, 3, 9), (2048, 1500, 1)])
def test_three_nn_and_interpolate_match_oracle(n, m, c):
    xyz, _ = synth.scannet_batch(n, 2, n)
    known = cpu.gather_point(xyz, cpu.farthest_point_sample(m, xyz))
    dist, idx = ops.three_nn(cu(xyz), cu(known))
    od, oi = cpu.three_nn(xyz, known)
    assert same(idx, oi) and same(dist, od)
    ow = cpu.three_weights(od)
    w = ops.three_weights(dist)
    assert same(w, ow)
    pts = synth.features(c, 2, m, c)
    want = cpu.three_interpolate(pts, oi, ow)
    assert same(ops.three_interpolate(cu(pts), idx, w), want)
    go = synth.features(c + 7, 2, n, c)
    wantg = cpu.three_interpolate_grad(pts, oi, ow, go)
    assert same(ops.three_interpolate_grad(cu(pts), idx, w, cu(go)), wantg)
    p = cu(pts).requires_grad_(True)
    ops.three_interpolate(p, idx, w).backward(cu(go))
    assert same(p.grad, wantg)


# ------------------------------------------------------------------------------------------------- attention (a13)
@pytest.mark.parametrize("G,S,H,D", [(50, 32, 16, 4), (20, 32, 128, 4), (7, 16, 5, 4), (9, 64, 4, 8), (5, 100, 3, 16),
                                     (11, 32, 6, 2), (13, 33, 3, 1)])
def test_attention_contraction_matches_oracle(G, S, H, D):
    rs = np.random.RandomState(G + S)
    Q = rs.standard_normal((G, H * D)).astype(np.float32)
    K = rs.standard_normal((G, S, H * D)).astype(np.float32)
    V = rs.standard_normal((G, S, H * D)).astype(np.float32)
    dout = rs.standard_normal((G, H * D)).astype(np.float32)
    q, k, v = (cu(t).requires_grad_(True) for t in (Q, K, V))
    out = ops.attention_contract(q, k, v, H, D)
    np.testing.assert_allclose(npy(out), cpu.attention_fwd(Q, K, V, H, D), rtol=1e-5, atol=1e-6)
    out.backward(cu(dout))
    dQ, dK, dV = cpu.attention_bwd(Q, K, V, dout, H, D)
    np.testing.assert_allclose(npy(q.grad), dQ, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(npy(k.grad), dK, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(npy(v.grad), dV, rtol=1e-5, atol=2e-6)


def test_attention_layer_golden(golden):
    g = golden("attention")
    x = cu(g["x"])
    H, D = int(g["heads"]), int(g["key_dim"])
    layer = ops.AttentionLayer(output_dim=D, key_dim=D, num_heads=H, in_features=x.shape[-1]).to(DEV)
    with torch.no_grad():
        for lin, W, b in ((layer.query_net, "Wq", "bq"), (layer.key_net, "Wk", "bk"), (layer.value_net, "Wv", "bv")):
            lin.weight.copy_(cu(g[W]).t())                          # TF Dense stores (in,out); torch Linear (out,in)
            lin.bias.copy_(cu(g[b]))
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = layer([x, x[:, :, 0:1, :]])                           # attention_layer.py:259-261
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    np.testing.assert_allclose(npy(out), g["out"], rtol=2e-5, atol=2e-5)


# -------------------------------------------------------------------- full-size, size-independent properties
def test_full_size_properties_b16():
    B, N = 16, 8192
    xyz_np, feat_np = synth.scannet_batch(1000, B, N)
    xyz, feat = cu(xyz_np), cu(feat_np)
    fi = ops.farthest_point_sample(1024, xyz)
    assert torch.equal(fi, ops.farthest_point_sample(1024, xyz))                      # deterministic
    assert int(fi.min()) >= 0 and int(fi.max()) < N and bool((fi[:, 0] == 0).all())
    new_xyz = ops.gather_point(xyz, fi)
    assert torch.equal(new_xyz, torch.gather(xyz, 1, fi.long().unsqueeze(-1).expand(-1, -1, 3)))
    # farthest-point property: the min distance to the chosen set never increases along the sequence
    d_first = (new_xyz[:, 1:] - new_xyz[:, :1]).pow(2).sum(-1)
    assert bool((d_first[:, 0] >= d_first.max(dim=1).values - 1e-4).all())
    idx, cnt = ops.query_ball_point(0.1, 32, xyz, new_xyz)
    g = ops.group_point(xyz, idx)
    assert torch.equal(g, torch.gather(xyz.unsqueeze(1).expand(-1, 1024, -1, -1), 2,
                                       idx.long().unsqueeze(-1).expand(-1, -1, -1, 3)))
    dist = (g - new_xyz.unsqueeze(2)).pow(2).sum(-1).sqrt()
    assert bool((dist < 0.1 + 1e-6).all())                                              # every slot inside the ball
    ar = torch.arange(32, device=DEV).view(1, 1, 32)
    valid = ar < cnt.unsqueeze(-1)
    inc = (idx[:, :, 1:] > idx[:, :, :-1]) | ~valid[:, :, 1:]
    assert bool(inc.all())                                                              # hits ascending
    pad_ok = (idx == idx[:, :, :1]) | valid
    assert bool(pad_ok.all()) and int(cnt.min()) >= 1                                   # padding repeats the first hit
    gf = ops.group_point(feat, idx)
    go = torch.randn_like(gf)
    gp = ops.group_point_grad(feat, idx, go)
    lhs, rhs = (gp.double() * feat.double()).sum(), (go.double() * gf.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-6 * max(1.0, abs(float(rhs)))                     # adjointness
    d3, i3 = ops.three_nn(xyz, new_xyz)
    assert bool((d3[:, :, 0] <= d3[:, :, 1]).all()) and bool((d3[:, :, 1] <= d3[:, :, 2]).all())
    assert bool((d3[:, :, 0].gather(1, fi.long()) == 0).all())                          # a centroid is its own 1-NN
    w = ops.three_weights(d3)
    torch.testing.assert_close(w.sum(-1), torch.ones_like(w[..., 0]), rtol=0, atol=1e-6)
    const = torch.full((B, 1024, 128), 2.5, device=DEV)
    torch.testing.assert_close(ops.three_interpolate(const, i3, w), torch.full((B, N, 128), 2.5, device=DEV),
                               rtol=0, atol=1e-5)
    # and one full-size scene against the oracle
    assert same(fi[3], cpu.farthest_point_sample(1024, xyz_np[3:4])[0])
    oi, oc = cpu.query_ball_point(0.1, 32, xyz_np[3:4], npy(new_xyz[3:4]))
    assert same(idx[3], oi[0]) and same(cnt[3], oc[0])


def test_sample_and_group_and_fp_front_end():
    xyz_np, feat_np = synth.scannet_batch(50, 2, 2048)
    xyz, feat = cu(xyz_np), cu(feat_np)
    new_xyz, new_points, idx, grouped_xyz = ops.sample_and_group(256, 0.2, 32, xyz, feat)
    fi = cpu.farthest_point_sample(256, xyz_np)
    nx = cpu.gather_point(xyz_np, fi)
    oi, _ = cpu.query_ball_point(0.2, 32, xyz_np, nx)
    gx = cpu.group_point(xyz_np, oi) - nx[:, :, None, :]
    assert same(new_xyz, nx) and same(idx, oi) and same(grouped_xyz, gx)
    assert same(new_points, np.concatenate([gx, cpu.group_point(feat_np, oi)], -1))     # [xyz_local, feats] order
    p2 = synth.features(1, 2, 256, 32)
    out = ops.fp_interpolate(xyz, new_xyz, feat, cu(p2))
    od, o3 = cpu.three_nn(xyz_np, nx)
    want = np.concatenate([cpu.three_interpolate(p2, o3, cpu.three_weights(od)), feat_np], 2)
    assert same(out, want)
    _, npk, idxk, _ = ops.sample_and_group(256, 0.2, 16, xyz, None, knn=True)
    assert same(idxk, cpu.knn_point(16, xyz_np, nx)[1])


# -------------------------------------------------------------------- cell-grid paths: adversarial geometry
def _check_ball_and_nn(xyz1, xyz2, r, ns):
    idx, cnt = ops.query_ball_point(r, ns, cu(xyz1), cu(xyz2))
    oi, oc = cpu.query_ball_point(r, ns, xyz1, xyz2)
    assert same(idx, oi) and same(cnt, oc)
    d, i3 = ops.three_nn(cu(xyz2), cu(xyz1))      # xyz2 as the dense cloud, xyz1 as the known one
    od, o3 = cpu.three_nn(xyz2, xyz1)
    assert same(d, od) and same(i3, o3)
    d, i3 = ops.three_nn(cu(xyz1), cu(xyz2))      # and the other way round (queries may leave the known box)
    od, o3 = cpu.three_nn(xyz1, xyz2)
    assert same(d, od) and same(i3, o3)


def test_grid_paths_on_adversarial_clouds():
    rng = np.random.default_rng(5)
    # (a) heavy duplicates + exact ties: points on a coarse lattice, queries on lattice points
    lat = rng.integers(0, 6, size=(2, 700, 3)).astype(np.float32) * 0.125
    _check_ball_and_nn(lat, lat[:, :90].copy(), 0.25, 16)
    # (b) planar cloud (zero extent along z) and a line (zero extent along two axes)
    plane = rng.random((2, 600, 3)).astype(np.float32)
    plane[..., 2] = 0.5
    _check_ball_and_nn(plane, rng.random((2, 70, 3)).astype(np.float32), 0.2, 32)
    line = np.zeros((1, 300, 3), np.float32)
    line[..., 0] = rng.random((1, 300)).astype(np.float32)
    _check_ball_and_nn(line, line[:, ::4].copy(), 0.05, 8)
    # (c) two far-apart clusters: rings run out for queries between them -> whole-cloud scan
    a = rng.normal(0, 0.01, size=(1, 400, 3)).astype(np.float32)
    b = rng.normal(0, 0.01, size=(1, 400, 3)).astype(np.float32) + 5.0
    both = np.concatenate([a, b], 1)
    q = np.concatenate([rng.random((1, 64, 3)).astype(np.float32) * 5.0, both[:, ::10]], 1)
    _check_ball_and_nn(both, q, 0.02, 32)
    # (d) radius larger than the scene (one cell), radius smaller than any spacing (empty balls), n % 32 != 0
    u = rng.random((2, 333, 3)).astype(np.float32)
    _check_ball_and_nn(u, u[:, :50].copy(), 10.0, 32)
    _check_ball_and_nn(u, rng.random((2, 50, 3)).astype(np.float32), 1e-4, 4)
    # (e) all points identical
    same_pt = np.full((1, 200, 3), 0.3, np.float32)
    _check_ball_and_nn(same_pt, same_pt[:, :70].copy(), 0.1, 32)
    # (f) ScanNet-shaped chunk, every SA radius
    xyz, _ = synth.scannet_batch(77, 2, 4096)
    fi = cpu.farthest_point_sample(512, xyz)
    nx = cpu.gather_point(xyz, fi)
    for r in (0.1, 0.2, 0.4, 0.8):
        _check_ball_and_nn(xyz, nx, r, 32)


@pytest.mark.timeout(300)
def test_non_finite_and_huge_coordinates_do_not_hang_and_match_the_oracle():
    """A +/-inf, NaN or 1e30 coordinate used to make grid_build_kernel's coarsening loop spin for ever (the extent was
    inf, so no cell size ever fitted); the reference and the all-pairs kernels simply never hit such points."""
    rng = np.random.default_rng(9)
    for bad in (np.inf, -np.inf, 1e30, -3e38, np.nan):
        u = rng.random((2, 400, 3)).astype(np.float32)
        q = rng.random((2, 64, 3)).astype(np.float32)
        u[0, 7, 1] = bad
        u[1, 399, 0] = bad
        u[1, 100] = bad
        q[0, 3, 2] = bad
        with np.errstate(all="ignore"):
            _check_ball_and_nn(u, q, 0.15, 16)
    # a scene without a single finite coordinate
    u = np.full((1, 96, 3), np.inf, np.float32)
    with np.errstate(all="ignore"):
        _check_ball_and_nn(u, rng.random((1, 40, 3)).astype(np.float32), 0.2, 8)


# -------------------------------------------------------------------- fused layer front ends (csrc/fused.cu)
@pytest.mark.parametrize("c", [0, 6, 64, 5])
def test_fused_sa_group_equals_composition_and_oracle(c):
    from pcops_b200 import pointnet_util
    xyz_np, _ = synth.scannet_batch(21 + c, 2, 2048)
    feat_np = synth.features(c, 2, 2048, c) if c else None
    xyz = cu(xyz_np)
    feat = cu(feat_np).requires_grad_(True) if c else None
    pointnet_util.FUSED = True
    new_xyz, new_points, idx, gxyz = ops.sample_and_group(256, 0.2, 32, xyz, feat)
    pointnet_util.FUSED = False
    try:
        feat2 = cu(feat_np).requires_grad_(True) if c else None
        r_xyz, r_points, r_idx, r_gxyz = ops.sample_and_group(256, 0.2, 32, xyz, feat2)
    finally:
        pointnet_util.FUSED = True
    assert torch.equal(new_xyz, r_xyz) and torch.equal(idx, r_idx)
    assert torch.equal(new_points, r_points) and torch.equal(gxyz, r_gxyz)
    oi = npy(idx)
    want = cpu.group_point(xyz_np, oi) - npy(new_xyz)[:, :, None, :]
    if c:
        want = np.concatenate([want, cpu.group_point(feat_np, oi)], -1)
    assert same(new_points, want)
    if c:  # gradient to the features only, identical to GroupPointGrad on the feature slice
        go = torch.randn_like(new_points)
        new_points.backward(go)
        r_points.backward(go)
        assert torch.equal(feat.grad, feat2.grad)
        assert same(feat.grad, cpu.group_point_grad(feat_np, oi, npy(go)[..., 3:]))


@pytest.mark.parametrize("c2,c1", [(32, 6), (128, 0), (7, 3), (256, 128)])
def test_fused_fp_interpolate_equals_composition_and_oracle(c2, c1):
    from pcops_b200 import pointnet_util
    xyz_np, _ = synth.scannet_batch(31 + c2, 2, 2048)
    nx = cpu.gather_point(xyz_np, cpu.farthest_point_sample(256, xyz_np))
    p2_np = synth.features(c2, 2, 256, c2)
    p1_np = synth.features(c1 + 1, 2, 2048, c1) if c1 else None
    xyz, new_xyz = cu(xyz_np), cu(nx)
    p2 = cu(p2_np).requires_grad_(True)
    p1 = cu(p1_np).requires_grad_(True) if c1 else None
    out = ops.fp_interpolate(xyz, new_xyz, p1, p2)
    pointnet_util.FUSED = False
    try:
        q2 = cu(p2_np).requires_grad_(True)
        q1 = cu(p1_np).requires_grad_(True) if c1 else None
        ref_out = ops.fp_interpolate(xyz, new_xyz, q1, q2)
    finally:
        pointnet_util.FUSED = True
    assert torch.equal(out, ref_out)
    od, o3 = cpu.three_nn(xyz_np, nx)
    w = cpu.three_weights(od)
    want = cpu.three_interpolate(p2_np, o3, w)
    if c1:
        want = np.concatenate([want, p1_np], 2)
    assert same(out, want)
    go = torch.randn_like(out)
    out.backward(go)
    ref_out.backward(go)
    assert torch.equal(p2.grad, q2.grad)
    assert same(p2.grad, cpu.three_interpolate_grad(p2_np, o3, w, npy(go)[..., :c2]))
    if c1:
        assert torch.equal(p1.grad, q1.grad)


# -------------------------------------------------------------------- fused attention layer on tcgen05
@pytest.mark.parametrize("G,C", [(4, 64), (7, 64), (256, 64), (1000, 64),
                                 (4, 128), (7, 128), (600, 128), (5, 256), (300, 256), (3, 512), (150, 512)])
def test_attention_layer_fused_matches_oracle(G, C):
    """C = 64: attention_layer.cu (W resident); C = 128 / 256 / 512: attention_layer_wide.cu (both operands streamed).
    G not a multiple of 4 exercises the partial last tile; more items than SMs the persistent loop and stage reuse."""
    from pcops_b200.attention_layer import attention_layer_fused
    rng = np.random.default_rng(G + C)
    S = 32
    x = rng.standard_normal((G, S, C), dtype=np.float32)
    xq = x[:, 0, :].copy()
    W = [(rng.standard_normal((C, C), dtype=np.float32) / np.float32(np.sqrt(C))) for _ in range(3)]
    bias = [rng.standard_normal(C, dtype=np.float32) * 0.1 for _ in range(3)]
    want = cpu.attention_layer(x, xq, W[0], bias[0], W[1], bias[1], W[2], bias[2], C // 4, 4)
    got = attention_layer_fused(cu(xq), cu(x), cu(W[0]), cu(bias[0]), cu(W[1]), cu(bias[1]), cu(W[2]), cu(bias[2]))
    if C == 64:
        np.testing.assert_allclose(npy(got), want, rtol=1e-5, atol=2e-6)
    else:
        # tolerance: 1e-5 relative to the output scale (max-norm), against the float64-accumulated oracle.  The 3xTF32
        # products carry 22 mantissa bits and the tensor core accumulates K = C terms in fp32; measured 2e-6 (C = 128),
        # 3e-6 (256), 7e-6 (512) of max|out|.  Entries near zero are not held to an elementwise relative bound.
        err = np.abs(npy(got) - want).max() / np.abs(want).max()
        assert err <= 1e-5, err
    # and the module: fused inference path == Dense + contraction path
    layer = ops.AttentionLayer(4, 4, num_heads=C // 4, in_features=C).to(DEV)
    xt = cu(x).reshape(1, G, S, C)
    with torch.no_grad():
        a = layer([xt, xt[:, :, 0:1, :]])
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        b = layer([xt.requires_grad_(True), xt[:, :, 0:1, :]])     # grad enabled -> composition
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    np.testing.assert_allclose(npy(a), npy(b), rtol=2e-5, atol=2e-5)   # fp32 cuBLAS composition vs fused


@pytest.mark.gpu
def test_concurrency_hint_changes_launch_shapes_not_results():
    """pc_set_concurrency_hint sizes the streaming kernels for a lone launch (1) or for co-residency with other launches
    (> 1): group_point, three_interpolate, the attention contraction and the cell-grid searches (build CTA size, histogram
    size) must return the same bits either way."""
    from pcops_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(5)
    B, n, m, ns = 4, 2048, 512, 32
    feat = torch.randn((B, n, 64), generator=g, device=dev)
    idx = torch.randint(0, n, (B, m, ns), generator=g, device=dev, dtype=torch.int32)
    p2 = torch.randn((B, m, 128), generator=g, device=dev)
    i3 = torch.randint(0, m, (B, n, 3), generator=g, device=dev, dtype=torch.int32)
    w = torch.rand((B, n, 3), generator=g, device=dev)
    Q = torch.randn((B * m, 64), generator=g, device=dev)
    K = torch.randn((B * m, ns, 64), generator=g, device=dev)
    V = torch.randn((B * m, ns, 64), generator=g, device=dev)
    with pytest.raises(ValueError):
        _lib.set_concurrency_hint(0)
    old = _lib.set_concurrency_hint(1)
    try:
        xyz = torch.rand((B, n, 3), generator=g, device=dev)
        new_xyz = xyz[:, :m].contiguous()

        def run():
            return (ops.group_point(feat, idx), ops.three_interpolate(p2, i3, w), ops.attention_contract(Q, K, V, 16, 4),
                    *ops.query_ball_point(0.1, ns, xyz, new_xyz), *ops.three_nn(xyz, new_xyz))
        want = run()
        for h in (2, 4, 8, 64):
            assert _lib.set_concurrency_hint(h) in (1, 2, 4, 8)
            got = run()
            for a, b in zip(got, want):
                assert torch.equal(a, b), h
    finally:
        _lib.set_concurrency_hint(old)
    assert _lib.lib().pc_get_concurrency_hint() == old
