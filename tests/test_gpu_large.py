"""Parity of the neighbour searches and FPS at the sizes of BASELINE.json configs[4] (N = 16 k ... 1 M points, npoint up
to 16 k, B up to 64) against the CPU oracle -- the regime where pc_query_ball_grid changes path (its per-query bitmap
over original indices stops fitting shared memory at n = 21 089, grid.cu) and where FPS runs on thread-block clusters.

The oracle is O(n * m) on the host, so the GPU op always runs on the FULL dataset cloud while the query set is a seeded
sample of it, and the oracle's OpenMP loop (over the batch dimension) is fed one (dataset, query-chunk) pair per
thread.  Everything is bit-exact: indices, counts, distances."""
import numpy as np
import pytest
import torch

import pcops_b200 as ops
from oracle import cpu, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


@pytest.fixture(params=["grid", "allpairs"])
def mode(request):
    from pcops_b200 import tf_grouping, tf_interpolate
    saved = tf_grouping.USE_GRID, tf_interpolate.USE_GRID
    tf_grouping.USE_GRID = tf_interpolate.USE_GRID = (request.param == "grid")
    yield request.param
    tf_grouping.USE_GRID, tf_interpolate.USE_GRID = saved


def cloud(kind, seed, b, n):
    """b clouds of n points: 'uniform' (unit cube, the reference smoke scripts' generator) or 'scan' (surface-like
    synthetic rooms with ~5 mm noise)."""
    if kind == "uniform":
        return synth.uniform_cube(seed, b, n, 3)
    return np.stack([synth.whole_scene(seed + i, n)[0] for i in range(b)], 0)


def chunked(fn, data, queries, chunk):
    """fn(data (B',n,3), queries (B',chunk,3)) over query chunks, every chunk paired with its scene's cloud, so the
    oracle's OpenMP-over-batch loops use all host threads.  Returns the outputs re-assembled to (b, m, ...)."""
    b, m = queries.shape[:2]
    assert m % chunk == 0
    per = m // chunk
    d = np.repeat(data, per, axis=0)                      # (b*per, n, 3): a view-free repeat is fine at these sizes
    q = queries.reshape(b * per, chunk, 3)
    outs = fn(d, q)
    return tuple(o.reshape((b, m) + o.shape[2:]) for o in outs)


def radius_for(n, kind, expect=24.0):
    if kind == "uniform":
        return float((expect / n * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0))
    return float(np.sqrt(expect / n * 120.0 / np.pi))      # ~120 m^2 of surface in a synthetic room


# n = 21 088 is the last size whose bitmap fits (grid path), 21 089 the first that takes the all-pairs kernel from
# inside pc_query_ball_grid
@pytest.mark.parametrize("kind,n,m,b", [("uniform", 16384, 2048, 2), ("scan", 16384, 1024, 3), ("uniform", 21088, 512, 2),
                                        ("uniform", 21089, 512, 2), ("scan", 65536, 1024, 2), ("uniform", 65536, 512, 2),
                                        ("scan", 262144, 512, 2), ("uniform", 262144, 256, 2)])
def test_ball_query_large_clouds(mode, kind, n, m, b):
    xyz = cloud(kind, 100 + n % 97, b, n)
    rng = np.random.default_rng(n + m)
    pick = np.sort(rng.choice(n, size=m, replace=False))
    q = xyz[:, pick].copy()
    q[:, ::7] += rng.normal(0, 0.01, size=q[:, ::7].shape).astype(np.float32)     # queries off the cloud as well
    r = radius_for(n, kind)
    idx, cnt = ops.query_ball_point(r, 32, cu(xyz), cu(q))
    oi, oc = chunked(lambda d, qq: cpu.query_ball_point(r, 32, d, qq, omp=True), xyz, q, min(m, 64))
    assert np.array_equal(npy(cnt), oc)
    assert np.array_equal(npy(idx), oi)
    assert 0 < oc.mean() <= 32


def test_ball_query_b64_config5_batch(mode):
    """B = 64 (config 5's batch) x 16 384 points, npoint 1024 from FPS, checked on every scene."""
    b, n, m = 64, 16384, 1024
    xyz = synth.uniform_cube(7, b, n, 3)
    x = cu(xyz)
    fi, nx = ops.farthest_point_sample_and_gather(m, x)
    r = radius_for(n, "uniform")
    idx, cnt = ops.query_ball_point(r, 32, x, nx)
    oi, oc = cpu.query_ball_point(r, 32, xyz, npy(nx), omp=True)
    assert np.array_equal(npy(idx), oi) and np.array_equal(npy(cnt), oc)
    # FPS on the 2-CTA cluster path, spot-checked on four scenes (the oracle needs 16 M distance updates per scene)
    sel = [0, 21, 42, 63]
    assert np.array_equal(npy(fi)[sel], cpu.farthest_point_sample(m, xyz[sel], omp=True))


@pytest.mark.parametrize("kind,n,m,b,k", [("uniform", 16384, 1024, 2, 32), ("scan", 65536, 512, 2, 32),
                                          ("uniform", 262144, 256, 2, 16), ("scan", 16384, 256, 2, 64)])
def test_knn_large_clouds(kind, n, m, b, k):
    xyz = cloud(kind, 300 + n % 89, b, n)
    rng = np.random.default_rng(n * 3 + m)
    pick = rng.choice(n, size=m, replace=False)
    q = xyz[:, pick].copy()
    q[:, 1::5] += rng.normal(0, 0.02, size=q[:, 1::5].shape).astype(np.float32)
    val, idx = ops.knn_point(k, cu(xyz), cu(q))
    chunk = 32 if n >= 262144 else 64               # the oracle forms the reference's (b, m, n) distance matrix
    ov, oi = chunked(lambda d, qq: cpu.knn_point(k, d, qq), xyz, q, chunk)
    assert np.array_equal(npy(idx), oi)
    assert np.array_equal(npy(val), ov)


# (dense n, known m): a large known cloud with sampled dense points, and a large dense cloud over a mid-size known one
@pytest.mark.parametrize("kind,n,m,b", [("uniform", 4096, 16384, 2), ("scan", 2048, 65536, 2), ("uniform", 1024, 262144, 2),
                                        ("scan", 65536, 4096, 2), ("uniform", 262144, 1024, 2), ("scan", 16384, 16384, 3)])
def test_three_nn_large_clouds(mode, kind, n, m, b):
    known = cloud(kind, 500 + m % 83, b, m)
    rng = np.random.default_rng(n + 7 * m)
    if n <= m:
        dense = known[:, rng.choice(m, size=n, replace=False)].copy()
        dense[:, ::3] += rng.normal(0, 0.01, size=dense[:, ::3].shape).astype(np.float32)
    else:
        dense = cloud(kind, 500 + m % 83, b, n)   # the same surfaces, n points: `known` is NOT a subset (other draws)
        dense[:, :m:4] = known[:, ::4]            # but exact coincidences (distance 0, ties) do occur
    d, i3 = ops.three_nn(cu(dense), cu(known))
    chunk = 64 if n % 64 == 0 else n
    od, o3 = chunked(lambda kn, dn: cpu.three_nn(dn, kn, omp=True), known, dense, chunk)
    assert np.array_equal(npy(i3), o3)
    assert np.array_equal(npy(d), od)


@pytest.mark.parametrize("n,m,b", [(16384, 4096, 2), (65536, 1024, 2), (131072, 512, 2), (262144, 256, 1), (262145, 40, 2),
                                   (300000, 64, 3), (1048576, 48, 3), (1100000, 24, 1)])
def test_fps_large_clouds(n, m, b):
    """Cluster path (2 / 8 / 16 CTAs, 16 x 16384-point slices) and, beyond 262 144 points, the cooperative grid whose
    CTAs exchange their winners through global memory (b = 3 at 1 M points = two launches: scenes {0, 1} and {2})."""
    xyz = np.stack([synth.whole_scene(900 + i + n % 61, n)[0] for i in range(b)], 0)
    got = ops.farthest_point_sample(m, cu(xyz))
    assert np.array_equal(npy(got), cpu.farthest_point_sample(m, xyz, omp=True))
