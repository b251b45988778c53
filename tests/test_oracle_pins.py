"""The CPU oracle pinned against the reference: golden fixtures produced by the reference's own code
(scripts/gen_golden.py), oracle/_ref when it is built here, and independent Python restatements."""
import numpy as np
import pytest

from oracle import cpu, ref, synth
from tests import pyref

needs_ref = pytest.mark.skipif(not ref.available_cpu(), reason="oracle/_ref not built (no /root/reference)")


# ---- golden fixtures: outputs of the reference's own code -------------------------------------------------
def test_ball_query_matches_reference_golden(golden):
    g = golden("grouping")
    for r, ns in ((0.1, 64), (0.2, 8), (0.4, 32)):
        idx, cnt = cpu.query_ball_point(r, ns, g["xyz1"], g["xyz2"])
        assert np.array_equal(idx, g["idx_r%g_ns%d" % (r, ns)])
        assert cnt.min() == 0 and (idx[cnt == 0] == 0).all()        # the far-away query: empty ball, zero row
        assert cnt.max() <= ns


def test_group_point_and_grad_match_reference_golden(golden):
    g = golden("grouping")
    idx = g["idx_r0.2_ns8"]
    assert np.array_equal(cpu.group_point(g["pts"], idx), g["group"])
    assert np.array_equal(cpu.group_point_grad(g["pts"], idx, g["grad_out"]), g["group_grad"])


def test_selection_sort_matches_reference_known_answer(golden):
    g = golden("selsort")
    # the reference program's own known answer: every row 3 2 1 0 (selection_sort.cpp:65-94)
    assert np.array_equal(g["i0"].reshape(-1, 4), np.tile([3, 2, 1, 0], (4, 1)))
    for tag, k in (("0", 3), ("1", 9), ("2", 4)):
        outi, out = cpu.select_top_k(k, g["d" + tag])
        assert np.array_equal(outi, g["i" + tag]) and np.array_equal(out, g["v" + tag])
    # swap-induced instability: a stable sort would give [5,1,4,0]
    assert list(g["i2"].ravel()[:4]) == [5, 1, 4, 3]


def test_three_nn_interpolate_match_reference_golden(golden):
    g = golden("interpolate")
    dist, idx = cpu.three_nn(g["xyz1"], g["xyz2"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    assert np.array_equal(cpu.three_weights(dist), g["weight"])
    assert np.array_equal(cpu.three_interpolate(g["pts"], idx, g["weight"]), g["out"])
    assert np.array_equal(cpu.three_interpolate_grad(g["pts"], idx, g["weight"], g["grad_out"]), g["grad_points"])
    d2, i2 = cpu.three_nn(g["xyz1"][:, :16], g["xyz2"][:, :2])
    assert np.array_equal(i2, g["idx_m2"]) and np.array_equal(d2, g["dist_m2"])
    assert np.isinf(d2[..., 2]).all() and (i2[..., 2] == 0).all()   # m < 3: (inf, 0)


def test_attention_matches_numpy_transcription(golden):
    g = golden("attention")
    x = g["x"]
    B, NP, S, C = x.shape
    H, D = int(g["heads"]), int(g["key_dim"])
    out = cpu.attention_layer(x.reshape(B * NP, S, C), x[:, :, 0, :].reshape(B * NP, C), g["Wq"], g["bq"], g["Wk"],
                              g["bk"], g["Wv"], g["bv"], H, D)
    np.testing.assert_allclose(out.reshape(B, NP, C), g["out"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(pyref.attention_numpy(x, g["Wq"], g["bq"], g["Wk"], g["bk"], g["Wv"], g["bv"], H, D),
                               g["out"], rtol=1e-12, atol=1e-12)


# ---- live cross-checks against oracle/_ref (the reference sources compiled here) --------------------------
@needs_ref
@pytest.mark.parametrize("seed,n,m,r,ns", [(1, 300, 40, 0.15, 16), (2, 1024, 256, 0.2, 32), (3, 64, 16, 0.8, 32)])
def test_ball_group_against_ref_cpu(seed, n, m, r, ns):
    xyz1 = synth.uniform_cube(seed, 3, n, 3)
    xyz2 = xyz1[:, :m].copy()
    idx, _ = cpu.query_ball_point(r, ns, xyz1, xyz2)
    assert np.array_equal(idx, ref.cpu_query_ball_point(r, ns, xyz1, xyz2))
    for c in (3, 6, 9, 64):
        pts = synth.features(seed + c, 3, n, c)
        assert np.array_equal(cpu.group_point(pts, idx), ref.cpu_group_point(pts, idx))
        go = synth.features(seed + 100 + c, 3, m, ns, c)
        assert np.array_equal(cpu.group_point_grad(pts, idx, go), ref.cpu_group_point_grad(pts, idx, go))


@needs_ref
def test_scannet_chunk_against_ref_cpu():
    xyz, feats = synth.scannet_batch(0, 2, 2048)
    new_xyz = cpu.gather_point(xyz, cpu.farthest_point_sample(256, xyz))
    idx, cnt = cpu.query_ball_point(0.2, 32, xyz, new_xyz)
    assert np.array_equal(idx, ref.cpu_query_ball_point(0.2, 32, xyz, new_xyz))
    dist, i3 = cpu.three_nn(xyz, new_xyz)
    rd, ri = ref.cpu_three_nn(xyz, new_xyz)
    assert np.array_equal(i3, ri) and np.array_equal(dist, rd)
    w = cpu.three_weights(dist)
    pts = synth.features(5, 2, 256, 32)
    assert np.array_equal(cpu.three_interpolate(pts, i3, w), ref.cpu_three_interpolate(pts, i3, w))
    go = synth.features(6, 2, 2048, 32)
    assert np.array_equal(cpu.three_interpolate_grad(pts, i3, w, go), ref.cpu_three_interpolate_grad(pts, i3, w, go))


# ---- independent restatements ------------------------------------------------------------------------------
def test_fps_against_kernel_emulation_with_ties():
    rs = np.random.RandomState(3)
    # quantised coordinates + duplicated points: many exact ties, n straddles the 512-lane partition
    for n, m in ((40, 12), (513, 24), (700, 40), (1100, 16)):
        xyz = (rs.randint(0, 6, size=(n, 3)) / 4.0).astype(np.float32)
        got = cpu.farthest_point_sample(m, xyz[None])[0]
        assert np.array_equal(got, pyref.fps_kernel_emulation(xyz, m))
        assert np.array_equal(got, pyref.fps_numpy(xyz, m))


def test_fps_scannet_chunk_against_numpy():
    xyz, _ = synth.scannet_chunk(synth.BASE_SEED + 7, 8192)
    got = cpu.farthest_point_sample(128, xyz[None])[0]
    assert got[0] == 0 and np.array_equal(got, pyref.fps_numpy(xyz, 128))


def test_fps_edge_cases():
    xyz = synth.uniform_cube(9, 2, 10, 3)
    assert np.array_equal(cpu.farthest_point_sample(1, xyz), np.zeros((2, 1), np.int32))
    idx = cpu.farthest_point_sample(16, xyz)                        # m > n: keeps emitting, first n are a permutation
    assert sorted(idx[0, :10]) == list(range(10))
    same = np.ones((1, 33, 3), np.float32)                          # all duplicates -> always index 0
    assert (cpu.farthest_point_sample(8, same) == 0).all()


def test_ball_query_against_python_loops():
    xyz1 = synth.uniform_cube(4, 1, 90, 3)[0]
    xyz2 = xyz1[:20]
    for r, ns in ((0.25, 8), (0.6, 4)):
        idx, cnt = cpu.query_ball_point(r, ns, xyz1[None], xyz2[None])
        pi, pc = pyref.ball_query_python(r, ns, xyz1, xyz2)
        assert np.array_equal(idx[0], pi) and np.array_equal(cnt[0], pc)


def radius_probe(r, span=24):
    """Candidates (x, y, 0) whose fp32 squared distance to the origin takes every float within `span` ulps of
    fl(r*r); expected hits by the reference formula max(sqrtf(s),1e-20f) < r (tf_grouping_g.cu:24-25)."""
    r = np.float32(r)
    x = np.float32(r * np.float32(0.99))
    a = np.float32(x * x)
    targets = [np.float32(r * r)]
    for _ in range(span):
        targets.insert(0, np.nextafter(targets[0], np.float32(0)))
        targets.append(np.nextafter(targets[-1], np.float32(1)))
    pts, svals = [], []
    for t in targets:
        y = np.float32(np.sqrt(np.float64(t) - np.float64(a)))
        lo = hi = y
        cands = [y]
        for _ in range(200):
            lo = np.nextafter(lo, np.float32(0))
            hi = np.nextafter(hi, np.float32(1))
            cands += [lo, hi]
        for y in cands:
            if np.float32(a + np.float32(y * y)) == t:
                pts.append((x, y, np.float32(0)))
                svals.append(t)
                break
    cand = np.array(pts, np.float32)[None]
    s = np.array(svals, np.float32)
    expect = np.maximum(np.sqrt(s), np.float32(1e-20)) < r
    naive = s < np.float32(r * r)
    return cand, expect, naive


def test_ball_query_threshold_is_sqrt_not_square():
    differs = 0
    for r in (0.1, 0.2, 0.4, 0.8, 0.3, 1.7):
        cand, expect, naive = radius_probe(r)
        assert cand.shape[1] > 20
        idx, cnt = cpu.query_ball_point(r, cand.shape[1], cand, np.zeros((1, 1, 3), np.float32))
        assert cnt[0, 0] == expect.sum()
        assert np.array_equal(idx[0, 0, :cnt[0, 0]], np.flatnonzero(expect))
        differs += int((expect != naive).any())
    assert differs > 0      # d2 < r*r is NOT the same test


def test_knn_matches_selection_sort_of_distance_matrix():
    rs = np.random.RandomState(5)
    xyz1 = (rs.randint(0, 5, size=(2, 60, 3)) / 2.0).astype(np.float32)   # ties galore
    xyz2 = xyz1[:, :9].copy()
    dist = cpu.knn_dist(xyz1, xyz2)
    outi, out = cpu.select_top_k(7, dist)
    val, idx = cpu.knn_point(7, xyz1, xyz2)
    assert np.array_equal(idx, outi[:, :, :7]) and np.array_equal(val, out[:, :, :7])


def test_gradients_are_transposes():
    # tf_grouping_op_test.py:23-25 / tf_interpolate_op_test.py:19-21 check analytic-vs-numeric Jacobians < 1e-4;
    # both ops are linear in `points`, so <grad(g), p> == <g, op(p)> exactly characterises the Jacobian.
    rs = np.random.RandomState(0)
    pts = rs.random_sample((1, 128, 16)).astype(np.float32)
    xyz = rs.random_sample((1, 128, 3)).astype(np.float32)
    idx, _ = cpu.query_ball_point(0.3, 32, xyz, xyz[:, :8])
    g = rs.standard_normal((1, 8, 32, 16)).astype(np.float32)
    lhs = np.sum(cpu.group_point_grad(pts, idx, g).astype(np.float64) * pts)
    rhs = np.sum(g.astype(np.float64) * cpu.group_point(pts, idx))
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(rhs))
    p2 = rs.random_sample((1, 8, 16)).astype(np.float32)
    d, i3 = cpu.three_nn(xyz, xyz[:, :8])
    w = np.full_like(d, 1.0 / 3.0)
    g2 = rs.standard_normal((1, 128, 16)).astype(np.float32)
    lhs = np.sum(cpu.three_interpolate_grad(p2, i3, w, g2).astype(np.float64) * p2)
    rhs = np.sum(g2.astype(np.float64) * cpu.three_interpolate(p2, i3, w))
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(rhs))


def test_attention_backward_against_finite_differences():
    rs = np.random.RandomState(1)
    G, S, H, D = 3, 8, 2, 4
    Q = rs.standard_normal((G, H * D)).astype(np.float32)
    K = rs.standard_normal((G, S, H * D)).astype(np.float32)
    V = rs.standard_normal((G, S, H * D)).astype(np.float32)
    dout = rs.standard_normal((G, H * D)).astype(np.float32)
    dQ, dK, dV = cpu.attention_bwd(Q, K, V, dout, H, D)

    def loss(q, k, v):
        return float(np.sum(cpu.attention_fwd(q, k, v, H, D).astype(np.float64) * dout))
    eps = 1e-2
    for arr, grad in ((Q, dQ), (K, dK), (V, dV)):
        flat = arr.reshape(-1)
        for pos in rs.choice(flat.size, 6, replace=False):
            old = flat[pos]
            flat[pos] = old + eps
            up = loss(Q, K, V)
            flat[pos] = old - eps
            dn = loss(Q, K, V)
            flat[pos] = old
            assert abs((up - dn) / (2 * eps) - grad.reshape(-1)[pos]) < 2e-3


def test_prob_sample_oracle_against_python_emulation_of_the_kernel():
    """cumsumKernel / binarysearchKernel (tf_sampling_g.cu:7-104) emulated literally in numpy float32 -- thread loops
    unrolled into python loops, padded buffer indices kept -- against the C restatement (unpadded, regrouped)."""
    def kernel_cumsum(inp):
        f = np.float32
        n = inp.shape[0]
        out = np.zeros(n, f)
        BS, PL = 2048, 5
        rs, rs2 = f(0), f(0)
        for j in range(0, n, BS * 4):
            n24_i = min(n - j, BS * 4)
            n24 = (n24_i + 3) & ~3
            n2 = n24 >> 2
            b4 = np.zeros(BS * 4, f)
            bf = np.zeros(BS + (BS >> PL), f)
            for k in range(0, n24_i, 4):
                if k + 3 < n24_i:
                    v1, v2, v3, v4 = (f(inp[j + k + t]) for t in range(4))
                    v2 = f(v2 + v1); v4 = f(v4 + v3); v3 = f(v3 + v2); v4 = f(v4 + v2)
                    b4[k:k + 4] = (v1, v2, v3, v4)
                    bf[(k >> 2) + (k >> (2 + PL))] = v4
                else:
                    v = f(0)
                    for k2 in range(k, n24_i):
                        v = f(v + inp[j + k2]); b4[k2] = v
                    b4[n24_i:n24] = v
                    bf[(k >> 2) + (k >> (2 + PL))] = v
            u = 0
            while (2 << u) <= n2:
                for k in range(n2 >> (u + 1)):
                    i1 = (((k << 1) + 2) << u) - 1; i2 = (((k << 1) + 1) << u) - 1
                    i1 += i1 >> PL; i2 += i2 >> PL
                    bf[i1] = f(bf[i1] + bf[i2])
                u += 1
            u -= 1
            while u >= 0:
                for k in range((n2 - (1 << u)) >> (u + 1)):
                    i1 = (((k << 1) + 3) << u) - 1; i2 = (((k << 1) + 2) << u) - 1
                    i1 += i1 >> PL; i2 += i2 >> PL
                    bf[i1] = f(bf[i1] + bf[i2])
                u -= 1
            for k in range(4, n24, 4):
                k2 = ((k >> 2) - 1) + (((k >> 2) - 1) >> PL)
                b4[k:k + 4] = b4[k:k + 4] + bf[k2]
            out[j:j + n24_i] = b4[:n24_i] + rs
            t = f(bf[(n2 - 1) + ((n2 - 1) >> PL)] + rs2)
            r2 = f(rs + t)
            rs2 = f(t - f(r2 - rs))
            rs = r2
        return out

    rs = np.random.RandomState(100)
    for n in (1, 2, 4, 7, 64, 129, 1000, 8192, 8199, 17000):
        w = rs.random_sample((1, n)).astype(np.float32)
        assert np.array_equal(cpu.cumsum(w)[0], kernel_cumsum(w[0])), n
    w = rs.random_sample((2, 777)).astype(np.float32)
    w[:, ::3] = 0
    r = rs.random_sample((2, 300)).astype(np.float32)
    cdf = cpu.cumsum(w)
    got = cpu.prob_sample(w, r)
    for i in range(2):
        q = (r[i] * cdf[i, -1]).astype(np.float32)
        want = np.searchsorted(cdf[i], q, side="left")      # smallest index with cdf >= q (cdf is non-decreasing)
        assert np.array_equal(got[i], np.minimum(want, 776))


# ---- whole-scene chunker: the restatement against the reference's own complete_scene_loader.py ------------------
def test_scene_chunker_matches_reference_golden(golden):
    from oracle import scene_chunks as sc
    from tests.scene_cases import scene_chunk_cases, sha
    g = golden("scene_chunks")
    for name, seed, p, l, c, n in scene_chunk_cases():
        np.random.seed(seed)
        if l is None:
            res = dict(zip(("points", "colors", "normals", "masks", "orig"),
                           sc.get_all_subsets_with_all_points_for_scene_numpy_test(p, c, n)))
        else:
            res = dict(zip(("points", "labels", "colors", "normals", "weights", "masks", "orig"),
                           sc.get_all_subsets_with_all_points_for_scene_numpy(p, l, c, n)))
        assert res["masks"].shape[0] == int(g[name + "_nchunks"])
        assert np.array_equal(np.packbits(res["masks"]), g[name + "_masks"])
        assert np.array_equal(res["orig"], g[name + "_orig"])
        for k, v in res.items():
            assert sha(v) == str(g[name + "_sha_" + k]), (name, k)
        flat_o, flat_m = res["orig"].reshape(-1), res["masks"].reshape(-1)
        assert sha(sc.map_back((flat_o + 1).astype(np.int64), flat_o, flat_m, (len(p),))) == str(g[name + "_sha_mapback"])
        assert sha(sc.map_back(res["points"].reshape(-1, 3), flat_o, flat_m, (len(p), 3))) == str(g[name + "_sha_mapback_points"])


def test_scene_chunker_rejects_what_the_reference_rejects():
    """A cell holding an exact multiple of 8192 points makes the reference concatenate an empty list with a 2-D array
    (complete_scene_loader.py:89-90): ValueError."""
    from oracle import scene_chunks as sc
    rng = np.random.Generator(np.random.PCG64(3))
    p = (rng.random((8192, 3)) * 1.4).astype(np.float32)
    with pytest.raises(ValueError):
        sc.chunk_scene(p, [np.zeros(8192, np.int32)], True)
