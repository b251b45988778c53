"""Generate tests/golden/*.npz: inputs and outputs of the REFERENCE's own code.

Run in the build container (where /root/reference is mounted and `make -C oracle` has produced oracle/_ref):

    python scripts/gen_golden.py

Sources of truth
  grouping.npz     query_ball_point_cpu / group_point_cpu / group_point_grad_cpu compiled from
                   pointnet2_tensorflow/tf_ops/grouping/test/query_ball_point.cpp:19-84
  selsort.npz      selection_sort_cpu from grouping/test/selection_sort.cpp:20-63 on that program's own
                   known-answer input (dist[i] = 10-i, b=2,n=4,m=2,k=3; :65-94) plus two seeded tie-heavy cases
  interpolate.npz  threenn_cpu / threeinterpolate_cpu / threeinterpolate_grad_cpu compiled from
                   interpolation_3d/tf_interpolate.cpp:60-153 (TF headers stubbed, oracle/tf_stub)
  attention.npz    a literal numpy transcription of AttentionLayer.call (attention_layer.py:29-45: Dense, reshape,
                   matmul, softmax, matmul) in float64 -- TensorFlow itself is not installed, so this one is a
                   restatement, not reference output ("parity unpinned" for attention)
  scene_chunks.npz the reference's own complete_scene_loader.py (pure numpy, imported from /root/reference) and
                   generate_predictions.map_back (function body executed from its source, the module imports TF) on seeded
                   synthetic scans (synth.whole_scene) under np.random.seed: masks, original indices and sha256 digests of
                   every returned array
Shapes follow the reference smoke scripts (np.random.seed(100); tf_grouping.py:79-83, tf_interpolate.py:39-42).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def grouping():
    rs = np.random.RandomState(100)
    pts = rs.random_sample((4, 512, 16)).astype(np.float32)
    xyz1 = rs.random_sample((4, 512, 3)).astype(np.float32)
    xyz2 = rs.random_sample((4, 128, 3)).astype(np.float32)
    # duplicates + a far-away query (empty ball) + an exactly-on-the-sphere candidate
    xyz1[:, 100] = xyz1[:, 7]
    xyz2[:, 5] = xyz1[:, 7]
    xyz2[:, 6] = 9.0
    out = {}
    for r, ns in ((0.1, 64), (0.2, 8), (0.4, 32)):
        idx = ref.cpu_query_ball_point(r, ns, xyz1, xyz2)
        out["idx_r%g_ns%d" % (r, ns)] = idx
    idx = out["idx_r0.2_ns8"]
    out["group"] = ref.cpu_group_point(pts, idx)
    go = rs.standard_normal(out["group"].shape).astype(np.float32)
    out["group_grad"] = ref.cpu_group_point_grad(pts, idx, go)
    np.savez_compressed(os.path.join(OUT, "grouping.npz"), pts=pts, xyz1=xyz1, xyz2=xyz2, grad_out=go, **out)


def selsort():
    d0 = (10 - np.arange(16, dtype=np.float32)).reshape(2, 2, 4)
    i0, v0 = ref.cpu_selection_sort(3, d0)
    rs = np.random.RandomState(7)
    d1 = rs.randint(0, 4, size=(3, 5, 24)).astype(np.float32)  # many exact ties -> swap-order matters
    i1, v1 = ref.cpu_selection_sort(9, d1)
    d2 = np.array([2, 1, 2, 2, 1, 0], np.float32).reshape(1, 1, 6)  # stable sort would differ
    i2, v2 = ref.cpu_selection_sort(4, d2)
    np.savez_compressed(os.path.join(OUT, "selsort.npz"), d0=d0, i0=i0, v0=v0, d1=d1, i1=i1, v1=v1, d2=d2, i2=i2,
                        v2=v2)


def interpolate():
    rs = np.random.RandomState(100)
    pts = rs.random_sample((4, 128, 24)).astype(np.float32)
    xyz1 = rs.random_sample((4, 512, 3)).astype(np.float32)
    xyz2 = rs.random_sample((4, 128, 3)).astype(np.float32)
    xyz2[:, 9] = xyz2[:, 3]        # duplicate known points -> equal distances
    xyz1[:, 11] = xyz2[:, 3]       # zero distance
    dist, idx = ref.cpu_three_nn(xyz1, xyz2)
    d = np.maximum(dist, np.float32(1e-10))
    w = ((np.float32(1.0) / d) / np.sum(np.float32(1.0) / d, axis=2, keepdims=True)).astype(np.float32)
    out = ref.cpu_three_interpolate(pts, idx, w)
    go = rs.standard_normal(out.shape).astype(np.float32)
    gp = ref.cpu_three_interpolate_grad(pts, idx, w, go)
    # m < 3 known points: unfilled slots stay (1e40 -> inf, 0)
    d2, i2 = ref.cpu_three_nn(xyz1[:, :16], xyz2[:, :2])
    np.savez_compressed(os.path.join(OUT, "interpolate.npz"), pts=pts, xyz1=xyz1, xyz2=xyz2, dist=dist, idx=idx,
                        weight=w, out=out, grad_out=go, grad_points=gp, dist_m2=d2, idx_m2=i2)


def attention():
    rs = np.random.RandomState(100)
    B, NP, S, C = 2, 5, 32, 16
    heads, kd = C // 4, 4
    r32 = lambda a: a.astype(np.float32).astype(np.float64)  # inputs are exactly the float32 values stored below
    x = r32(rs.standard_normal((B, NP, S, C)))
    Wq, Wk, Wv = (r32(rs.standard_normal((C, C)) * 0.3) for _ in range(3))
    bq, bk, bv = (r32(rs.standard_normal((C,)) * 0.1) for _ in range(3))
    query = x[:, :, 0:1, :]                                   # attention_layer.py:259
    Q = query @ Wq + bq                                       # :31
    Q = np.expand_dims(Q, axis=2)                             # :32
    K = x @ Wk + bk                                           # :33
    V = x @ Wv + bv                                           # :34
    Q, K, V = [np.reshape(t, (t.shape[0], t.shape[1], heads, t.shape[2], kd)) for t in (Q, K, V)]  # :35
    w = np.matmul(Q, np.swapaxes(K, -1, -2)) / np.sqrt(float(kd))                                 # :37-38
    w = np.exp(w - w.max(-1, keepdims=True))
    w = w / w.sum(-1, keepdims=True)                          # :39
    out = np.matmul(w, V)                                     # :40
    out = np.reshape(out, (out.shape[0], out.shape[1], heads * kd))                               # :42
    np.savez_compressed(os.path.join(OUT, "attention.npz"), x=x.astype(np.float32), Wq=Wq.astype(np.float32),
                        Wk=Wk.astype(np.float32), Wv=Wv.astype(np.float32), bq=bq.astype(np.float32),
                        bk=bk.astype(np.float32), bv=bv.astype(np.float32), out=out.astype(np.float64),
                        heads=heads, key_dim=kd)


def scene_chunks():
    from tests.scene_cases import scene_chunk_cases, sha as _sha
    sys.path.insert(0, "/root/reference")
    from attention_points.scannet_dataset import complete_scene_loader as R
    # map_back lives in a module that imports tensorflow at the top: execute just that function's source
    import ast
    src = open("/root/reference/attention_points/benchmark/generate_predictions.py").read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "map_back"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "generate_predictions.py", "exec"), ns)
    out = {}
    for name, seed, p, l, c, n in scene_chunk_cases():
        np.random.seed(seed)
        if l is None:
            res = R.get_all_subsets_with_all_points_for_scene_numpy_test(p, c, n)
            names = ("points", "colors", "normals", "masks", "orig")
        else:
            res = R.get_all_subsets_with_all_points_for_scene_numpy(p, l, c, n)
            names = ("points", "labels", "colors", "normals", "weights", "masks", "orig")
        d = dict(zip(names, res))
        out[name + "_masks"] = np.packbits(d["masks"])
        out[name + "_orig"] = d["orig"].astype(np.int32)
        out[name + "_nchunks"] = np.int32(d["masks"].shape[0])
        for k, v in d.items():
            out[name + "_sha_" + k] = np.array(_sha(v))
        # map_back of a per-row value (the row's own original index + 1, so unwritten points stay 0) and of the coordinates
        vals = (d["orig"].reshape(-1) + 1).astype(np.int64)
        mb = ns["map_back"](vals, d["orig"].reshape(-1), d["masks"].reshape(-1), (len(p),))
        mp = ns["map_back"](d["points"].reshape(-1, 3), d["orig"].reshape(-1), d["masks"].reshape(-1), (len(p), 3))
        out[name + "_sha_mapback"] = np.array(_sha(mb))
        out[name + "_sha_mapback_points"] = np.array(_sha(mp))
    np.savez_compressed(os.path.join(OUT, "scene_chunks.npz"), **out)


if __name__ == "__main__":
    if not ref.available_cpu():
        sys.exit("oracle/_ref is not built: run `make -C oracle` where /root/reference is mounted")
    os.makedirs(OUT, exist_ok=True)
    grouping()
    selsort()
    interpolate()
    attention()
    scene_chunks()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
