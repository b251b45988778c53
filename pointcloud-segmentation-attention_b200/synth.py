"""Seeded synthetic ScanNet-shaped workloads: the inputs bench.py, smoke() and the tests feed to
the CUDA path; the CPU checker re-exports this very file, so both sides consume identical bytes.

Pure numpy; no reference files are read; nothing here computes any op result.

``scannet_chunk``  one 8192-point ScanNet-shaped chunk: a 1.9 m x 1.9 m x ~3 m cell (1.5 m cell +
                   0.2 m padding, complete_scene_loader.py:33-35) made of a floor, 2-4 walls and a few
                   box / cylinder clutter surfaces with ~5 mm noise; n points drawn WITH replacement
                   from a 20-60 k pool (data_transformation.py:145) so exact duplicates occur; features
                   are colours/255 (train.py:95) and unit normals -> 6 channels.
``scannet_batch``  B chunks, seed = 20260000 + scene_id.
``uniform_cube``   np.random.random clouds like the reference smoke scripts (seed 100).
"""
import numpy as np

BASE_SEED = 20260000


def _rng(seed):
    return np.random.Generator(np.random.PCG64(int(seed)))


def _plane(rng, k, origin, u, v, normal):
    a = rng.random((k, 1))
    b = rng.random((k, 1))
    pts = origin[None, :] + a * u[None, :] + b * v[None, :]
    nrm = np.broadcast_to(normal[None, :], (k, 3))
    return pts, nrm


def scannet_chunk(seed, n=8192):
    rng = _rng(seed)
    pool = int(rng.integers(20000, 60001))
    W, H = 1.9, float(rng.uniform(2.4, 3.0))
    parts = []  # (area, sampler)
    ex, ey, ez = np.eye(3)
    parts.append((W * W, lambda k: _plane(rng, k, np.zeros(3), W * ex, W * ey, ez)))
    nwalls = int(rng.integers(2, 5))
    walls = [(np.zeros(3), W * ex, H * ez, ey), (np.zeros(3), W * ey, H * ez, ex),
             (np.array([0, W, 0.0]), W * ex, H * ez, -ey), (np.array([W, 0, 0.0]), W * ey, H * ez, -ex)]
    for w in range(nwalls):
        o, u, v, nn = walls[w]
        parts.append((W * H, (lambda o, u, v, nn: (lambda k: _plane(rng, k, o, u, v, nn)))(o, u, v, nn)))
    for _ in range(int(rng.integers(2, 6))):  # box clutter: top + two sides
        sx, sy, sz = rng.uniform(0.2, 0.7), rng.uniform(0.2, 0.7), rng.uniform(0.3, 1.0)
        ox, oy = rng.uniform(0, W - sx), rng.uniform(0, W - sy)
        o = np.array([ox, oy, 0.0])
        parts.append((sx * sy, (lambda o, sx, sy, sz: (lambda k: _plane(rng, k, o + sz * ez, sx * ex, sy * ey, ez)))(o, sx, sy, sz)))
        parts.append((sx * sz, (lambda o, sx, sz: (lambda k: _plane(rng, k, o, sx * ex, sz * ez, -ey)))(o, sx, sz)))
        parts.append((sy * sz, (lambda o, sy, sz: (lambda k: _plane(rng, k, o, sy * ey, sz * ez, -ex)))(o, sy, sz)))
    for _ in range(int(rng.integers(0, 3))):  # cylinder clutter
        r, h = rng.uniform(0.1, 0.3), rng.uniform(0.4, 1.2)
        cx, cy = rng.uniform(r, W - r), rng.uniform(r, W - r)

        def cyl(k, r=r, h=h, cx=cx, cy=cy):
            th = rng.random(k) * 2 * np.pi
            z = rng.random(k) * h
            nrm = np.stack([np.cos(th), np.sin(th), np.zeros(k)], 1)
            pts = np.stack([cx + r * np.cos(th), cy + r * np.sin(th), z], 1)
            return pts, nrm
        parts.append((2 * np.pi * r * h, cyl))
    areas = np.array([p[0] for p in parts])
    counts = rng.multinomial(pool, areas / areas.sum())
    pts, nrms, cols = [], [], []
    for (area, sampler), k in zip(parts, counts):
        if k == 0:
            continue
        p, nn = sampler(int(k))
        pts.append(p)
        nrms.append(nn)
        base = rng.integers(30, 226, size=3)
        cols.append(np.clip(base[None, :] + rng.integers(-25, 26, size=(int(k), 3)), 0, 255))
    pts = np.concatenate(pts, 0) + rng.normal(0.0, 0.005, size=(pool, 3))
    nrms = np.concatenate(nrms, 0)
    cols = np.concatenate(cols, 0).astype(np.uint8)
    pick = rng.integers(0, pool, size=n)  # with replacement -> duplicates
    xyz = pts[pick].astype(np.float32)
    feats = np.concatenate([cols[pick].astype(np.float32) / 255.0, nrms[pick].astype(np.float32)], 1)
    return xyz, feats.astype(np.float32)


def scannet_batch(first_scene, b, n=8192):
    xs, fs = zip(*(scannet_chunk(BASE_SEED + first_scene + i, n) for i in range(b)))
    return np.stack(xs, 0), np.stack(fs, 0)


def split_features(feats):
    """(…,6) float32 features of scannet_batch -> (colours uint8 (…,3), normals float32 (…,3)): the storage form of
    the reference's data set (colours are bytes; train.py:95 divides by 255 after loading).  Exact: the features were
    made as uint8 / 255 and float32(round(f * 255)) / 255 gives the same bits back (asserted)."""
    col = np.rint(feats[..., :3] * np.float32(255.0)).astype(np.uint8)
    assert np.array_equal(col.astype(np.float32) / np.float32(255.0), feats[..., :3])
    return col, np.ascontiguousarray(feats[..., 3:6])


def uniform_cube(seed, *shape):
    return np.random.RandomState(seed).random_sample(shape).astype(np.float32)


def features(seed, *shape):
    return _rng(seed).standard_normal(shape, dtype=np.float32)


def whole_scene(seed, npoints=None):
    """One synthetic whole scan (config 4): a room of 4-9 m x 4-9 m x 2.4-3 m -- floor, four walls, box clutter --
    sampled with ~5 mm noise at 100-200 k points (default), as complete_scene_loader.py expects it:
    points (N,3) float32, labels (N,) int32 in 0..20, colors (N,3) uint8, normals (N,3) float32."""
    rng = _rng(int(seed) + 7_000_000)
    n = int(npoints) if npoints is not None else int(rng.integers(100_000, 200_001))
    Lx, Ly, H = float(rng.uniform(4.0, 9.0)), float(rng.uniform(4.0, 9.0)), float(rng.uniform(2.4, 3.0))
    ex, ey, ez = np.eye(3)
    parts = [(Lx * Ly, np.zeros(3), Lx * ex, Ly * ey, ez, 2),
             (Lx * H, np.zeros(3), Lx * ex, H * ez, ey, 1), (Lx * H, np.array([0, Ly, 0.0]), Lx * ex, H * ez, -ey, 1),
             (Ly * H, np.zeros(3), Ly * ey, H * ez, ex, 1), (Ly * H, np.array([Lx, 0, 0.0]), Ly * ey, H * ez, -ex, 1)]
    for _ in range(int(rng.integers(6, 20))):
        sx, sy, sz = rng.uniform(0.3, 1.5), rng.uniform(0.3, 1.5), rng.uniform(0.3, 1.2)
        o = np.array([rng.uniform(0, Lx - sx), rng.uniform(0, Ly - sy), 0.0])
        lab = int(rng.integers(3, 21))
        parts += [(sx * sy, o + sz * ez, sx * ex, sy * ey, ez, lab), (sx * sz, o, sx * ex, sz * ez, -ey, lab),
                  (sy * sz, o, sy * ey, sz * ez, -ex, lab)]
    areas = np.array([p[0] for p in parts])
    counts = rng.multinomial(n, areas / areas.sum())
    pts, nrm, col, lab = [], [], [], []
    for (area, o, u, v, nn, lb), k in zip(parts, counts):
        k = int(k)
        if k == 0:
            continue
        p, q = _plane(rng, k, o, u, v, nn)
        pts.append(p)
        nrm.append(q)
        base = rng.integers(30, 226, size=3)
        col.append(np.clip(base[None, :] + rng.integers(-25, 26, size=(k, 3)), 0, 255))
        lab.append(np.full(k, lb if rng.random() > 0.05 else 0, np.int32))
    perm = rng.permutation(n)   # scan order is not surface order
    pts = (np.concatenate(pts, 0) + rng.normal(0.0, 0.005, size=(n, 3)))[perm].astype(np.float32)
    return pts, np.concatenate(lab)[perm], np.concatenate(col, 0).astype(np.uint8)[perm], \
        np.concatenate(nrm, 0).astype(np.float32)[perm]
