"""Slow, independent pure-Python / numpy restatements used to cross-check the C oracle on tiny inputs."""
import numpy as np

f32 = np.float32


def fps_kernel_emulation(xyz, m, block=512):
    """Line-by-line emulation of farthestpointsamplingKernel (tf_sampling_g.cu:105-170) for ONE scene:
    512 strided lanes with strict '>' and the left-biased shared-memory tree."""
    n = xyz.shape[0]
    temp = np.full(n, f32(1e38), f32)
    out = np.zeros(m, np.int32)
    old = 0
    for j in range(1, m):
        x1, y1, z1 = xyz[old]
        dists = np.full(block, f32(-1), f32)
        dists_i = np.zeros(block, np.int64)
        for t in range(block):
            best, besti = f32(-1), 0
            for k in range(t, n, block):
                dx, dy, dz = f32(xyz[k, 0] - x1), f32(xyz[k, 1] - y1), f32(xyz[k, 2] - z1)
                d = f32(f32(f32(dx * dx) + f32(dy * dy)) + f32(dz * dz))
                d2 = min(d, temp[k])
                if d2 != temp[k]:
                    temp[k] = d2
                if d2 > best:
                    best, besti = d2, k
            dists[t], dists_i[t] = best, besti
        u = 0
        while (1 << u) < block:
            for t in range(block >> (u + 1)):
                i1, i2 = (t * 2) << u, (t * 2 + 1) << u
                if dists[i1] < dists[i2]:
                    dists[i1], dists_i[i1] = dists[i2], dists_i[i2]
            u += 1
        old = int(dists_i[0])
        out[j] = old
    return out


def fps_numpy(xyz, m):
    """Vectorised FPS with the documented tie rule: max value, then smallest (k mod 512, k)."""
    n = xyz.shape[0]
    td = np.full(n, f32(1e38), f32)
    out = np.zeros(m, np.int32)
    k = np.arange(n)
    tie = (k % 512) * (n + 1) + k // 512
    old = 0
    for j in range(1, m):
        d = xyz - xyz[old]
        d = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        td = np.minimum(d.astype(f32), td)
        cand = np.flatnonzero(td == td.max())
        old = int(cand[np.argmin(tie[cand])])
        out[j] = old
    return out


def ball_query_python(radius, nsample, xyz1, xyz2):
    """tf_grouping_g.cu:3-36 for one scene, rows pre-zeroed."""
    n, m = xyz1.shape[0], xyz2.shape[0]
    idx = np.zeros((m, nsample), np.int32)
    cnt = np.zeros(m, np.int32)
    r = f32(radius)
    for j in range(m):
        c = 0
        for k in range(n):
            if c == nsample:
                break
            dx, dy, dz = (f32(xyz2[j, i] - xyz1[k, i]) for i in range(3))
            d = max(np.sqrt(f32(f32(f32(dx * dx) + f32(dy * dy)) + f32(dz * dz))), f32(1e-20))
            if d < r:
                if c == 0:
                    idx[j, :] = k
                idx[j, c] = k
                c += 1
        cnt[j] = c
    return idx, cnt


def attention_numpy(x, Wq, bq, Wk, bk, Wv, bv, heads, kd):
    """AttentionLayer.call (attention_layer.py:29-45) in float64 with numpy reshape semantics.
    x (B,NP,S,C); query = x[:, :, 0:1, :] (:259)."""
    x = x.astype(np.float64)
    Q = np.expand_dims(x[:, :, 0:1, :] @ Wq.astype(np.float64) + bq, 2)
    K = x @ Wk.astype(np.float64) + bk
    V = x @ Wv.astype(np.float64) + bv
    Q, K, V = [np.reshape(t, (t.shape[0], t.shape[1], heads, t.shape[2], kd)) for t in (Q, K, V)]
    w = np.matmul(Q, np.swapaxes(K, -1, -2)) / np.sqrt(float(kd))
    w = np.exp(w - w.max(-1, keepdims=True))
    w = w / w.sum(-1, keepdims=True)
    out = np.matmul(w, V)
    return np.reshape(out, (out.shape[0], out.shape[1], heads * kd))
