"""The whole attention-model inference forward (model_pipeline.ScanNetAttentionModel: 4 SA-attention levels, 4 FP levels,
fc1, fc2 -- attention_points/models/pointnet2_sem_seg_attention.py:28-62) against a float64 restatement of the same graph
built from the oracle's pieces: the geometry decisions are bit-exact (they only depend on xyz), every floating-point
tensor is compared at the output scale.  Errors of ~2e-6 per tensor-core layer accumulate over the 24 dense layers of
the graph; the bounds below are the per-depth figures measured on B200 with a 4x margin."""
import numpy as np
import pytest
import torch

from oracle import cpu, synth
from pcops_b200.model_pipeline import ScanNetAttentionModel

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def mlp64(x, layers):
    for d in layers:
        x = x @ npy(d.w).astype(np.float64) + npy(d.b).astype(np.float64)
        if d.relu:
            x = np.maximum(x, 0.0)
    return x


def reference_forward(model, xyz, feats):
    """float64 numpy restatement of the graph on the model's own weights; returns per-level outputs and the logits."""
    B = xyz.shape[0]
    outs, cur_xyz, cur_feat = [], xyz, feats
    geo = []
    for lv in model.sa:
        fi = cpu.farthest_point_sample(lv["m"], cur_xyz)
        nx = cpu.gather_point(cur_xyz, fi)
        idx, _ = cpu.query_ball_point(lv["r"], lv["ns"], cur_xyz, nx)
        gx = (cpu.group_point(cur_xyz, idx) - nx[:, :, None, :]).astype(np.float64)
        if cur_feat is not None:
            b_ix = np.arange(B)[:, None, None]
            new_points = np.concatenate([gx, np.asarray(cur_feat, np.float64)[b_ix, idx]], -1)
        else:
            new_points = gx
        X = mlp64(new_points, lv["mlp"])
        m, ns, C = lv["m"], lv["ns"], lv["cout"]
        W, b = [npy(w) for w in lv["W"]], [npy(v) for v in lv["bq"]]
        x = X.reshape(B * m, ns, C)
        att = cpu.attention_layer_f64(x, x[:, 0, :], W[0], b[0], W[1], b[1], W[2], b[2], C // 4, 4).reshape(B, m, C)
        geo.append((cur_xyz, nx, fi, idx, cur_feat))
        outs.append(att)
        cur_xyz, cur_feat = nx, att
    points2 = outs[3]
    fps = []
    for k, fp in enumerate(model.fp):
        xyz1, xyz2, _, _, points1 = geo[3 - k]
        d, i3 = cpu.three_nn(xyz1, xyz2)
        w = cpu.three_weights(d).astype(np.float64)
        b_ix = np.arange(B)[:, None]
        interp = sum(points2[b_ix, i3[..., t]] * w[..., t:t + 1] for t in range(3))
        cat = interp if points1 is None else np.concatenate([interp, np.asarray(points1, np.float64)], -1)
        points2 = mlp64(cat, fp["mlp"])
        fps.append(points2)
    net = mlp64(points2, [model.fc1])
    return outs, fps, mlp64(net, [model.fc2]), geo


def rel_scale(got, want):
    return float(np.abs(npy(got).astype(np.float64) - want).max() / np.abs(want).max())


@pytest.mark.parametrize("feat_channels", [6, 0])
def test_whole_model_forward_matches_float64_graph(feat_channels):
    B = 1
    xyz, feats = synth.scannet_batch(900 + feat_channels, B, 8192)
    model = ScanNetAttentionModel(B, 8192, feat_channels, seed=3)
    model.set_inputs(torch.from_numpy(xyz), torch.from_numpy(feats) if feat_channels else None)
    model.forward()
    torch.cuda.synchronize()
    outs, fps, logits, geo = reference_forward(model, xyz, feats if feat_channels else None)
    for lv, (cx, nx, fi, idx, _) in zip(model.sa, geo):
        assert np.array_equal(npy(lv["fps_idx"]), fi) and np.array_equal(npy(lv["new_xyz"]), nx)
        assert np.array_equal(npy(lv["idx"]), idx)
    errs = [rel_scale(lv["out"], o) for lv, o in zip(model.sa, outs)]
    errs += [rel_scale(fp["h"][-1].view(B, fp["n"], -1), o) for fp, o in zip(model.fp, fps)]
    errs.append(rel_scale(model.logits, logits))
    print("max err / max|out| at SA1..4, FP1..4, logits:", " ".join("%.1e" % e for e in errs))
    assert max(errs[:4]) <= 2e-5 and max(errs) <= 1e-4
    assert model.launches_per_step == 58
    # the same forward as ONE CUDA graph on another batch, bit-identical to its eager run
    xyz2, feats2 = synth.scannet_batch(950, B, 8192)
    model.set_inputs(torch.from_numpy(xyz2), torch.from_numpy(feats2) if feat_channels else None)
    model.forward()
    torch.cuda.synchronize()
    eager = model.logits.clone()
    model.capture()
    model.logits.zero_()
    model.replay()
    torch.cuda.synchronize()
    assert torch.equal(model.logits, eager)
