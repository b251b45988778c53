"""The pre-allocated two-stream geometry pipeline (what bench.py times) against the oracle, eager and as a CUDA graph."""
import numpy as np
import pytest
import torch

from oracle import cpu, synth
from pcops_b200.pipeline import KEY_DIM, ScanNetGeometry

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def check_against_oracle(pipe, xyz_np, feat_np):
    xyz = xyz_np
    for lv in pipe.levels:
        fi = cpu.farthest_point_sample(lv["m"], xyz)
        assert np.array_equal(npy(lv["fps_idx"]), fi)
        new_xyz = cpu.gather_point(xyz, fi)
        assert np.array_equal(npy(lv["new_xyz"]), new_xyz)
        idx, cnt = cpu.query_ball_point(lv["r"], lv["ns"], xyz, new_xyz)
        assert np.array_equal(npy(lv["idx"]), idx) and np.array_equal(npy(lv["cnt"]), cnt)
        if pipe.fuse_layers:   # centred grouped xyz and the [xyz_local, feats] tensor sample_and_group returns
            want = cpu.group_point(xyz, idx) - new_xyz[:, :, None, :]
            assert np.array_equal(npy(lv["gxyz"]), want)
            assert np.array_equal(npy(lv["new_points"]), np.concatenate([want, cpu.group_point(npy(lv["feat"]), idx)], -1))
        else:
            assert np.array_equal(npy(lv["gxyz"]), cpu.group_point(xyz, idx))
            assert np.array_equal(npy(lv["gfeat"]), cpu.group_point(npy(lv["feat"]), idx))
        if pipe.attention:
            want = cpu.attention_fwd(npy(lv["Q"]), npy(lv["K"]), npy(lv["V"]), lv["cout"] // KEY_DIM, KEY_DIM)
            np.testing.assert_allclose(npy(lv["att"]), want, rtol=1e-5, atol=1e-6)
        xyz = new_xyz
    for fp in pipe.fps:
        d, i3 = cpu.three_nn(npy(fp["xyz1"]), npy(fp["xyz2"]))
        assert np.array_equal(npy(fp["idx"]), i3) and np.array_equal(npy(fp["dist"]), d)
        w = cpu.three_weights(d)
        assert np.array_equal(npy(fp["w"]), w)
        assert np.array_equal(npy(fp["out"]), cpu.three_interpolate(npy(fp["points2"]), i3, w))


@pytest.mark.parametrize("grid,fuse", [(True, True), (False, False), (True, False)])
def test_pipeline_matches_oracle_eager_and_graph(grid, fuse):
    B = 2
    xyz_np, feat_np = synth.scannet_batch(300, B, 8192)
    pipe = ScanNetGeometry(B, grid=grid, fuse_gather=fuse or grid, fuse_layers=fuse)
    pipe.set_inputs(torch.from_numpy(xyz_np), torch.from_numpy(feat_np))
    pipe.forward(overlap=False)
    torch.cuda.synchronize()
    check_against_oracle(pipe, xyz_np, feat_np)
    # grid: +1 build (both clouds in one launch) per binned op (4 ball + 3 three_nn); fused: -1 gather, -1 group per SA
    # level, -1 weights per FP level
    assert pipe.launches_per_step == (31 if fuse else (39 if grid else 36))

    eager = [t.clone() for t in pipe.result_tensors()] + [fp["out"].clone() for fp in pipe.fps]
    pipe.capture(overlap=True)
    xyz2, feat2 = synth.scannet_batch(400, B, 8192)                 # new inputs through the captured graph
    pipe.set_inputs(torch.from_numpy(xyz2), torch.from_numpy(feat2))
    pipe.replay()
    torch.cuda.synchronize()
    check_against_oracle(pipe, xyz2, feat2)
    pipe.set_inputs(torch.from_numpy(xyz_np), torch.from_numpy(feat_np))
    pipe.replay()
    torch.cuda.synchronize()
    again = pipe.result_tensors() + [fp["out"] for fp in pipe.fps]
    for a, b in zip(eager, again):
        assert torch.equal(a, b)                                     # overlap + graph change nothing, bit for bit


def test_training_step_gradients_match_oracle():
    """Config 3: forward + the registered gradients (GroupPointGrad, ThreeInterpolateGrad) and the attention contraction's
    backward, eager and as one CUDA graph; the gradient ops are bit-exact against the reference's serial order."""
    B = 2
    xyz_np, feat_np = synth.scannet_batch(500, B, 8192)
    pipe = ScanNetGeometry(B)
    pipe.allocate_backward(seed=9)
    pipe.set_inputs(torch.from_numpy(xyz_np), torch.from_numpy(feat_np))

    def check():
        for li, lv in enumerate(pipe.levels):
            dq, dk, dv = cpu.attention_bwd(npy(lv["Q"]), npy(lv["K"]), npy(lv["V"]), npy(lv["d_att"]),
                                           lv["cout"] // KEY_DIM, KEY_DIM)
            for got, want in ((lv["dQ"], dq), (lv["dK"], dk), (lv["dV"], dv)):
                np.testing.assert_allclose(npy(got), want.reshape(got.shape), rtol=1e-5, atol=1e-5)
            if li > 0:
                want = cpu.group_point_grad(npy(lv["feat"]), npy(lv["idx"]), npy(lv["d_gfeat"]))
                assert np.array_equal(npy(lv["d_feat"]), want)
        for fp in pipe.fps:
            want = cpu.three_interpolate_grad(npy(fp["points2"]), npy(fp["idx"]), npy(fp["w"]), npy(fp["d_out"]))
            assert np.array_equal(npy(fp["d_points2"]), want)

    pipe.forward(overlap=False, train=True)
    torch.cuda.synchronize()
    check_against_oracle(pipe, xyz_np, feat_np)
    check()
    assert pipe.launches_per_train_step == pipe.launches_per_step + 4 + 2 * 3 + 2 * 4
    grads = [pipe.levels[1]["d_feat"].clone(), pipe.fps[3]["d_points2"].clone()]
    pipe.capture(overlap=True, train=True)
    for t in (pipe.levels[1]["d_feat"], pipe.fps[3]["d_points2"]):
        t.zero_()
    pipe.replay()
    torch.cuda.synchronize()
    check()
    assert torch.equal(grads[0], pipe.levels[1]["d_feat"]) and torch.equal(grads[1], pipe.fps[3]["d_points2"])


def test_pipeline_with_whole_attention_layers():
    """attention_layers=True: every level runs pc_attention_layer_fwd (tcgen05) on stand-in grouped activations; eager and
    graph replay agree bit for bit and match the oracle's Dense + contraction within 1e-5 of the output scale."""
    B = 1
    xyz_np, feat_np = synth.scannet_batch(600, B, 8192)
    pipe = ScanNetGeometry(B, attention_layers=True)
    pipe.set_inputs(torch.from_numpy(xyz_np), torch.from_numpy(feat_np))
    pipe.forward(overlap=False)
    torch.cuda.synchronize()
    assert pipe.launches_per_step == 39 + 8
    eager = []
    for lv in pipe.levels[1:]:          # SA2-SA4 (G = 256, 64, 16): small enough for the CPU oracle
        C = lv["cout"]
        W, b = [npy(w) for w in lv["W"]], [npy(v) for v in lv["b"]]
        want = cpu.attention_layer(npy(lv["X"]), npy(lv["XQ"]), W[0], b[0], W[1], b[1], W[2], b[2], C // KEY_DIM, KEY_DIM)
        err = np.abs(npy(lv["att"]) - want).max() / np.abs(want).max()
        assert err <= 1e-5, (C, err)
        eager.append(lv["att"].clone())
    pipe.capture(overlap=True)
    for lv in pipe.levels:
        lv["att"].zero_()
    pipe.replay()
    torch.cuda.synchronize()
    for lv, a in zip(pipe.levels[1:], eager):
        assert torch.equal(lv["att"], a)


def test_packed_host_batch_and_narrowed_results_are_lossless():
    """The end-to-end path of bench.py: one packed pinned input arena (xyz | normals | colours as uint8, 27 bytes per
    point; the / 255 of train.py:95 runs on the device) and the integer results narrowed to uint16 before the
    device-to-host copy -- same features bit for bit, same indices, eager and as ONE captured graph per step."""
    B = 2
    xyz_np, feat_np = synth.scannet_batch(700, B, 8192)
    col, nrm = synth.split_features(feat_np)
    pipe = ScanNetGeometry(B, own_streams=True)
    host_in = ScanNetGeometry.pack_host_batch(xyz_np, col, nrm)
    assert host_in.is_pinned() and host_in.numel() == pipe.packed_input_bytes() == B * 8192 * 27
    pipe.set_inputs_packed(host_in)
    pipe.forward(overlap=True)
    host16 = torch.empty(pipe.result_arena().numel(), dtype=torch.int16).pin_memory()
    pipe.read_results_u16(host16)
    torch.cuda.synchronize()
    assert np.array_equal(npy(pipe.xyz0), xyz_np) and np.array_equal(npy(pipe.feat0), feat_np)
    check_against_oracle(pipe, xyz_np, feat_np)
    want = npy(pipe.result_arena())
    assert want.min() >= 0 and want.max() < 65536
    assert np.array_equal(host16.numpy().view(np.uint16).astype(np.int32), want)
    assert pipe.result_bytes_u16() * 2 == pipe.result_arena().numel() * 4
    # the whole host-to-host step as one graph, on a different batch
    xyz2, feat2 = synth.scannet_batch(800, B, 8192)
    col2, nrm2 = synth.split_features(feat2)
    host_in2 = ScanNetGeometry.pack_host_batch(xyz2, col2, nrm2)
    g = pipe.capture_e2e(host_in2, host16)
    host16.zero_()
    with torch.cuda.stream(pipe.stream()):
        g.replay()
    torch.cuda.synchronize()
    check_against_oracle(pipe, xyz2, feat2)
    assert np.array_equal(host16.numpy().view(np.uint16).astype(np.int32), npy(pipe.result_arena()))
