"""Whole-scene chunker (config 4; reference complete_scene_loader.py:4-117, generate_predictions.py:19-37).

CPU part: the HOST half of the product path (fp32 threshold rounding, numpy RNG planning, chunk descriptors) against
the golden fixtures produced by the reference itself, with the kernels' contracts emulated in numpy inside this test.
GPU part: the real path through libpcops.so against the oracle and the same fixtures."""
import numpy as np
import pytest
import torch

from oracle import scene_chunks as oracle_sc
from pcops_b200 import complete_scene_loader as csl
from tests.scene_cases import scene_chunk_cases, sha


def _emulated_kernels(points, npoints=8192):
    """What pc_scene_cells / pc_scene_chunk_masksum / pc_scene_chunk_assemble compute, in numpy (test-only)."""
    boxes = csl._cell_boxes(np.min(points, axis=0), np.max(points, axis=0))
    lists, inners, base = [], [], [0]
    for b in boxes:
        hit = np.all((points >= b[0:3]) & (points <= b[3:6]), axis=1)            # float32 compares, as the kernel does
        sel = np.nonzero(hit)[0]
        lists.append(sel)
        inners.append(np.all((points[sel] >= b[6:9]) & (points[sel] <= b[9:12]), axis=1))
        base.append(base[-1] + len(sel))
    cell_list, inner = np.concatenate(lists), np.concatenate(inners)
    desc, order, fill = csl._plan_chunks(np.asarray(base), npoints)
    src, mask, orig = [], [], []
    for lb, oo, start, rest, fo in desc:
        t = np.arange(npoints)
        pos = np.where(t < rest, order[oo + np.minimum(start + t, len(order) - 1 - oo)], 0)
        fpos = order[oo + fill[fo + np.maximum(t - rest, 0) % max(1, npoints - rest)]] if rest < npoints else pos
        pos = np.where(t < rest, pos, fpos)
        m = np.where(t < rest, inner[lb + pos], False)
        if m.sum() == 0:
            continue
        src.append(cell_list[lb + pos])
        mask.append(m)
        orig.append(np.where(t < rest, cell_list[lb + pos], 0))
    return np.stack(src), np.stack(mask), np.stack(orig).astype(np.int64)


def test_fp32_thresholds_equal_the_float64_compares():
    for name, seed, p, l, c, n in scene_chunk_cases()[:2]:
        boxes = csl._cell_boxes(np.min(p, axis=0), np.max(p, axis=0))
        cells = oracle_sc.cell_bounds(p)
        assert len(cells) == len(boxes)
        for b, (curmin, curmax) in zip(boxes, cells):
            want = np.sum((p >= (curmin - 0.2)) * (p <= (curmax + 0.2)), axis=1) == 3
            assert np.array_equal(np.all((p >= b[0:3]) & (p <= b[3:6]), axis=1), want)
            want = np.sum((p >= curmin) * (p <= curmax), axis=1) == 3
            assert np.array_equal(np.all((p >= b[6:9]) & (p <= b[9:12]), axis=1), want)
    # thresholds that are not representable in fp32, hit exactly from both sides
    t = np.array([0.1, -0.3, 1.7000000001])
    up, down = csl._round_up_f32(t), csl._round_down_f32(t)
    assert (up.astype(np.float64) >= t).all() and (np.nextafter(up, np.float32(-np.inf)).astype(np.float64) < t).all()
    assert (down.astype(np.float64) <= t).all() and (np.nextafter(down, np.float32(np.inf)).astype(np.float64) > t).all()


def test_host_plan_reproduces_the_reference_rng_stream(golden):
    g = golden("scene_chunks")
    for name, seed, p, l, c, n in scene_chunk_cases():
        np.random.seed(seed)
        src, mask, orig = _emulated_kernels(p)
        assert np.array_equal(np.packbits(mask), g[name + "_masks"]), name
        assert np.array_equal(orig, g[name + "_orig"]), name
        assert sha(p[src]) == str(g[name + "_sha_points"]), name
        assert sha(c[src]) == str(g[name + "_sha_colors"]), name


def test_host_plan_rejects_a_multiple_of_npoints():
    with pytest.raises(ValueError):
        csl._plan_chunks(np.array([0, 8192]))
    with pytest.raises(ValueError):
        csl._plan_chunks(np.array([0, 0, 0]))


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_chunker_matches_reference_golden_numpy_in_numpy_out(golden):
    g = golden("scene_chunks")
    for name, seed, p, l, c, n in scene_chunk_cases():
        np.random.seed(seed)
        if l is None:
            res = dict(zip(("points", "colors", "normals", "masks", "orig"),
                           csl.get_all_subsets_with_all_points_for_scene_numpy_test(p, c, n)))
        else:
            res = dict(zip(("points", "labels", "colors", "normals", "weights", "masks", "orig"),
                           csl.get_all_subsets_with_all_points_for_scene_numpy(p, l, c, n)))
        assert res["masks"].dtype == np.bool_ and res["orig"].dtype == np.int64
        assert np.array_equal(np.packbits(res["masks"]), g[name + "_masks"]), name
        assert np.array_equal(res["orig"], g[name + "_orig"]), name
        for k, v in res.items():
            assert sha(v) == str(g[name + "_sha_" + k]), (name, k)
        flat_o, flat_m = res["orig"].reshape(-1), res["masks"].reshape(-1)
        mb = csl.map_back((flat_o + 1).astype(np.int64), flat_o, flat_m, (len(p),))
        assert sha(mb) == str(g[name + "_sha_mapback"]), name
        mp = csl.map_back(res["points"].reshape(-1, 3), flat_o, flat_m, (len(p), 3))
        assert sha(mp) == str(g[name + "_sha_mapback_points"]), name


@pytest.mark.gpu
def test_chunker_matches_oracle_on_a_full_size_scan_device_tensors():
    from oracle import synth
    p, l, c, n = synth.whole_scene(11)            # 100-200 k points
    np.random.seed(123)
    want = oracle_sc.get_all_subsets_with_all_points_for_scene_numpy(p, l, c, n)
    np.random.seed(123)
    dev = torch.device("cuda")
    got = csl.get_all_subsets_with_all_points_for_scene_numpy(*(torch.from_numpy(a).to(dev) for a in (p, l, c, n)))
    assert all(isinstance(t, torch.Tensor) and t.is_cuda for t in got)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b)
    # every point of the scan is predicted exactly once after map_back (masks select the un-padded cells)
    orig, masks = got[6].reshape(-1), got[5].reshape(-1)
    back = csl.map_back(got[0].reshape(-1, 3), orig, masks, (len(p), 3))
    assert torch.equal(back.cpu(), torch.from_numpy(p))
    # duplicates: the LAST masked occurrence wins, as in numpy fancy assignment
    vals = torch.arange(6, dtype=torch.float32, device=dev)
    o = torch.tensor([2, 0, 2, 1, 2, 0], device=dev)
    m = torch.tensor([1, 1, 1, 1, 0, 1], dtype=torch.bool, device=dev)
    want = oracle_sc.map_back(vals.cpu().numpy(), o.cpu().numpy(), m.cpu().numpy(), (4,))
    assert np.array_equal(csl.map_back(vals, o, m, (4,)).cpu().numpy(), want.astype(np.float32))
    # N-D masks, exactly the (chunks, npoints) arrays the chunker returns (numpy boolean indexing consumes mask.ndim
    # leading dimensions), numpy in -> float64 numpy out; an out-of-range index raises like numpy
    pts_np, m_np, o_np = got[0].cpu().numpy(), got[5].cpu().numpy(), got[6].cpu().numpy()
    want = oracle_sc.map_back(pts_np.reshape(-1, 3), o_np.reshape(-1), m_np.reshape(-1), (len(p), 3))
    got_nd = csl.map_back(pts_np, o_np, m_np, (len(p), 3))
    assert got_nd.dtype == np.float64 and np.array_equal(got_nd, want)
    assert torch.equal(csl.map_back(got[0], got[6], got[5], (len(p), 3)).cpu(), torch.from_numpy(p))
    with pytest.raises(IndexError):
        csl.map_back(np.ones(3, np.float32), np.array([0, 9, 1]), np.array([True, True, True]), (4,))
    with pytest.raises(IndexError):
        csl.map_back(np.ones((2, 3), np.float32), np.zeros(6, np.int64), np.ones((2, 3), bool), (4,))


@pytest.mark.gpu
def test_chunker_has_no_cpu_path_and_rejects_bad_input():
    with pytest.raises(TypeError):
        csl.chunk_scene(np.zeros((10, 3), np.float64))
    rng = np.random.Generator(np.random.PCG64(3))
    p = (rng.random((8192, 3)) * 1.4).astype(np.float32)
    with pytest.raises(ValueError):                       # the reference raises here too (:89-90)
        csl.chunk_scene(p)


@pytest.mark.gpu
def test_pipelined_chunker_equals_scan_after_scan():
    """chunk_scenes() keeps later scans in flight but draws numpy's RNG in scan order: same chunks as chunk_scene()
    called scan after scan under the same seed, and as the oracle."""
    from oracle import synth
    dev = torch.device("cuda")
    scans = [synth.whole_scene(s)[0] for s in (3, 4, 5, 6, 7)]
    d_scans = [torch.from_numpy(p).to(dev) for p in scans]
    np.random.seed(77)
    want = [csl.chunk_scene(p) for p in d_scans]
    np.random.seed(77)
    want_cpu = [oracle_sc.get_all_subsets_with_all_points_for_scene_numpy_test(p, p.astype(np.uint8), p) for p in scans[:2]]
    for lookahead, background in ((0, False), (1, False), (2, True), (7, False), (0, True)):
        np.random.seed(77)
        got = []
        for chunks, ev in csl.chunk_scenes(iter(d_scans), lookahead=lookahead, background=background):
            torch.cuda.current_stream().wait_event(ev)
            got.append(chunks)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            for name in ("point_sets", "src_index", "masks", "orig_idx"):
                assert torch.equal(getattr(a, name), getattr(b, name)), (lookahead, name)
    for a, w in zip(got, want_cpu):
        assert np.array_equal(a.point_sets.cpu().numpy(), w[0])
    assert list(csl.chunk_scenes([])) == []
    for background in (False, True):                      # errors of the worker thread surface in the consumer
        with pytest.raises(TypeError):
            list(csl.chunk_scenes([np.zeros((10, 3), np.float32)], background=background))
    gen = csl.chunk_scenes(iter(d_scans), lookahead=1, background=True)   # closing early stops the worker
    next(gen)
    gen.close()


def test_native_rng_stream_equals_numpy_legacy():
    """pc_host_legacy_shuffle / pc_host_legacy_randint continue numpy's global legacy MT19937 stream bit for bit: the
    planner's orders and fill indices equal np.random.shuffle / np.random.choice under the same seed, and numpy
    continues from the same state afterwards (also across the 624-word regeneration and for one-point cells, where
    numpy's randint draws nothing)."""
    def numpy_plan(base, npoints=8192):
        orders, fills = [], []
        for cell in range(len(base) - 1):
            Lc = int(base[cell + 1] - base[cell])
            if Lc == 0:
                continue
            o = np.arange(Lc)
            np.random.shuffle(o)                                   # complete_scene_loader.py:17-18
            fills.append(np.random.choice(Lc, npoints - Lc % npoints, replace=True))   # :87
            orders.append(o)
        return np.concatenate(orders), np.concatenate(fills)
    rs = np.random.RandomState(7)
    for trial in range(12):
        counts = rs.randint(0, 30000, size=rs.randint(1, 20))
        counts[counts % 8192 == 0] += 1
        if trial % 4 == 0:
            counts[0] = 1
        base = np.concatenate([[0], np.cumsum(counts)])
        np.random.seed(trial)
        np.random.random(trial * 53)                               # any position inside the 624-word block
        want_o, want_f = numpy_plan(base)
        want_tail = np.random.random(4)
        np.random.seed(trial)
        np.random.random(trial * 53)
        _desc, got_o, got_f = csl._plan_chunks(base)
        assert got_o.dtype == np.int32 and got_f.dtype == np.int32
        assert np.array_equal(got_o, want_o) and np.array_equal(got_f, want_f)
        assert np.array_equal(np.random.random(4), want_tail)      # numpy continues from the same state
    # the state is written back when planning raises (a cell holding a multiple of npoints, as in the reference)
    np.random.seed(3)
    o = np.arange(8192)
    np.random.shuffle(o)
    want_tail = np.random.random(2)
    np.random.seed(3)
    with pytest.raises(ValueError):
        csl._plan_chunks(np.array([0, 8192]))
    assert np.array_equal(np.random.random(2), want_tail)
