"""Host mirror of attention_points/scannet_dataset/complete_scene_loader.py and of map_back
(attention_points/benchmark/generate_predictions.py:19-37) on libpcops.so's scene-chunk kernels (csrc/scene_chunks.cu).

Same function names, argument order and return tuples as the reference:

    get_all_subsets_with_all_points_for_scene_features(points, features, get_sample_weights)
        -> point_sets, feature_sets, sample_weights, masks_sets, points_orig_idxs_sets
    get_all_subsets_with_all_points_for_scene_numpy(points, labels, colors, normals)
    get_all_subsets_with_all_points_for_scene_numpy_test(points, colors, normals)
    map_back(values, original_idx, mask, res_shape)

numpy arrays in -> numpy arrays out (drop-in for the reference's generators, precompute_dataset.py:91,109); torch CUDA
tensors in -> torch CUDA tensors out (the chunks stay in HBM for the geometry pipeline).  Results equal the reference's
under the same ``np.random`` state: the shuffle order of every cell (np.random.shuffle, :17-18) and the fill-up indices
(np.random.choice, :87) are drawn on the host from numpy's global generator in the reference's order, everything that
touches per-point data -- cell membership, compaction, shuffling, chunking, fill-up, masks, original indices, feature
gathers, sample weights, the inverse scatter of predictions -- runs on the GPU.  There is no CPU fallback.

Work per scan: TWO host round trips (per-cell point counts; per-candidate-chunk mask sums -- the reference drops chunks
with no point inside the un-padded cell, :63,:99) instead of one numpy pass over the whole scan per cell.
"""
import numpy as np
import torch

from . import _lib

NPOINTS = 8192          # complete_scene_loader.py:11
_CELL, _PAD = 1.5, 0.2  # :23-24,:33-35


def _round_up_f32(t):
    """Smallest float32 >= t (t float64 array): p >= t for float32 p, compared in float64 as numpy does, iff p >= this."""
    f = t.astype(np.float32)
    low = f.astype(np.float64) < t
    f[low] = np.nextafter(f[low], np.float32(np.inf))
    return f


def _round_down_f32(t):
    """Largest float32 <= t."""
    f = t.astype(np.float32)
    high = f.astype(np.float64) > t
    f[high] = np.nextafter(f[high], np.float32(-np.inf))
    return f


def _cell_boxes(coordmin, coordmax):
    """The reference's cell grid (:23-24,:31-35,:41) as (ncells,12) float32 thresholds [padded lo, padded hi, lo, hi]."""
    nsubvolume_x = np.ceil((coordmax[0] - coordmin[0]) / _CELL).astype(np.int32)
    nsubvolume_y = np.ceil((coordmax[1] - coordmin[1]) / _CELL).astype(np.int32)
    boxes = []
    for i in range(nsubvolume_x):
        for j in range(nsubvolume_y):
            curmin = coordmin + [i * _CELL, j * _CELL, 0]
            curmax = coordmin + [(i + 1) * _CELL, (j + 1) * _CELL, coordmax[2] - coordmin[2]]
            boxes.append(np.concatenate([_round_up_f32(curmin - _PAD), _round_down_f32(curmax + _PAD),
                                         _round_up_f32(np.asarray(curmin, np.float64)),
                                         _round_down_f32(np.asarray(curmax, np.float64))]))
    return np.stack(boxes).astype(np.float32) if boxes else np.zeros((0, 12), np.float32)


class _LegacyStream:
    """numpy's global legacy RandomState, advanced in C for the duration of one scan's planning: the draws of
    np.random.shuffle (complete_scene_loader.py:17-18) and np.random.choice (:87) through pc_host_legacy_shuffle /
    pc_host_legacy_randint (csrc/host_rng.cu: MT19937 + numpy's masked-rejection draw, the same stream bit for bit,
    ~1.5-2 x faster than numpy's generic shuffle, which was the largest share of the chunker's host time).  The state
    is copied out once (np.random.get_state costs 0.1 ms), advanced in place and written back on exit -- also when
    planning raises -- so later np.random calls continue the same stream.  A global state that is not the legacy
    MT19937 (it always is for np.random.*) takes numpy's own functions."""

    def __enter__(self):
        st = np.random.get_state()
        self.native = st[0] == "MT19937"
        if self.native:
            self.st = st
            self.key = np.array(st[1], dtype=np.uint32, order="C", copy=True)
            self.pos = _lib.ctypes.c_int(int(st[2]))
            self.L = _lib.lib()
        return self

    def __exit__(self, *exc):
        if self.native:
            np.random.set_state((self.st[0], self.key, self.pos.value, self.st[3], self.st[4]))
        return False

    def shuffled_arange(self, n):
        if not self.native:
            order = np.arange(n)
            np.random.shuffle(order)
            return order.astype(np.int32)
        perm = np.empty(n, dtype=np.int32)
        _lib.check(self.L.pc_host_legacy_shuffle(self.key.ctypes.data, _lib.ctypes.addressof(self.pos), int(n),
                                                 perm.ctypes.data), "pc_host_legacy_shuffle")
        return perm

    def choice(self, high, count):
        if not self.native:
            return np.random.choice(high, count, replace=True).astype(np.int32)
        out = np.empty(count, dtype=np.int32)
        _lib.check(self.L.pc_host_legacy_randint(self.key.ctypes.data, _lib.ctypes.addressof(self.pos), int(high), int(count),
                                                 out.ctypes.data), "pc_host_legacy_randint")
        return out


def _plan_chunks(base, npoints=NPOINTS):
    """Host half of the chunker: numpy's global RNG stream in the reference's order -- per non-empty cell one
    np.random.shuffle (:17-18) and one np.random.choice (:87) -- and the candidate-chunk descriptors
    {list_base, order_off, start, rest, fill_off} the kernels consume.  base: (ncells+1) list offsets of the cells."""
    orders, fills, desc = [], [], []
    order_off = fill_off = 0
    with _LegacyStream() as rng:
        for cell in range(len(base) - 1):
            Lc = int(base[cell + 1] - base[cell])
            if Lc == 0:
                continue                                                         # :39-40
            order = rng.shuffled_arange(Lc)                                      # :17-18 (same stream as on a list)
            rest = Lc % npoints                                                  # :81
            if rest == 0:
                # the reference concatenates an empty list with a 2-D array here (:89-90): numpy's error, verbatim
                raise ValueError("all the input arrays must have same number of dimensions, but the array at index 0 "
                                 "has 1 dimension(s) and the array at index 1 has 2 dimension(s)")
            nfull = int(Lc / npoints)
            fill = rng.choice(Lc, npoints - rest)                                # :87
            for k in range(nfull):                                               # :56-79
                desc.append((base[cell], order_off, k * npoints, npoints, 0))
            desc.append((base[cell], order_off, nfull * npoints, rest, fill_off))    # :81-109
            orders.append(order)
            fills.append(fill)
            order_off += Lc
            fill_off += npoints - rest
    if not desc:
        raise ValueError("need at least one array to concatenate")               # :111 on an empty scan
    return np.asarray(desc, dtype=np.int32), np.concatenate(orders), np.concatenate(fills)


def _dev(a, device):
    """numpy / torch (any device) -> contiguous CUDA tensor, bytes unchanged."""
    if isinstance(a, torch.Tensor):
        return a.to(device).contiguous()
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a).to(device, non_blocking=False)


class SceneChunks:
    """Device-side result of chunking one scan.  ``src_index`` (C,npoints) int32 is the source point of every row
    (fill-up rows included): any per-point array can be chunked later with ``gather(feature)``."""

    def __init__(self, point_sets, src_index, masks, orig_idx, desc, npoints):
        self.point_sets, self.src_index, self.masks, self.orig_idx = point_sets, src_index, masks, orig_idx
        self.desc, self.npoints = desc, npoints
        self.nchunks = int(src_index.shape[0])

    def gather(self, feature):
        """feature (N, ...) of any dtype on the chunks' device -> (C, npoints, ...)."""
        L = _lib.lib()
        feature = feature.contiguous()
        row_bytes = feature.element_size() * (feature[0].numel() if feature.dim() > 1 else 1)
        out = torch.empty((self.nchunks, self.npoints) + tuple(feature.shape[1:]), dtype=feature.dtype,
                          device=feature.device)
        _lib.check(L.pc_gather_rows_bytes(self.nchunks * self.npoints, row_bytes, _lib.ptr(feature),
                                          _lib.ptr(self.src_index), _lib.ptr(out), _lib.stream()), "gather_rows_bytes")
        return out

    def sample_weights(self, gathered_labels=None):
        """(C,npoints) float64: label_weights[label] (1 without labels), masked on full chunks only (:66-70,:100-103).
        Labels are assumed to lie in 0..20 as ScanNet's do (the reference indexes a 21-entry table with them, :12-13,
        and would raise IndexError beyond it); here any non-zero label weighs 1."""
        L = _lib.lib()
        lab = None
        if gathered_labels is not None:
            lab = gathered_labels.to(torch.int32).contiguous()
        out = torch.empty((self.nchunks, self.npoints), dtype=torch.float64, device=self.masks.device)
        _lib.check(L.pc_scene_sample_weights(self.nchunks, self.npoints, _lib.ptr(self.desc), _lib.ptr(lab),
                                             _lib.ptr(self.masks), _lib.ptr(out), _lib.stream()), "scene_sample_weights")
        return out


def chunk_scene(points, npoints=NPOINTS, device=None):
    """Chunk one scan: points (N,3) float32 (numpy or torch) -> SceneChunks on ``device`` (default: the current CUDA
    device).  Draws from numpy's global RNG exactly as the reference does."""
    L = _lib.lib()
    if not torch.cuda.is_available():
        raise _lib.PcopsError("the scene chunker has no CPU implementation: a CUDA device is required")
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if isinstance(points, np.ndarray):
        if points.dtype != np.float32:
            raise TypeError("points must be float32, got %s" % points.dtype)
        coordmax, coordmin = np.max(points, axis=0), np.min(points, axis=0)           # :21-22
        pts = _dev(points, device)
    else:
        if points.dtype != torch.float32:
            raise TypeError("points must be float32, got %s" % points.dtype)
        pts = points.to(device).contiguous()
        bb = torch.empty(6, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.check(L.pc_scene_bbox(pts.shape[0], _lib.ptr(pts), _lib.ptr(bb), _lib.stream()), "scene_bbox")
        bb = bb.cpu().numpy()
        coordmin, coordmax = bb[:3], bb[3:]
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise ValueError("points must have shape (N, 3)")
    n = int(pts.shape[0])
    boxes = _cell_boxes(coordmin, coordmax)
    ncells = boxes.shape[0]
    with torch.cuda.device(device):
        st = _lib.stream()
        # ---- pass 1: per-cell compaction (ascending point index) -------------------------------------------
        d_boxes = torch.from_numpy(boxes).to(device)
        cell_base = torch.empty(ncells + 1, dtype=torch.int32, device=device)
        cell_list = torch.empty(max(1, 4 * n), dtype=torch.int32, device=device)   # a point lies in <= 2 x 2 padded cells
        inner = torch.empty(max(1, 4 * n), dtype=torch.uint8, device=device)
        ws = _lib.workspace(L.pc_scene_cells_workspace_bytes(n, ncells), device)
        _lib.check(L.pc_scene_cells(n, ncells, _lib.ptr(pts), _lib.ptr(d_boxes), _lib.ptr(cell_base), _lib.ptr(cell_list),
                                    _lib.ptr(inner), _lib.ptr(ws), st), "scene_cells")
        base = cell_base.cpu().numpy().astype(np.int64)                              # host round trip 1
        desc, order, fill = _plan_chunks(base, npoints)
        d_order = torch.from_numpy(order).to(device)
        d_fill = torch.from_numpy(fill).to(device)
        d_desc = torch.from_numpy(desc).to(device)
        # ---- pass 2: which candidate chunks hold a point of the un-padded cell (:63,:99) ---------------------
        masksum = torch.empty(len(desc), dtype=torch.int32, device=device)
        _lib.check(L.pc_scene_chunk_masksum(len(desc), npoints, _lib.ptr(d_desc), _lib.ptr(d_order), _lib.ptr(inner),
                                            _lib.ptr(masksum), st), "scene_chunk_masksum")
        keep = np.nonzero(masksum.cpu().numpy() > 0)[0]                              # host round trip 2
        if len(keep) == 0:
            raise ValueError("need at least one array to concatenate")
        d_desc = torch.from_numpy(np.ascontiguousarray(desc[keep])).to(device)
        C = len(keep)
        # ---- pass 3: assemble the kept chunks ----------------------------------------------------------------
        src_index = torch.empty((C, npoints), dtype=torch.int32, device=device)
        point_sets = torch.empty((C, npoints, 3), dtype=torch.float32, device=device)
        masks = torch.empty((C, npoints), dtype=torch.uint8, device=device)
        orig = torch.empty((C, npoints), dtype=torch.int64, device=device)
        _lib.check(L.pc_scene_chunk_assemble(C, npoints, _lib.ptr(d_desc), _lib.ptr(d_order), _lib.ptr(d_fill),
                                             _lib.ptr(cell_list), _lib.ptr(inner), _lib.ptr(pts), _lib.ptr(src_index),
                                             _lib.ptr(point_sets), _lib.ptr(masks), _lib.ptr(orig), st),
                   "scene_chunk_assemble")
    return SceneChunks(point_sets, src_index, masks, orig, d_desc, npoints)


class _PendingScan:
    """One scan inside chunk_scenes(): the four device passes of chunk_scene() as stages separated by the host
    decisions they need -- A: bounding box; B: cell membership (needs the box on the host: float64 cell bounds);
    C: candidate-chunk mask sums (needs the cell counts on the host: numpy's RNG draws, chunk descriptors);
    D: assembly (needs the mask sums on the host: which chunks survive).  Every stage ends with an asynchronous copy
    into pinned memory and an event, so the host only ever waits for this scan's own small kernels."""

    def __init__(self, pts, npoints, stream):
        self.pts, self.npoints, self.st, self.stage = pts, npoints, stream, 0
        self.dev = pts.device
        L, c_st = _lib.lib(), _lib.ctypes.c_void_p(stream.cuda_stream)
        with torch.cuda.stream(stream):
            self.bb = torch.empty(6, dtype=torch.float32, device=self.dev)
            _lib.check(L.pc_scene_bbox(pts.shape[0], _lib.ptr(pts), _lib.ptr(self.bb), c_st), "scene_bbox")
            self.h_bb = torch.empty(6, dtype=torch.float32).pin_memory()
            self.h_bb.copy_(self.bb, non_blocking=True)
            self.ev = torch.cuda.Event()
            self.ev.record(stream)

    def ready(self):
        return self.ev.query()

    def advance(self):
        """Run the next stage (waits for the previous one).  Stage C consumes numpy's global RNG: the caller runs it in
        scan order."""
        L, c_st = _lib.lib(), _lib.ctypes.c_void_p(self.st.cuda_stream)
        self.ev.synchronize()
        n, dev = int(self.pts.shape[0]), self.dev
        with torch.cuda.stream(self.st):
            if self.stage == 0:      # B
                bb = self.h_bb.numpy()
                self.boxes = _cell_boxes(bb[:3], bb[3:])
                ncells = self.boxes.shape[0]
                self.d_boxes = torch.from_numpy(self.boxes).to(dev)
                self.cell_base = torch.empty(ncells + 1, dtype=torch.int32, device=dev)
                self.cell_list = torch.empty(max(1, 4 * n), dtype=torch.int32, device=dev)
                self.inner = torch.empty(max(1, 4 * n), dtype=torch.uint8, device=dev)
                self.ws = _lib.workspace(L.pc_scene_cells_workspace_bytes(n, ncells), dev)
                _lib.check(L.pc_scene_cells(n, ncells, _lib.ptr(self.pts), _lib.ptr(self.d_boxes), _lib.ptr(self.cell_base),
                                            _lib.ptr(self.cell_list), _lib.ptr(self.inner), _lib.ptr(self.ws), c_st), "scene_cells")
                self.h_base = torch.empty(ncells + 1, dtype=torch.int32).pin_memory()
                self.h_base.copy_(self.cell_base, non_blocking=True)
            elif self.stage == 1:    # C
                base = self.h_base.numpy().astype(np.int64)
                self.desc, order, fill = _plan_chunks(base, self.npoints)
                self.d_order = torch.from_numpy(order).to(dev)
                self.d_fill = torch.from_numpy(fill).to(dev)
                d_desc = torch.from_numpy(self.desc).to(dev)
                self.masksum = torch.empty(len(self.desc), dtype=torch.int32, device=dev)
                _lib.check(L.pc_scene_chunk_masksum(len(self.desc), self.npoints, _lib.ptr(d_desc), _lib.ptr(self.d_order),
                                                    _lib.ptr(self.inner), _lib.ptr(self.masksum), c_st), "scene_chunk_masksum")
                self.h_mask = torch.empty(len(self.desc), dtype=torch.int32).pin_memory()
                self.h_mask.copy_(self.masksum, non_blocking=True)
                self._keep_alive = d_desc
            else:                    # D
                keep = np.nonzero(self.h_mask.numpy() > 0)[0]
                if len(keep) == 0:
                    raise ValueError("need at least one array to concatenate")
                d_desc = torch.from_numpy(np.ascontiguousarray(self.desc[keep])).to(dev)
                C, npts = len(keep), self.npoints
                src_index = torch.empty((C, npts), dtype=torch.int32, device=dev)
                point_sets = torch.empty((C, npts, 3), dtype=torch.float32, device=dev)
                masks = torch.empty((C, npts), dtype=torch.uint8, device=dev)
                orig = torch.empty((C, npts), dtype=torch.int64, device=dev)
                _lib.check(L.pc_scene_chunk_assemble(C, npts, _lib.ptr(d_desc), _lib.ptr(self.d_order), _lib.ptr(self.d_fill),
                                                     _lib.ptr(self.cell_list), _lib.ptr(self.inner), _lib.ptr(self.pts),
                                                     _lib.ptr(src_index), _lib.ptr(point_sets), _lib.ptr(masks), _lib.ptr(orig),
                                                     c_st), "scene_chunk_assemble")
                self.result = SceneChunks(point_sets, src_index, masks, orig, d_desc, npts)
            self.ev = torch.cuda.Event()
            self.ev.record(self.st)
            self.stage += 1
        return self.stage == 3


def chunk_scenes(scans, npoints=NPOINTS, lookahead=2, stream=None, background=False):
    """Chunk a SEQUENCE of scans (float32 CUDA tensors (N_i, 3)) with up to ``lookahead`` later scans in flight: the
    bounding-box and cell-membership passes of scans i+1, i+2 run (on ``stream``, default a stream of its own) while
    scan i is being planned on the host and consumed by the caller, so the three host round trips of a scan overlap
    with device work instead of idling the GPU.  Yields (SceneChunks, event): wait for the event on the consuming
    stream before using the tensors.  numpy's global RNG is consumed strictly in scan order (stage C runs only for
    the oldest scan), so the chunks equal those of calling chunk_scene() scan after scan under the same seed.

    background=True runs all of that on a worker thread that stays up to ``lookahead`` + 1 finished scans ahead of
    the consumer: the host planning of a scan (numpy's shuffle of every cell, a few ms, GIL released) then overlaps
    the consumer's own launches instead of alternating with them.  The worker owns numpy's global RNG and the
    ``scans`` iterator until the generator is exhausted or closed: the consumer must not draw from np.random
    meanwhile, and scans must be complete on the stream that is current when iteration starts."""
    if not background:
        yield from _chunk_scenes(scans, npoints, lookahead, stream, None)
        return
    import queue
    import threading
    q, stop = queue.Queue(maxsize=lookahead + 1), threading.Event()
    caller_stream = _CurrentStreams().get             # current streams are thread-local: capture the consumer's here

    def put(item):
        while not stop.is_set():
            try:
                q.put(item, timeout=0.05)
                return True
            except queue.Full:
                pass
        return False

    def work():
        try:
            for item in _chunk_scenes(scans, npoints, lookahead, stream, caller_stream):
                if not put(item):
                    return
            put(None)
        except BaseException as e:                    # re-raised in the consumer
            put(e)

    # Two Python threads of ~5 ms of work per scan each: with CPython's default 5 ms switch interval a thread that wants
    # the GIL can wait a whole scan for it (config 4 measured anywhere between 133 and 196 scans/s); hand over faster
    # while the worker runs.
    import sys
    old_interval = sys.getswitchinterval()
    sys.setswitchinterval(min(old_interval, 2e-4))
    t = threading.Thread(target=work, name="pcops-chunker", daemon=True)
    t.start()
    try:
        while True:
            item = q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()
        t.join()
        sys.setswitchinterval(old_interval)


class _CurrentStreams:
    """The CONSUMER thread's current stream per device, captured when chunk_scenes(background=True) starts."""

    def __init__(self):
        self._s = {torch.device("cuda", d): torch.cuda.current_stream(d) for d in range(torch.cuda.device_count())}

    def get(self, dev):
        return self._s[torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())]


def _chunk_scenes(scans, npoints, lookahead, stream, caller_stream):
    from collections import deque
    it = iter(scans)
    pend = deque()
    st = stream

    def start():
        nonlocal st
        pts = next(it, None)
        if pts is None:
            return
        if not (isinstance(pts, torch.Tensor) and pts.is_cuda and pts.dtype == torch.float32 and pts.dim() == 2 and pts.shape[1] == 3):
            raise TypeError("chunk_scenes expects float32 CUDA tensors of shape (N, 3)")
        if st is None:
            st = torch.cuda.Stream(device=pts.device)
        # the scan may just have been produced on the consumer's stream
        st.wait_stream(caller_stream(pts.device) if caller_stream else torch.cuda.current_stream(pts.device))
        with torch.cuda.device(pts.device):
            pend.append(_PendingScan(pts.contiguous(), npoints, st))

    for _ in range(lookahead + 1):
        start()
    while pend:
        for ps in list(pend)[1:]:          # younger scans: the RNG-free stage B as soon as their box has arrived
            if ps.stage == 0 and ps.ready():
                with torch.cuda.device(ps.dev):
                    ps.advance()
        head = pend[0]
        with torch.cuda.device(head.dev):
            while not head.advance():
                pass
        pend.popleft()
        start()
        yield head.result, head.ev


def get_all_subsets_with_all_points_for_scene_features(points, features, get_sample_weights):
    """complete_scene_loader.py:4-117.  ``features[0]`` are the labels when ``get_sample_weights`` (:66)."""
    as_numpy = isinstance(points, np.ndarray)
    chunks = chunk_scene(points)
    dev = chunks.src_index.device
    with torch.cuda.device(dev):
        feats = [chunks.gather(_dev(f, dev)) for f in features]
        weights = chunks.sample_weights(feats[0] if get_sample_weights else None)
        masks = chunks.masks.view(torch.bool)
        if not as_numpy:
            return chunks.point_sets, feats, weights, masks, chunks.orig_idx
        torch.cuda.synchronize(dev)
        return (chunks.point_sets.cpu().numpy(), [f.cpu().numpy() for f in feats], weights.cpu().numpy(),
                masks.cpu().numpy(), chunks.orig_idx.cpu().numpy())


def get_all_subsets_with_all_points_for_scene_numpy(points, labels, colors, normals):
    """complete_scene_loader.py:120-125"""
    point_sets, feature_sets, sample_weights, masks_sets, points_orig_idxs_sets = \
        get_all_subsets_with_all_points_for_scene_features(points, [labels, colors, normals], True)
    return point_sets, feature_sets[0], feature_sets[1], feature_sets[2], \
        sample_weights, masks_sets, points_orig_idxs_sets


def get_all_subsets_with_all_points_for_scene_numpy_test(points, colors, normals):
    """complete_scene_loader.py:128-131"""
    point_sets, feature_sets, sample_weights, masks_sets, points_orig_idxs_sets = \
        get_all_subsets_with_all_points_for_scene_features(points, [colors, normals], False)
    return point_sets, feature_sets[0], feature_sets[1], masks_sets, points_orig_idxs_sets


def map_back(values, original_idx, mask, res_shape):
    """generate_predictions.py:19-37: res = zeros(res_shape); res[original_idx[mask]] = values[mask] (the last
    occurrence of an index wins, as in numpy).  numpy in -> float64 numpy out like the reference; torch CUDA in ->
    a CUDA tensor of values' dtype.

    ``mask`` may have any number of leading dimensions (numpy boolean indexing: it consumes the first ``mask.ndim``
    dimensions of ``values`` and ``original_idx``), e.g. the (chunks, npoints) arrays the chunker returns.  An index
    outside [0, res_shape[0]) raises IndexError for numpy inputs, as numpy does; for CUDA tensors the check would cost
    a host round trip, so such rows are dropped instead (documented difference)."""
    L = _lib.lib()
    as_numpy = isinstance(values, np.ndarray)
    dev = torch.device("cuda", torch.cuda.current_device()) if as_numpy else values.device
    shape = (res_shape,) if isinstance(res_shape, (int, np.integer)) else tuple(int(x) for x in res_shape)
    nres = shape[0]
    with torch.cuda.device(dev):
        v = _dev(values, dev)
        o = _dev(original_idx, dev)
        m = _dev(mask, dev)
        k = m.dim()
        if tuple(o.shape) != tuple(m.shape):
            raise IndexError("boolean index did not match indexed array: original_idx has shape %s, mask %s"
                             % (tuple(o.shape), tuple(m.shape)))
        if tuple(v.shape[:k]) != tuple(m.shape) or tuple(v.shape[k:]) != shape[1:]:
            raise ValueError("shape mismatch: value array of shape %s could not be broadcast to indexing result"
                             % (tuple(v.shape),))
        rows = int(m.numel())
        v = v.reshape((rows,) + tuple(v.shape[k:])).contiguous()
        o = o.reshape(rows).to(torch.int64).contiguous()
        m = m.reshape(rows).to(torch.uint8).contiguous()
        if as_numpy and rows:
            sel = o[m.bool()]
            if sel.numel() and (int(sel.max()) >= nres or int(sel.min()) < -nres):
                raise IndexError("index %d is out of bounds for axis 0 with size %d"
                                 % (int(sel.max()) if int(sel.max()) >= nres else int(sel.min()), nres))
            o = torch.where(o < 0, o + nres, o)    # numpy's negative indices
        winner = torch.empty(max(1, nres), dtype=torch.int32, device=dev)
        _lib.check(L.pc_map_back_winner(rows, nres, _lib.ptr(o), _lib.ptr(m), _lib.ptr(winner), _lib.stream()),
                   "map_back_winner")
        res = torch.empty(shape, dtype=v.dtype, device=dev)
        row_bytes = v.element_size() * (v[0].numel() if v.dim() > 1 else 1)
        _lib.check(L.pc_gather_rows_bytes(nres, row_bytes, _lib.ptr(v), _lib.ptr(winner), _lib.ptr(res), _lib.stream()),
                   "gather_rows_bytes")
        if not as_numpy:
            return res
        return res.cpu().numpy().astype(np.float64)
