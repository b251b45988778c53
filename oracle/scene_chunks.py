"""CPU restatement of the reference's whole-scene chunker and map-back -- TEST INFRASTRUCTURE ONLY.

Restates, in numpy, attention_points/scannet_dataset/complete_scene_loader.py:4-117
(get_all_subsets_with_all_points_for_scene_features) and attention_points/benchmark/generate_predictions.py:19-37
(map_back).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product
never does.

Pinned: tests/test_oracle_pins.py checks this restatement against tests/golden/scene_chunks.npz, which
scripts/gen_golden.py produced by importing and running the reference's own complete_scene_loader.py (pure numpy) on
seeded synthetic scans under the same np.random seed.

The restatement keeps the reference's arithmetic (float32 bounds promoted to float64 by the Python-list addition, :33-34,
inclusive compares :35,:41) and its use of the GLOBAL numpy RNG: one np.random.shuffle of range(L) (:17-18) and one
np.random.choice(L, npoints - rest, replace=True) (:87) per non-empty cell, in cell order, whether or not the resulting
chunks are kept.  It differs in form: index arithmetic on arrays instead of Python lists of rows.
"""
import numpy as np

NPOINTS = 8192  # complete_scene_loader.py:11


def cell_bounds(points):
    """The grid of 1.5 m cells (:21-24,:31-34): list of (curmin, curmax) float64 arrays in the reference's (i, j) order."""
    coordmax = np.max(points, axis=0)
    coordmin = np.min(points, axis=0)
    nsubvolume_x = np.ceil((coordmax[0] - coordmin[0]) / 1.5).astype(np.int32)
    nsubvolume_y = np.ceil((coordmax[1] - coordmin[1]) / 1.5).astype(np.int32)
    cells = []
    for i in range(nsubvolume_x):
        for j in range(nsubvolume_y):
            curmin = coordmin + [i * 1.5, j * 1.5, 0]
            curmax = coordmin + [(i + 1) * 1.5, (j + 1) * 1.5, coordmax[2] - coordmin[2]]
            cells.append((curmin, curmax))
    return cells


def chunk_scene(points, features, get_sample_weights, npoints=NPOINTS):
    """-> point_sets (C,npoints,3), feature_sets [ (C,npoints,...) ], sample_weights (C,npoints) f64, masks (C,npoints) bool,
    points_orig_idxs (C,npoints) int64 -- the reference's return tuple (:111-116)."""
    label_weights = np.ones(21)
    label_weights[0] = 0
    src_rows, mask_rows, orig_rows, weight_rows = [], [], [], []

    def emit(src, mask, orig, masked_weight):
        if mask.sum() == 0:                                        # :63, :99
            return
        if get_sample_weights:
            w = label_weights[features[0][src]]                     # :66, :101
        else:
            w = np.ones(len(src))                                   # :68, :103
        if masked_weight:
            w = w * mask                                            # :70 (full chunks only)
        src_rows.append(src)
        mask_rows.append(mask)
        orig_rows.append(orig)
        weight_rows.append(w)

    for curmin, curmax in cell_bounds(points):
        curchoice = np.sum((points >= (curmin - 0.2)) * (points <= (curmax + 0.2)), axis=1) == 3   # :35
        sel = np.nonzero(curchoice)[0]
        L = len(sel)
        if L == 0:
            continue                                                # :39-40
        inner = np.sum((points[sel] >= curmin) * (points[sel] <= curmax), axis=1) == 3              # :41
        order = list(range(L))
        np.random.shuffle(order)                                    # :17-18
        order = np.asarray(order, dtype=np.int64)
        sel, inner = sel[order], inner[order]
        nfull = int(L / npoints)
        for k in range(nfull):                                      # :56-79
            rows = slice(k * npoints, (k + 1) * npoints)
            emit(sel[rows], inner[rows], sel[rows], True)
        rest = L % npoints                                          # :81
        if rest == 0:
            # the reference concatenates an empty Python list with a 2-D array here (:89-90) and numpy raises
            raise ValueError("all the input arrays must have same number of dimensions, but the array at index 0 has "
                             "1 dimension(s) and the array at index 1 has 2 dimension(s)")
        offset = nfull * npoints                                    # :84-86; when L == npoints the reference's offset is 0 but rest == 0, so nothing is read
        fill = np.random.choice(L, npoints - rest, replace=True)    # :87
        src = np.concatenate((sel[offset:offset + rest], sel[fill]))
        mask = np.concatenate((inner[offset:offset + rest], np.zeros(npoints - rest, dtype=bool)))  # :92
        orig = np.concatenate((sel[offset:offset + rest], np.zeros(npoints - rest, dtype=int)))     # :93-94
        emit(src, mask, orig, False)

    src = np.stack(src_rows)  # the reference raises on an empty result too (np.concatenate of nothing, :111)
    point_sets = points[src]
    feature_sets = [np.asarray(f)[src] for f in features]
    return point_sets, feature_sets, np.stack(weight_rows), np.stack(mask_rows), np.stack(orig_rows).astype(np.int64)


def get_all_subsets_with_all_points_for_scene_numpy(points, labels, colors, normals):
    """complete_scene_loader.py:120-125"""
    p, f, w, m, o = chunk_scene(points, [labels, colors, normals], True)
    return p, f[0], f[1], f[2], w, m, o


def get_all_subsets_with_all_points_for_scene_numpy_test(points, colors, normals):
    """complete_scene_loader.py:128-131"""
    p, f, w, m, o = chunk_scene(points, [colors, normals], False)
    return p, f[0], f[1], m, o


def map_back(values, original_idx, mask, res_shape):
    """generate_predictions.py:19-37"""
    values = values[mask]
    original_idx = original_idx[mask]
    res = np.zeros(res_shape)
    res[original_idx] = values
    return res
