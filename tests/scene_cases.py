"""Seeded whole-scan inputs of the scene-chunker fixtures (tests/golden/scene_chunks.npz); shared by scripts/gen_golden.py
(which runs the reference on them) and the tests (which rebuild the same inputs)."""
import hashlib

import numpy as np

from oracle import synth


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


def scene_chunk_cases():
    """(name, np.random seed, points, labels or None, colors, normals)"""
    cases = []
    p, l, c, n = synth.whole_scene(1, 60000)
    cases.append(("room60k", 5, p, l, c, n))
    p, l, c, n = synth.whole_scene(2, 40000)
    cases.append(("room40k_test", 6, p, None, c, n))
    rng = np.random.Generator(np.random.PCG64(77))
    # a single cell with one full chunk + 1 point, two full chunks + 100, fewer points than one chunk
    for name, cnt, seed in (("one_cell_8193", 8193, 7), ("one_cell_16484", 16484, 8), ("one_cell_300", 300, 9)):
        p = (rng.random((cnt, 3)) * np.array([1.4, 1.4, 2.0])).astype(np.float32)
        l = rng.integers(0, 21, size=cnt).astype(np.int32)
        c = rng.integers(0, 256, size=(cnt, 3)).astype(np.uint8)
        n = rng.standard_normal((cnt, 3)).astype(np.float32)
        cases.append((name, seed, p, l, c, n))
    return cases
