"""Where config 4's ~4.4 ms per scan goes: worker side (chunk_scenes alone, nothing consumed) and consumer side (gathers,
forward over the chunks, map_back on chunks prepared beforehand), each timed alone over the same scans."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcops_b200 import _lib, synth                       # noqa: E402
from pcops_b200 import complete_scene_loader as csl      # noqa: E402
from pcops_b200.pipeline import ScanNetGeometry          # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
S, B, D = 39, 16, 8
_lib.set_concurrency_hint(D)
scans = []
k = 0
while len(scans) < S:
    p, l, c, n = synth.whole_scene(1000 + k)
    k += 1
    f6 = np.concatenate([c.astype(np.float32) / 255.0, n], 1)
    t = tuple(torch.from_numpy(a).to(dev) for a in (p, l, f6))
    try:
        csl.chunk_scene(t[0])
    except ValueError:
        continue
    scans.append(t)
pipes = [ScanNetGeometry(B, 8192, 6, dev, attention=True, seed=d, own_streams=True, grid=True) for d in range(D)]
x, f = synth.scannet_batch(0, B, 8192)
for pl in pipes:
    pl.set_inputs(torch.from_numpy(x).to(dev), torch.from_numpy(f).to(dev))
    pl.forward(True)
torch.cuda.synchronize()
for pl in pipes:
    pl.capture(True)
cur = torch.cuda.current_stream(dev)


def consume(i, chunks):
    p, l, f6 = scans[i % S]
    feats = chunks.gather(f6)
    labels = chunks.gather(l)
    C = chunks.nchunks
    keep = [chunks]
    for b in range((C + B - 1) // B):
        pl = pipes[b % D]
        sel = torch.arange(b * B, b * B + B, device=dev) % C
        bx, bf = chunks.point_sets[sel], feats[sel]
        keep.append((bx, bf))
        pl.main.wait_stream(cur)
        pl.set_inputs(bx, bf)
        pl.replay()
    orig, masks = chunks.orig_idx.reshape(-1), chunks.masks.reshape(-1)
    csl.map_back(chunks.point_sets.reshape(-1, 3), orig, masks, (p.shape[0], 3))
    csl.map_back(labels.reshape(-1), orig, masks, (p.shape[0],))
    return keep


for bg in (False, True):
    np.random.seed(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for chunks, ev in csl.chunk_scenes((scans[k % S][0] for k in range(2 * S)), lookahead=2, background=bg):
        n += 1
    torch.cuda.synchronize()
    print("chunk_scenes alone, background=%s: %.2f ms per scan" % (bg, (time.perf_counter() - t0) * 1e3 / n))
np.random.seed(1)
ready = [csl.chunk_scene(scans[i][0]) for i in range(S)]
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    held = [consume(i, ready[i]) for i in range(S)]
    t_host = time.perf_counter() - t0
    for pl in pipes:
        cur.wait_stream(pl.main)
    torch.cuda.synchronize()
    print("consumer alone: host enqueue %.2f ms per scan, until the GPU is done %.2f ms per scan" %
          (t_host * 1e3 / S, (time.perf_counter() - t0) * 1e3 / S))
