"""Marginal cost of the parts of the pipelined step: us per batch (D batches in flight, graph replay) with the FPS chain
only, with everything but the FPS chain, with and without the attention contraction.  If fps-only + side-only is
about the full step the two contend for the same resource; if the full step is about the larger one they overlap."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcops_b200 import _lib, synth               # noqa: E402
from pcops_b200.pipeline import ScanNetGeometry   # noqa: E402

B, N, D, K = 16, 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 8, 400
if len(sys.argv) > 2:
    _lib.LIB_PATH = os.path.abspath(sys.argv[2])
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
_lib.set_concurrency_hint(int(os.environ.get('HINT', D)))   # HINT=1: lone-launch shapes, for comparison
x, f = synth.scannet_batch(0, B, N)
dx, df = torch.from_numpy(x).to(dev), torch.from_numpy(f).to(dev)
cur = torch.cuda.current_stream(dev)


def measure(parts, attention, skip=(), fuse_fp=False):
    pipes = [ScanNetGeometry(B, N, 6, dev, attention=attention, seed=d, own_streams=True, grid=True) for d in range(D)]
    for pl in pipes:                 # a full forward first: the side ops of a "side"-only run need FPS results
        pl.set_inputs(dx, df)
        pl.forward(True)
    torch.cuda.synchronize()
    for pl in pipes:
        pl.parts = tuple(parts)
        pl.skip = tuple(skip)
        pl.fuse_fp = fuse_fp
        pl.capture(True)

    def run(steps):
        for pl in pipes:
            pl.main.wait_stream(cur)
        for i in range(steps):
            pipes[i % D].replay()
        for pl in pipes:
            cur.wait_stream(pl.main)
    run(4 * D)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    run(K)
    e1.record(cur)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K


for skip in (("query_ball",), ("three_nn",), ("group_",), ("three_interpolate",), ("three_weights",),
             ("group_", "three_interpolate", "three_weights"), ("query_ball", "three_nn")):
    print("side only, no attention, without %-40s %.1f us per batch" % (" ".join(skip), measure(("side",), False, skip)))
print("side only, no attention, pc_fp_interpolate instead of weights + interpolate: %.1f us per batch" % measure(("side",), False, (), True))
print("whole step with pc_fp_interpolate: %.1f us per batch" % measure(("fps", "side"), True, (), True))
for parts, att in ((("fps", "side"), True), (("fps", "side"), False), (("fps",), False), (("side",), True), (("side",), False)):
    print("parts %-16s attention %-5s  %.1f us per batch" % ("+".join(parts), att, measure(parts, att)))
