#!/usr/bin/env python
"""One whole-model forward (model_pipeline.ScanNetAttentionModel, B = 16) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/model.csv python scripts/model_launches.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pcops_b200 import synth  # noqa: E402
from pcops_b200.model_pipeline import ScanNetAttentionModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
if len(sys.argv) > 2:      # A/B runs: path of an alternative libpcops build
    from pcops_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(sys.argv[2])
x, f = synth.scannet_batch(0, B, 8192)
m = ScanNetAttentionModel(B, 8192, 6)
m.set_inputs(torch.from_numpy(x), torch.from_numpy(f))
for _ in range(2):
    m.forward()
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.forward()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
