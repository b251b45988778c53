#!/bin/bash
# A/B of FPS variants: lone SA1 launch (B = 16 x 8192 -> 1024) timed with events, indices compared with the default library
PKG=pointcloud-segmentation-attention_b200
python - "$@" <<'PY'
import ctypes, sys, os, torch
sys.path.insert(0, os.getcwd())
from pcops_b200 import synth
import numpy as np
x, f = synth.scannet_batch(0, 16, 8192)
xyz = torch.from_numpy(x).cuda()
res = {}
for lib in sys.argv[1:]:
    L = ctypes.CDLL(os.path.join("pointcloud-segmentation-attention_b200", lib))
    L.pc_fps_workspace_bytes.restype = ctypes.c_size_t
    out = torch.empty((16, 1024), dtype=torch.int32, device="cuda")
    oxyz = torch.empty((16, 1024, 3), dtype=torch.float32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: L.pc_fps_gather(16, 8192, 1024, ctypes.c_void_p(xyz.data_ptr()), None, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(oxyz.data_ptr()), st)
    for _ in range(3): assert call() == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[lib] = out.clone()
    print(lib, "fps 16 x 8192 -> 1024: %.1f us (median of 10)" % (sorted(ts)[5] * 1e3), "same indices as first:", torch.equal(out, res[sys.argv[1]]))
PY
