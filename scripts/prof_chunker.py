import sys, cProfile, pstats, numpy as np, torch
sys.path.insert(0, ".")   # run from the repository root
from pcops_b200 import synth
from pcops_b200 import complete_scene_loader as csl
dev = torch.device("cuda:0")
scans = []
k = 0
while len(scans) < 20:
    p = synth.whole_scene(1000 + k)[0]; k += 1
    t = torch.from_numpy(p).to(dev)
    try: csl.chunk_scene(t)
    except ValueError: continue
    scans.append(t)
np.random.seed(1)
for c, ev in csl.chunk_scenes(iter(scans), lookahead=2): pass
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for c, ev in csl.chunk_scenes(iter(scans * 2), lookahead=2): pass
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(18)
