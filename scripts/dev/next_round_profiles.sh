#!/bin/bash
# Round-2 starter: evidence that round 1 ran out of GPU budget for.
#  1. full GPU suite + smoke + default bench (first run of the 2000-step default and of the config-1 / config-4 changes)
#  2. ncu launch list of a training step (config 3) and of the config-4 region
#  3. ncu --set full of the scene-chunk kernels
# Multi-GPU (separately, it is charged N x): gpurun --gpus 2 -- 'bash scripts/gpu_multi.sh 2', then 8.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','steps','e2e','clocks','with_attention_layers','config3_training_step','config4_whole_scene','config1_single_scene_sa1')})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train.csv \
  python bench.py --steps 8 --warmup 3 --skip-cpu --skip-probe --attention-layers 0 --scenes 2 > gpurun_out/ncu_train.log 2>&1
echo "ncu train/config4 list rc=$?"
cat > /tmp/chunk_once.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import pcops_b200
from pcops_b200 import complete_scene_loader as csl, synth
p, l, c, n = synth.whole_scene(1000)
t = [torch.from_numpy(a).cuda() for a in (p, l, c, n)]
for _ in range(3):
    np.random.seed(0)
    out = csl.get_all_subsets_with_all_points_for_scene_numpy(*t)
    csl.map_back(out[0].reshape(-1, 3), out[6].reshape(-1), out[5].reshape(-1), (len(p), 3))
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cell_|chunk_|gather_rows|winner|sample_weights|bbox" -s 9 -c 9 -f \
  -o gpurun_out/prof_scene_chunks python /tmp/chunk_once.py > gpurun_out/ncu_scene_chunks.log 2>&1
echo "ncu scene chunks rc=$?"
