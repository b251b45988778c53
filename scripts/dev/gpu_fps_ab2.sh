#!/bin/bash
for s in 256 512 1256 1512 2024; do echo "== FPS variant $s"; PCOPS_FPS_SHAPE=$s timeout 600 python -m pytest tests -m gpu -x -q -k "fps and allpairs" 2>&1 | tail -2; PCOPS_FPS_SHAPE=$s timeout 300 python scripts/opbench.py --ops fps --levels 0 --iters 10 2>&1 | grep -E "^fps_sa1 |fps_sa1_B148"; done
