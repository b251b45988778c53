#!/bin/bash
for s in 256 1 2 3 512 4 1024; do echo -n "FPS variant $s: "; PCOPS_FPS_SHAPE=$s timeout 300 python scripts/opbench.py --ops fps --levels 0 --iters 10 2>&1 | grep -E "^fps_sa1 "; done
