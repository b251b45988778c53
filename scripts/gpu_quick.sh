#!/bin/bash
# quick parity + opbench of selected ops: gpu_quick.sh "<pytest -k expr>" "<opbench ops>" [levels]
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_quick.log
timeout 600 python scripts/opbench.py --ops $2 --levels ${3:-0,1,2,3} 2>&1 | grep -v "^{" | grep -v "fps_sa1_B"
