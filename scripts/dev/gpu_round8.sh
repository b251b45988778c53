#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "attention_layer" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_quick.log
timeout 300 python scripts/dev/time_attlayer.py 2>&1 | tail -12
