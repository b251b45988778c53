#!/bin/bash
# Round-2 starter: try the 2-CTA cluster / multicast variant of the wide attention layer (written without a GPU at the
# end of round 1, never run).  Applies the patch, rebuilds, runs the attention-layer parity tests and the timing script.
#   PCOPS_ATTN_WIDE_CLUSTER=1 forces the single-CTA schedule for an A/B in the same build.
set -e
cd "$(dirname "$0")/../.."
git apply scripts/dev/attention_layer_wide_cluster.patch
sh pointcloud-segmentation-attention_b200/build.sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "attention_layer" 2>&1 | tail -5
timeout 300 python scripts/dev/time_attlayer.py 2>&1 | tail -5
PCOPS_ATTN_WIDE_CLUSTER=1 timeout 300 python scripts/dev/time_attlayer.py 2>&1 | tail -5
