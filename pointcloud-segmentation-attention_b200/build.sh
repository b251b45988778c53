#!/bin/sh
# Builds libpcops.so (sm_100a only) next to this script.  Usage: sh build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-nvcc}"
exec "$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared \
  -I"$HERE/../include" -I"$HERE/csrc" "$@" \
  -o "$HERE/libpcops.so" \
  "$HERE/csrc/api.cu" "$HERE/csrc/fps.cu" "$HERE/csrc/prob_sample.cu" "$HERE/csrc/ball_query.cu" "$HERE/csrc/group.cu" \
  "$HERE/csrc/segreduce.cu" "$HERE/csrc/interpolate.cu" "$HERE/csrc/grid.cu" "$HERE/csrc/fused.cu" "$HERE/csrc/topk.cu" "$HERE/csrc/attention.cu" "$HERE/csrc/attention_layer.cu" "$HERE/csrc/attention_layer_wide.cu" "$HERE/csrc/scene_chunks.cu"
