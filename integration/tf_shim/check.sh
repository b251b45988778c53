#!/bin/sh
# Syntax-check the TensorFlow shim without TensorFlow: g++ -fsyntax-only against stub/ (declarations only).
# The real build line is in INTEGRATION.md section 2.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
for f in sampling_ops.cc grouping_ops.cc interpolation_ops.cc attention_ops.cc; do
  g++ -std=c++14 -fsyntax-only -Wall -I"$HERE/stub" -I"$HERE/../../include" "$HERE/$f"
  echo "ok: $f"
done
