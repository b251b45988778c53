#!/bin/sh
# Builds libpcops.so (sm_100a only) next to this script.  Usage: sh build.sh [extra nvcc flags]
# Every csrc/*.cu is compiled to build/<name>/*.o (in parallel, skipped when the object is newer than the source and
# every header) and the objects are linked into one shared library.  PCOPS_OUT=<path>.so builds a variant library
# next to the default one (A/B runs of kernel variants: scripts pass its path to bench.py --lib).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-nvcc}"
OUT="${PCOPS_OUT:-$HERE/libpcops.so}"
OBJ="$HERE/build/$(basename "$OUT" .so)"
mkdir -p "$OBJ"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I$HERE/../include -I$HERE/csrc $*"
# a change of flags rebuilds everything
echo "$FLAGS" | cmp -s - "$OBJ/.flags" || { rm -f "$OBJ"/*.o; echo "$FLAGS" > "$OBJ/.flags"; }
NEWEST_HDR="$(ls -t "$HERE"/csrc/*.cuh "$HERE"/csrc/*.h "$HERE"/../include/*.h 2>/dev/null | head -1)"
TODO=""
for src in "$HERE"/csrc/*.cu; do
  o="$OBJ/$(basename "$src" .cu).o"
  if [ ! -f "$o" ] || [ "$src" -nt "$o" ] || { [ -n "$NEWEST_HDR" ] && [ "$NEWEST_HDR" -nt "$o" ]; }; then
    TODO="$TODO $src"
  fi
done
if [ -n "$TODO" ]; then
  # shellcheck disable=SC2086
  printf '%s\n' $TODO | xargs -P "$(nproc)" -I{} sh -c "$NVCC $FLAGS -c {} -o $OBJ/\$(basename {} .cu).o"
fi
# objects whose source disappeared must not be linked
for o in "$OBJ"/*.o; do
  [ -f "$HERE/csrc/$(basename "$o" .o).cu" ] || rm -f "$o"
done
exec "$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ"/*.o
