#!/usr/bin/env bash
# Builds libdvid_b200.so (sm_100a only) in-tree. Used by __graft_entry__.build().  Fails if any translation unit fails.
set -euo pipefail
cd "$(dirname "$0")"
SRC=diffusionvid_b200/csrc
OUT=diffusionvid_b200/_C
mkdir -p "$OUT" build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-Xptxas -v -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
objs=()
pids=()
for f in $SRC/*.cu; do
  o=build/$(basename "${f%.cu}").o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC include -name '*.h' -newer "$o" -o -name '*.cuh' -newer "$o")" ]; then
    echo "nvcc $f"
    rm -f "$o"
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
  objs+=("$o")
done
fail=0
for p in "${pids[@]:-}"; do
  if [ -n "$p" ] && ! wait "$p"; then fail=1; fi
done
if [ "$fail" -ne 0 ]; then echo "build.sh: compilation failed" >&2; exit 1; fi
$NVCC -shared -o "$OUT/libdvid_b200.so" "${objs[@]}" -ldl
echo "built $OUT/libdvid_b200.so"
