#!/usr/bin/env bash
# Build libarpde.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/_obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v)
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/_obj/$(basename "${f%.cu}").o"
  if [[ ! -f "$o" || "$f" -nt "$o" || "$HERE/common.cuh" -nt "$o" || "$HERE/tc_ptx.cuh" -nt "$o" || "$HERE/conv_common.cuh" -nt "$o" || "$HERE/../../include/arpde.h" -nt "$o" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; rm -f "$o"; exit 1; } ) &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && { wait "$p" || rc=1; }; done
[[ $rc -eq 0 ]] || { echo "build failed"; exit 1; }
"$NVCC" -shared -o "$OUT/libarpde.so" "$HERE"/_obj/*.o -lcudart
echo "built $OUT/libarpde.so"
