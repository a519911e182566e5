#!/usr/bin/env bash
# Build libarpde.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/_obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v)
# ARPDE_DEV=1: development build -- the ARPDE_* environment switches and the profiling hooks are compiled in (see common.cuh)
STAMP="$HERE/_obj/.flavour"
FLAVOUR="release"
if [[ "${ARPDE_DEV:-0}" == "1" ]]; then FLAGS+=(-DARPDE_DEV); FLAVOUR="dev"; fi
if [[ ! -f "$STAMP" || "$(cat "$STAMP")" != "$FLAVOUR" ]]; then rm -f "$HERE"/_obj/*.o; echo "$FLAVOUR" > "$STAMP"; fi
# any header newer than an object rebuilds it (every .cu includes most of them)
NEWEST_HDR="$(ls -t "$HERE"/*.cuh "$HERE/../../include/arpde.h" | head -1)"
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/_obj/$(basename "${f%.cu}").o"
  if [[ ! -f "$o" || "$f" -nt "$o" || "$NEWEST_HDR" -nt "$o" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; rm -f "$o"; exit 1; } ) &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && { wait "$p" || rc=1; }; done
[[ $rc -eq 0 ]] || { echo "build failed"; exit 1; }
"$NVCC" -shared -o "$OUT/libarpde.so" "$HERE"/_obj/*.o -lcudart
echo "built $OUT/libarpde.so"
