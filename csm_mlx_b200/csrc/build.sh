#!/bin/bash
# Builds libcsm_b200.so in-tree for sm_100a: one nvcc -c per source in parallel, then one link.
# Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${CSMB_OUT:-$HERE/../libcsm_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OBJ="${CSMB_OBJ_DIR:-$HERE/.obj}"
mkdir -p "$OBJ"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr --extended-lambda -Xptxas -v "$@")
pids=()
for src in "$HERE"/*.cu; do
  o="$OBJ/$(basename "${src%.cu}").o"
  # rebuild when the source, any header of this directory, the public header or this script is newer than the object
  if [ ! -f "$o" ] || [ -n "$(find "$src" "$HERE"/frame_kernel.cu "$HERE"/*.cuh "$HERE/../../include/csm_b200.h" "${BASH_SOURCE[0]}" -newer "$o" 2>/dev/null | head -1)" ] || [ $# -gt 0 ]; then
    "$NVCC" "${FLAGS[@]}" -c -o "$o" "$src" 2> "$o.log" &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [ -z "$p" ] || wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$OBJ"/*.log >&2; exit 1; fi
cat "$OBJ"/*.log >&2 || true
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "$OUT" "$OBJ"/*.o
echo "built $OUT"
