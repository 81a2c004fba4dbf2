#!/bin/bash
# Builds libcsm_b200.so in-tree for sm_100a.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${CSMB_OUT:-$HERE/../libcsm_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS=("$HERE"/*.cu)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared \
  --expt-relaxed-constexpr --extended-lambda -Xptxas -v "$@" \
  -o "$OUT" "${SRCS[@]}"
echo "built $OUT"
