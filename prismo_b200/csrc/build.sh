#!/usr/bin/env bash
# Build libfdtd_b200.so in-tree for sm_100a.  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libfdtd_b200.so"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
     -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared ${NVCC_EXTRA:-} \
     -o "$out" "$here/fdtd_engine.cu"
echo "built $out"
