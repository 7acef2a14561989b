#!/bin/bash
# Build libaf2b200.so for sm_100a (cross-compiles without a GPU). Static cudart: no runtime dependency on torch's.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v \
  -Xcompiler -fPIC -shared -cudart static \
  -o libaf2b200.so api.cu 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning: v|spill" build.log | grep -v " 0 bytes spill" | head -20 || true
echo "built $(pwd)/libaf2b200.so"
