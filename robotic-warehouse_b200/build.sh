#!/bin/bash
# Builds robotic-warehouse_b200/lib/librware_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
mkdir -p lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v \
  -Xcompiler -fPIC -shared --cudart static \
  -o lib/librware_b200.so csrc/rw_kernels.cu csrc/rw_capi.cu "$@"
