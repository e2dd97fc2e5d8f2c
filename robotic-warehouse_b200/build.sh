#!/bin/bash
# Builds robotic-warehouse_b200/lib/librware_b200.so for sm_100a (nvcc cross-compiles without a GPU).
# The fused kernel is instantiated in 7 translation units (csrc/rw_inst.cu, -DRW_INST_GROUP=k) compiled in parallel.
set -euo pipefail
cd "$(dirname "$0")"
mkdir -p lib build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v -Xcompiler -fPIC"
pids=()
for g in 0 1 2 3 4 5 6; do
  $NVCC $FLAGS -DRW_INST_GROUP=$g -c -o build/rw_inst_$g.o csrc/rw_inst.cu "$@" > build/rw_inst_$g.log 2>&1 &
  pids+=($!)
done
$NVCC $FLAGS -c -o build/rw_kernels.o csrc/rw_kernels.cu "$@" > build/rw_kernels.log 2>&1 &
pids+=($!)
$NVCC $FLAGS -c -o build/rw_views.o csrc/rw_views.cu "$@" > build/rw_views.log 2>&1 &
pids+=($!)
g++ -O3 -std=c++17 -fPIC -pthread -Wall -c -o build/rw_hostpack.o csrc/rw_hostpack.cpp > build/rw_hostpack.log 2>&1 &
pids+=($!)
$NVCC $FLAGS -c -o build/rw_capi.o csrc/rw_capi.cu "$@" > build/rw_capi.log 2>&1 &
pids+=($!)
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done
cat build/*.log
if [ "$fail" != 0 ]; then echo "nvcc failed" >&2; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared --cudart static -o lib/librware_b200.so build/rw_inst_*.o build/rw_kernels.o build/rw_views.o build/rw_hostpack.o build/rw_capi.o -lpthread
