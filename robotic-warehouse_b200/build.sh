#!/bin/bash
# Builds robotic-warehouse_b200/$OUT/librware_b200.so for sm_100a (nvcc cross-compiles without a GPU).
# The fused kernel is instantiated in 7 translation units (csrc/rw_inst.cu, -DRW_INST_GROUP=k) compiled in parallel.
set -euo pipefail
cd "$(dirname "$0")"
OUT=${RW_LIB_OUT:-lib}; BUILD=${RW_BUILD_DIR:-build}
mkdir -p $OUT $BUILD
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v -Xcompiler -fPIC"
pids=()
for g in 0 1 2 3 4 5 6; do
  $NVCC $FLAGS -DRW_INST_GROUP=$g -c -o $BUILD/rw_inst_$g.o csrc/rw_inst.cu "$@" > $BUILD/rw_inst_$g.log 2>&1 &
  pids+=($!)
done
$NVCC $FLAGS -c -o $BUILD/rw_kernels.o csrc/rw_kernels.cu "$@" > $BUILD/rw_kernels.log 2>&1 &
pids+=($!)
$NVCC $FLAGS -c -o $BUILD/rw_views.o csrc/rw_views.cu "$@" > $BUILD/rw_views.log 2>&1 &
pids+=($!)
g++ -O3 -std=c++17 -fPIC -pthread -Wall -c -o $BUILD/rw_hostpack.o csrc/rw_hostpack.cpp > $BUILD/rw_hostpack.log 2>&1 &
pids+=($!)
$NVCC $FLAGS -c -o $BUILD/rw_capi.o csrc/rw_capi.cu "$@" > $BUILD/rw_capi.log 2>&1 &
pids+=($!)
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done
cat $BUILD/*.log
if [ "$fail" != 0 ]; then echo "nvcc failed" >&2; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared --cudart static -o $OUT/librware_b200.so $BUILD/rw_inst_*.o $BUILD/rw_kernels.o $BUILD/rw_views.o $BUILD/rw_hostpack.o $BUILD/rw_capi.o -lpthread
