#!/bin/bash
# Build librd_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --use_fast_math=false"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
if [ -n "$RD_B200_CLEAN_BUILD" ]; then rm -rf build ../librd_b200.so; fi      # RD_B200_CLEAN_BUILD=1: recompile every source
mkdir -p build
pids=()
for f in rd_elementwise rd_conv_direct rd_conv_tc rd_conv_tma rd_conv_halo rd_loss rd_metrics rd_compose rd_runtime; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ rd_common.cuh -nt build/$f.o ] || [ rd_tc_common.cuh -nt build/$f.o ] || [ ../../include/rd_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${RD_PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../librd_b200.so build/rd_elementwise.o build/rd_conv_direct.o build/rd_conv_tc.o build/rd_conv_tma.o build/rd_conv_halo.o build/rd_loss.o build/rd_metrics.o build/rd_compose.o build/rd_runtime.o -lcudart -ldl
echo "built $(cd .. && pwd)/librd_b200.so"
