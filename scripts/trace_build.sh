#!/bin/bash
# Debug build with phase tracing (clock64 deltas printed by one warpgroup): build/libgnnb_trace.so; use with GNNB_LIB=...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/trace
for u in gnnb_api gnnb_simt gnnb_prop gnnb_tc gnnb_prop_tc gnnb_babsr gnnb_train gnnb_queue gnnb_kw; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DGNNB_TRACE \
      -c gnn_branching_b200/csrc/$u.cu -o build/trace/$u.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o build/libgnnb_trace.so build/trace/*.o -gencode arch=compute_100a,code=sm_100a
ls -la build/libgnnb_trace.so
