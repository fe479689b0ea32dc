#!/bin/bash
# debug build of the library with per-phase clock64 accounting in the extrapolation sweep
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/_dbg
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I include"
for u in stencil_ops advect; do nvcc $F -fmad=false -c pyrmt_b200/csrc/$u.cu -o scripts/_dbg/$u.o; done
for u in momentum projection fft; do nvcc $F -c pyrmt_b200/csrc/$u.cu -o scripts/_dbg/$u.o; done
nvcc $F -fmad=false -DRMT_EXT_TIMING -c pyrmt_b200/csrc/extrap.cu -o scripts/_dbg/extrap.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o scripts/_dbg/librmt_b200_dbg.so scripts/_dbg/*.o -lcudart
