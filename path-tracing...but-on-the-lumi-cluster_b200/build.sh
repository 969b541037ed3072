#!/bin/sh
# Builds libptgpu.so (CUDA kernels + C ABI) in-tree for sm_100a.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr ${PTGPU_NVCC_FLAGS}"
OUT=${PTGPU_OUT:-libptgpu.so}          # experimental builds: PTGPU_OUT=libptgpu_x.so PTGPU_NVCC_FLAGS=-D... (load with PTGPU_LIB=)
B=${PTGPU_BUILD_DIR:-build}
mkdir -p $B
$NVCC $FLAGS -Xptxas -v -c csrc/ptgpu_api.cu -o $B/ptgpu_api.o 2> $B/ptxas_api.log || { cat $B/ptxas_api.log; exit 1; }
$NVCC $FLAGS -c csrc/bvh_wide.cu -o $B/bvh_wide.o
$NVCC $FLAGS -c csrc/frame_setup.cu -o $B/frame_setup.o
$NVCC $FLAGS -x cu -c csrc/mesh_loader.cc -o $B/mesh_loader.o
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $B/ptgpu_api.o $B/bvh_wide.o $B/frame_setup.o $B/mesh_loader.o -lcudart_static -lpthread -ldl -lrt
echo "built $(pwd)/$OUT"
