#!/bin/sh
# Builds libptgpu.so (CUDA kernels + C ABI) in-tree for sm_100a.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr ${PTGPU_NVCC_FLAGS}"
mkdir -p build
$NVCC $FLAGS -Xptxas -v -c csrc/ptgpu_api.cu -o build/ptgpu_api.o 2> build/ptxas_api.log || { cat build/ptxas_api.log; exit 1; }
$NVCC $FLAGS -c csrc/bvh_wide.cu -o build/bvh_wide.o
$NVCC $FLAGS -c csrc/frame_setup.cu -o build/frame_setup.o
$NVCC $FLAGS -x cu -c csrc/mesh_loader.cc -o build/mesh_loader.o
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libptgpu.so build/ptgpu_api.o build/bvh_wide.o build/frame_setup.o build/mesh_loader.o -lcudart_static -lpthread -ldl -lrt
echo "built $(pwd)/libptgpu.so"
