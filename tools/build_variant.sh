#!/bin/bash
# Build an A/B variant of libgdpt_cuda.so (extra -D flags) into gdpathtracing_b200/ab/ (git-ignored; travels with gpurun).
# usage: tools/build_variant.sh <name> [-DNAME=VALUE ...]
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gdpathtracing_b200/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared "$@" -I include \
  -o gdpathtracing_b200/ab/libgdpt_cuda_$name.so gdpathtracing_b200/csrc/cuda/pt_kernels.cu gdpathtracing_b200/csrc/cuda/gdpt_capi.cu
echo built $name
