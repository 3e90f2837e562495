#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]  ->  tools/var_<name>.so (A/B builds of the CUDA library)
n=$1; shift
cd "$(dirname "$0")/../simuscop_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off "$@" -shared -o ../../tools/var_$n.so kernels.cu gen_fast.cu gz.cu api.cu tables.cpp
