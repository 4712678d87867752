#!/bin/bash
# development aid: build an A/B variant of the library (only dpmm_b200.cu is recompiled) -> dpmmsubclusters.jl_b200/build/var_<name>.so
# usage: tools/build_variant.sh <name> "<extra nvcc flags>"; run with DPMM_LIB_PATH=dpmmsubclusters.jl_b200/build/var_<name>.so
set -e
cd "$(dirname "$0")/../dpmmsubclusters.jl_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $2 -c -o build/api_$1.o csrc/dpmm_b200.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/var_$1.so build/api_$1.o build/smart.o build/niw_[0-5].o -ldl
echo built build/var_$1.so
