#!/bin/bash
# tools/build_variant.sh NAME [nvcc -D flags...]: development build of the library with extra defines ->
# dpmmsubclusters.jl_b200/build/libdpmm_NAME.so (select it with DPMM_LIB_PATH)
set -e
name=$1; shift
cd "$(dirname "$0")/../dpmmsubclusters.jl_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o ../build/api_$name.o dpmm_b200.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/variants/libdpmm_$name.so ../build/api_$name.o ../build/niw_[0-5].o -ldl
echo built tools/variants/libdpmm_$name.so
