#!/bin/bash
# usage: tools/build_variant.sh NAME -DFOO=1 ...   -> tools/variants/libnnmd_NAME.so (tuning experiments; select with NNMD_B200_LIB)
set -e
cd "$(dirname "$0")/../nnmd_b200/csrc"
name=$1; shift
mkdir -p ../../tools/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c sf.cu -o /tmp/sf_$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../tools/variants/libnnmd_$name.so $(ls *.o | grep -v '^sf\.o$') /tmp/sf_$name.o -lcudart
echo built $name
