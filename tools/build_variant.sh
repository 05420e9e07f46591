#!/bin/bash
# usage: tools/build_variant.sh NAME SRC.cu -DFOO=1 ...   -> tools/variants/libnnmd_NAME.so
# Tuning experiments: SRC.cu (in nnmd_b200/csrc) is recompiled with the given flags and linked with the objects of the
# regular build; -DNNMD_TUNING turns on the NNMD_* environment knobs the product library ignores (run `python -m nnmd_b200._build` first); select the result with NNMD_B200_LIB (tools/run_variants.sh).
set -e
cd "$(dirname "$0")/../nnmd_b200/csrc"
name=$1; src=$2; shift 2
base=${src%.cu}
mkdir -p ../../tools/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DNNMD_TUNING "$@" -c $src -o /tmp/${base}_$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../tools/variants/libnnmd_$name.so $(ls *.o | grep -v "^${base}\.o$") /tmp/${base}_$name.o -lcudart
echo built $name
