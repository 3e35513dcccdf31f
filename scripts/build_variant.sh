#!/bin/bash
# build_variant.sh <name> <file.cu> [nvcc -D flags...]: an A/B build of ONE translation unit, linked with the other objects
# into decision-pretrained-transformer_b200/variants/libdpt_b200_<name>.so (selected at run time with DPT_B200_LIB=...)
set -e
cd "$(dirname "$0")/../decision-pretrained-transformer_b200/csrc"
name=$1; file=$2; shift 2
mkdir -p build/var ../variants
base=$(basename $file .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" -c $file -o build/var/${base}_${name}.o 2> build/var/${base}_${name}.ptxas.log
objs=$(ls build/*.o | grep -v "build/${base}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libdpt_b200_${name}.so $objs build/var/${base}_${name}.o
echo built ../variants/libdpt_b200_${name}.so
