#!/bin/bash
# A/B builds of one translation unit with extra -D flags: variants/<name>/libdpt_b200.so (same ABI; select it with DPT_B200_LIB).
#   scripts/build_variant.sh <name> <file.cu> "<flags>"
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC="$ROOT/decision-pretrained-transformer_b200/csrc"
name=$1; src=$2; flags=$3
mkdir -p "$ROOT/variants/$name"
make -C "$CSRC" -j8 > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
  -c "$CSRC/$src" -o "$ROOT/variants/$name/${src%.cu}.o"
objs=$(ls "$CSRC"/build/*.o | grep -v "/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/variants/$name/libdpt_b200.so" $objs "$ROOT/variants/$name/${src%.cu}.o"
echo "variants/$name/libdpt_b200.so"
