#!/bin/bash
# tools/build_variant.sh NAME [nvcc flags...]: a tuning build of the library as variants/libpngb200_NAME.so
# (select it with PNGB200_LIB=...; experiments only, the product is swift-png_b200/libpngb200.so)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
    -DPNGB200_BUILD "$@" -o variants/libpngb200_$name.so swift-png_b200/csrc/*.cu
echo variants/libpngb200_$name.so
