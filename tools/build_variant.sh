#!/bin/bash
# Build a variant of deconv.o with extra -D flags and link it with the other (unchanged) objects into
# gpurun_out/variants/lib_<tag>.so   usage: tools/build_variant.sh <tag> "<-D flags>"
set -e
cd "$(dirname "$0")/.."
tag=$1; flags=$2
mkdir -p build/variants gpurun_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr $flags -c impulcifer_pip313_b200/csrc/deconv.cu -o build/variants/deconv_$tag.o
others=$(ls build/obj/*.o | grep -v deconv.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_variants/lib_$tag.so build/variants/deconv_$tag.o $others -lcudart_static -lpthread -ldl -lrt
echo built gpurun_variants/lib_$tag.so
