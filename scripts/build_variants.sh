#!/bin/bash
# Build kernel-tuning variants of the library as gpurun_variants/lib_<tag>.so (for sweeps on the GPU box).
# usage: scripts/build_variants.sh "tag1:-DFOO=1 -DBAR=2" "tag2:..."
set -e
cd "$(dirname "$0")/../halotools_b200/csrc"
mkdir -p ../../variants
make -s mesh.o capi.o
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC $flags -c count.cu -o /tmp/count_$tag.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$tag.so mesh.o /tmp/count_$tag.o capi.o
  echo built $tag
done
