#!/bin/bash
# Build kernel-tuning variants of the library as variants/lib_<tag>.so (for sweeps on the GPU box).
# usage: scripts/build_variants.sh "tag1:-DFOO=1 -DBAR=2" "tag2:..."
set -e
cd "$(dirname "$0")/../halotools_b200/csrc"
mkdir -p ../../variants
make -s mesh.o capi.o
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -Xptxas -v $flags -c count.cu -o /tmp/count_$tag.o 2> /tmp/count_$tag.log \
    && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$tag.so mesh.o /tmp/count_$tag.o capi.o \
    && echo "built $tag: $(grep -A2 'Fast3' /tmp/count_$tag.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')" || (echo "FAILED $tag"; tail -5 /tmp/count_$tag.log) ) &
done
wait
