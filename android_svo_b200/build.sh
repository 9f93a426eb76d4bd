#!/bin/bash
# Builds libsvob200.so for sm_100a in-tree (android_svo_b200/lib/).  -fmad=false: parity with the
# reference's FMA-free x86-64 host build depends on unfused float/double arithmetic.
set -e
cd "$(dirname "$0")"
mkdir -p lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v"
for f in pyramid fast sparse_align matcher map_ops glue synth tracker capi; do
  $NVCC $FLAGS -c csrc/$f.cu -o lib/$f.o 2> lib/$f.ptxas.log || { cat lib/$f.ptxas.log; exit 1; }
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o lib/libsvob200.so lib/pyramid.o lib/fast.o lib/sparse_align.o lib/matcher.o lib/map_ops.o lib/glue.o lib/tracker.o lib/synth.o lib/capi.o -lcudart
echo built lib/libsvob200.so
