#!/bin/bash
# Builds libsvob200.so for sm_100a in-tree (android_svo_b200/lib/).  -fmad=false: parity with the
# reference's FMA-free x86-64 host build depends on unfused float/double arithmetic.
# The translation units compile in parallel (one nvcc per file); a file is recompiled only when it or a
# header is newer than its object (FORCE=1 rebuilds everything).
set -e
cd "$(dirname "$0")"
mkdir -p lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v $EXTRA_NVCC_FLAGS"
UNITS="pyramid fast sparse_align matcher map_ops glue synth tracker capi"
newest_header=$(ls -t csrc/*.h csrc/*.cuh ../include/*.h build.sh | head -1)
pids=()
names=()
for f in $UNITS; do
  if [ -z "$FORCE" ] && [ -z "$EXTRA_NVCC_FLAGS" ] && [ -f lib/$f.o ] && [ lib/$f.o -nt csrc/$f.cu ] && [ lib/$f.o -nt "$newest_header" ]; then continue; fi
  ( $NVCC $FLAGS -c csrc/$f.cu -o lib/$f.o 2> lib/$f.ptxas.log ) &
  pids+=($!)
  names+=($f)
done
rc=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then cat lib/${names[$i]}.ptxas.log; rm -f lib/${names[$i]}.o; rc=1; fi
done
[ $rc -eq 0 ] || exit 1
OBJS=""
for f in $UNITS; do OBJS="$OBJS lib/$f.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o lib/libsvob200.so $OBJS -lcudart
echo built lib/libsvob200.so
