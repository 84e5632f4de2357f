#!/bin/bash
# Build an experimental variant of the library: tools/build_variant.sh NAME "-DSP_X=1 ..."
# SRC=<dir> takes the .cu/.cuh sources from another directory (e.g. an older revision checked out under /tmp).
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
SRC=${SRC:-sponet_b200/csrc}
mkdir -p sponet_b200/_variants/obj_$NAME
for f in api replay native_cnvm native_cntm cv graphgen; do
  EXTRA=""; [ $f = replay ] && EXTRA="-fmad=false"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off --expt-relaxed-constexpr $EXTRA "$@" -c $SRC/$f.cu -o sponet_b200/_variants/obj_$NAME/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o sponet_b200/_variants/lib_$NAME.so sponet_b200/_variants/obj_$NAME/*.o -lcudart
echo built sponet_b200/_variants/lib_$NAME.so
