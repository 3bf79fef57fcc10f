#!/bin/bash
# build an experimental variant of the library: scripts/build_variant.sh <suffix> <extra nvcc flags...>
suffix=$1; shift
out=build/variants/libkge_b200_$suffix.so
mkdir -p build/variants/obj_$suffix
for f in hopwise_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -w "$@" -I include -I hopwise_b200/csrc -c $f -o build/variants/obj_$suffix/$(basename $f .cu).o &
done
wait
nvcc -shared -o $out build/variants/obj_$suffix/*.o -lcudart_static -ldl -lrt -lpthread && echo $out
