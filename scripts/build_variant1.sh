#!/bin/bash
# variant of ONE translation unit: scripts/build_variant1.sh <suffix> <file.cu> <extra nvcc flags...>; other objects come from build/kge_b200
suffix=$1; f=$2; shift; shift
out=build/variants/libkge_b200_$suffix.so
mkdir -p build/variants/obj_$suffix
b=$(basename $f .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -w "$@" -I include -I hopwise_b200/csrc -c $f -o build/variants/obj_$suffix/$b.o || exit 1
objs=$(ls build/kge_b200/*.o | grep -v "/$b.o")
nvcc -shared -o $out build/variants/obj_$suffix/$b.o $objs -lcudart_static -ldl -lrt -lpthread && echo $out
