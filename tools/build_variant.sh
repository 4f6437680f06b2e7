#!/bin/bash
# build_variant.sh <out.so> <file.cu> <extra nvcc flags...>: relink libsgx_b200.so with ONE translation unit rebuilt
# with extra -D flags, for A/B timing through SGX_LIB=<out.so>.  Run the normal build first.
set -e
cd "$(dirname "$0")/.."
out=$1; src=$2; shift 2
pkg=group_gan_gcn_gat_b200
base=$(basename "$src" .cu)
mkdir -p group_gan_gcn_gat_b200/build/variants
obj=group_gan_gcn_gat_b200/build/variants/$(basename "$out" .so)_$base.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $pkg/csrc/$base.cu -o $obj
objs=$(ls $pkg/build/*.o | grep -v "/$base.o")
nvcc -shared -o $out $objs $obj -gencode arch=compute_100a,code=sm_100a -lcudart
echo built $out
