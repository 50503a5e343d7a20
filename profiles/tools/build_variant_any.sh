#!/bin/bash
# Builds profiles/variants/libtsdgpu_<name>.so = the in-tree library with ONE source file recompiled with extra flags
# (select with TSDGPU_LIB=profiles/variants/libtsdgpu_<name>.so).  Usage: build_variant_any.sh name file.cu "-DX=1 ..."
set -e
cd "$(dirname "$0")/../.."
make -C libtsd_b200/csrc -j8 >/dev/null
mkdir -p profiles/variants
b=$(basename $2 .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Ilibtsd_b200/csrc $3 \
  -c libtsd_b200/csrc/$2 -o /tmp/${b}_$1.o
OBJS=$(ls build/obj/*.o | grep -v "/$b.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o profiles/variants/libtsdgpu_$1.so $OBJS /tmp/${b}_$1.o
echo built profiles/variants/libtsdgpu_$1.so
