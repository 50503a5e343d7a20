#!/bin/bash
# Builds profiles/variants/libtsdgpu_<name>.so = the in-tree library with ols16k.cu recompiled with extra flags
# (kernel A/B experiments; select with TSDGPU_LIB=profiles/variants/libtsdgpu_<name>.so).  Usage: build_variant.sh name "-DX=1 ..."
set -e
cd "$(dirname "$0")/../.."
make -C libtsd_b200/csrc -j8 >/dev/null
mkdir -p profiles/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Ilibtsd_b200/csrc $2 \
  -c libtsd_b200/csrc/ols16k.cu -o /tmp/ols16k_$1.o
OBJS=$(ls build/obj/*.o | grep -v ols16k.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o profiles/variants/libtsdgpu_$1.so $OBJS /tmp/ols16k_$1.o
echo built profiles/variants/libtsdgpu_$1.so
