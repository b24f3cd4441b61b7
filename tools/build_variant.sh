#!/bin/bash
# Builds a kernel-variant library for A/B measurements on the GPU box: gpurun_variants/libvsg_<name>.so
# (selected at run time with VSG_LIB_PATH).  Usage: tools/build_variant.sh <name> <file.cu> [-DFLAG=VALUE ...]
set -e
NAME=$1; SRC=$2; shift 2
CS=visual_sgraphs_b200/csrc
mkdir -p gpurun_variants/obj_$NAME
make -C $CS -j8 > /dev/null
OBJS=""
for f in $CS/*.cu; do
  b=$(basename $f .cu)
  if [ "$b.cu" == "$SRC" ]; then
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall -Xptxas -v --fmad=false "$@" \
      -c -o gpurun_variants/obj_$NAME/$b.o $f 2> gpurun_variants/obj_$NAME/$b.ptxas.log
    grep -E "Used|spill" gpurun_variants/obj_$NAME/$b.ptxas.log | head -8
    OBJS="$OBJS gpurun_variants/obj_$NAME/$b.o"
  else
    OBJS="$OBJS $CS/$b.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_variants/libvsg_$NAME.so $OBJS
echo built gpurun_variants/libvsg_$NAME.so
