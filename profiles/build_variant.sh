#!/bin/bash
# build_variant.sh TAG "-DEX_MINBLK=1 ..." : a tuning variant of the library (extend.cu recompiled with
# the given macros) as fastga_b200/libfastga_b200_TAG.so; select it with FGB_LIB=<path>.
set -e
cd "$(dirname "$0")/../fastga_b200/csrc"
TAG=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $@ -c extend.cu -o /tmp/extend_$TAG.o
nvcc -shared -o ../libfastga_b200_$TAG.so sort128.o gix.o merge.o /tmp/extend_$TAG.o filter.o seams.o trace.o api.o -gencode arch=compute_100a,code=sm_100a
echo built ../libfastga_b200_$TAG.so
