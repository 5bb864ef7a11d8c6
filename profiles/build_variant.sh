#!/bin/bash
# build_variant.sh TAG FILE.cu "-DX=1 ..." : a tuning variant of the library (FILE.cu recompiled with the
# given macros) as fastga_b200/libfastga_b200_TAG.so; select it with FGB_LIB=<path>.
set -e
cd "$(dirname "$0")/../fastga_b200/csrc"
TAG=$1; FILE=$2; shift; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $@ -c $FILE -o /tmp/${FILE%.cu}_$TAG.o
OBJS=""
for f in sort128 gix merge extend filter seams trace api; do
  if [ "$f.cu" == "$FILE" ]; then OBJS="$OBJS /tmp/${f}_$TAG.o"; else OBJS="$OBJS $f.o"; fi
done
nvcc -shared -o ../libfastga_b200_$TAG.so $OBJS -gencode arch=compute_100a,code=sm_100a
echo built ../libfastga_b200_$TAG.so
