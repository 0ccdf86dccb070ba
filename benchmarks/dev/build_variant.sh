#!/bin/bash
# Dev helper: build libmie_b200_<tag>.so with one translation unit recompiled under extra -D flags.
#   benchmarks/dev/build_variant.sh <tag> <file.cu> [-DNAME=VALUE ...]
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
CSRC=$ROOT/medical-image-enhancement-system_b200/csrc
tag=$1; src=$2; shift 2
mkdir -p $CSRC/build/var_$tag
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-extended-lambda -Xptxas -v "$@" \
     -c $CSRC/$src -o $CSRC/build/var_$tag/${src%.cu}.o 2> $CSRC/build/var_$tag/${src%.cu}.ptxas.log
objs=""
for o in $CSRC/build/*.o; do
  b=$(basename $o)
  if [ "$b" == "${src%.cu}.o" ]; then objs="$objs $CSRC/build/var_$tag/$b"; else objs="$objs $o"; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/medical-image-enhancement-system_b200/libmie_b200_$tag.so $objs -ldl
echo built libmie_b200_$tag.so
