#!/bin/bash
# Builds an experimental variant of the library next to the product one (A/B runs on the GPU box swap the files):
#   bash tools/build_variant.sh NAME -DFOO=1 ...   ->  hmse_b200/libhmse_b200_NAME.so   (git-ignored)
set -eu
NAME=$1; shift
HERE=$(cd "$(dirname "$0")/.." && pwd)
OBJ=$HERE/hmse_b200/build_$NAME
mkdir -p $OBJ
for s in $HERE/hmse_b200/csrc/*.cu; do
  o=$OBJ/$(basename ${s%.cu}).o
  if [ "$(basename $s)" = "deflate.cu" ] || [ ! -f $o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c $s -o $o &
  fi
done
wait
nvcc -shared -o $HERE/hmse_b200/libhmse_b200_$NAME.so $OBJ/*.o -Xcompiler -fPIC -lcudart_static -ldl -lrt -lpthread
echo built $HERE/hmse_b200/libhmse_b200_$NAME.so
