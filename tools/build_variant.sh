#!/bin/bash
# Build an experimental variant of the library next to the default one:
#   tools/build_variant.sh TAG "-DFOO=1 -DBAR=0" file1.cu file2.cu ...
# recompiles the named sources with the extra defines into build/TAG/ and links them with the default objects of every
# other source into p4-fr-sorry-math-but-love-you_b200/lib/libfrx_TAG.so (load it with FRX_LIBRARY=<path>).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TAG=$1; DEFS=$2; shift 2
CSRC=$ROOT/p4-fr-sorry-math-but-love-you_b200/csrc
mkdir -p $ROOT/build/$TAG
OBJS=""
for src in $CSRC/*.cu; do
  base=$(basename $src .cu)
  obj=$ROOT/build/$base.o
  for v in "$@"; do
    if [ "$v" == "$base.cu" ]; then
      obj=$ROOT/build/$TAG/$base.o
      /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $DEFS -c $src -o $obj &
    fi
  done
  OBJS="$OBJS $obj"
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a $OBJS -o $ROOT/p4-fr-sorry-math-but-love-you_b200/lib/libfrx_$TAG.so
echo built libfrx_$TAG.so
