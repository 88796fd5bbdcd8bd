#!/bin/bash
# usage: build_variant.sh NAME "<extra nvcc flags>" file1.cu [file2.cu ...]
# Links a variant of libmvs_b200.so (multiviewstitch_b200/variants/libmvs_NAME.so) in which the listed
# sources are compiled with extra flags (A/B experiments on the GPU box: MVS_B200_LIB=<path>).
set -e
NAME=$1; FLAGS=$2; shift 2
PKG=multiviewstitch_b200
python -c "from multiviewstitch_b200 import build; build.build()"
mkdir -p $PKG/variants/_obj_$NAME
OBJS=""
for o in $PKG/_build/*.o; do
  b=$(basename $o .o); skip=0
  for f in "$@"; do [ "$(basename $f .cu)" == "$b" ] && skip=1; done
  [ $skip == 0 ] && OBJS="$OBJS $o"
done
for f in "$@"; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function --expt-relaxed-constexpr -DMVS_WITH_NCCL=1 $FLAGS -c $PKG/csrc/$b.cu -o $PKG/variants/_obj_$NAME/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -lcudart -ldl -o $PKG/variants/libmvs_$NAME.so $OBJS $PKG/variants/_obj_$NAME/*.o
echo built $PKG/variants/libmvs_$NAME.so
