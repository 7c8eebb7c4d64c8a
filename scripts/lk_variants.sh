#!/bin/bash
# Builds librdvio_fe.so variants with LK compile-time switches into rd_vio_b200/lib_variants/<name>/ (A/B on the GPU box:
# RDFE_LIB_PATH=rd_vio_b200/lib_variants/<name>/librdvio_fe.so python bench.py ...)
set -e
cd "$(dirname "$0")/../rd_vio_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden --expt-relaxed-constexpr"
for v in "base:" "nounroll:-DLK_NO_UNROLL1" "f2i:-DLK_F2I" "dblcheck:-DLK_DOUBLE_CHECK" "all:-DLK_NO_UNROLL1 -DLK_F2I -DLK_DOUBLE_CHECK"; do
  name="${v%%:*}"; defs="${v#*:}"
  mkdir -p ../lib_variants/$name
  nvcc $FLAGS $defs -c -o /tmp/lk_$name.o lk.cu
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib_variants/$name/librdvio_fe.so _obj/api.o _obj/clahe.o _obj/pyramid.o _obj/harris.o _obj/select.o /tmp/lk_$name.o _obj/undistort.o
  echo built $name
done
