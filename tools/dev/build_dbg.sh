#!/bin/bash
# development builds of libshdr with -DSHDR_C1_DBG=<v> in conv1_fused.cu only -> tools/_build/libshdr_dbg<v>.so
set -e
cd "$(dirname "$0")/../../singlehdr-tf2_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets"
for v in "$@"; do
  nvcc $FLAGS -DSHDR_C1_DBG=$v -c conv1_fused.cu -o build/conv1_dbg$v.o
  nvcc -gencode arch=compute_100a,code=sm_100a -Wno-deprecated-gpu-targets -shared -o ../../tools/_build/libshdr_dbg$v.so \
    build/api.o build/invcrf.o build/frontend.o build/pooled.o build/pooled_ws.o build/pooled_slide.o build/pooled_slide_nb.o build/backward.o build/conv1_dbg$v.o
done
