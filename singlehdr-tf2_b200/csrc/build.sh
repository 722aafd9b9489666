#!/bin/bash
# Build libshdr.so (sm_100a only) next to the Python package.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libshdr.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v -Wno-deprecated-gpu-targets"
mkdir -p build
pids=()
for f in api invcrf frontend pooled pooled_ws pooled_slide pooled_slide_nb backward conv1_fused; do
  $NVCC $FLAGS "$@" -c $f.cu -o build/$f.o > build/$f.log 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
if [ $rc -ne 0 ]; then cat build/*.log; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -Wno-deprecated-gpu-targets -shared -o $OUT build/api.o build/invcrf.o build/frontend.o build/pooled.o build/pooled_ws.o build/pooled_slide.o build/pooled_slide_nb.o build/backward.o build/conv1_fused.o
echo "built $(realpath $OUT)"
