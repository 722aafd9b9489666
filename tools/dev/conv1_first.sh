#!/bin/bash
# first bring-up of the tcgen05 path: descriptor unit test, then the conv1 parity tests (each under its own timeout)
mkdir -p gpurun_out
timeout 60 tools/_build/umma_nosw > gpurun_out/umma.log 2>&1; echo "umma rc=$?" >> gpurun_out/umma.log
cat gpurun_out/umma.log
if grep -q "PASS-SWAPPED" gpurun_out/umma.log; then
  echo "rebuilding with SHDR_UMMA_SWAP"; bash singlehdr-tf2_b200/csrc/build.sh -DSHDR_UMMA_SWAP > /dev/null
fi
timeout 600 python -m pytest tests/test_gpu_conv1.py -x -q 2>&1 | tail -30 | tee gpurun_out/conv1_tests.log
