#!/bin/bash
# where does the fused conv1 kernel spend its time: full | no feature generation | no weight traffic | neither
mkdir -p gpurun_out
P="python tools/dev/probe_conv1.py 1 --nolib"
echo "full:      $(timeout 120 $P)"
for v in 1 2 3; do echo "dbg$v:      $(SHDR_LIB=$PWD/tools/_build/libshdr_dbg$v.so timeout 120 $P)"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frontend_conv1 -c 1 -f -o gpurun_out/conv1_v1 $P > gpurun_out/ncu_conv1.log 2>&1
tail -3 gpurun_out/ncu_conv1.log
