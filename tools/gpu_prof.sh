#!/bin/bash
mkdir -p gpurun_out
# decode: skip encoder (9 kernels) + setup (6 gemm + init) + 2 full steps (53 each), capture 1 step
timeout 1200 ncu --set full --clock-control none --import-source on -s 122 -c 53 \
   -o gpurun_out/prof_decode -f python tools/prof_decode.py 4 > gpurun_out/prof_decode.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/prof_decode.log; ls -la gpurun_out/*.ncu-rep
