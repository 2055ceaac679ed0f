#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_tc_kernel|conv1" -s 7 -c 7 \
   -o gpurun_out/prof_layer -f python tools/prof_layer.py 2 > gpurun_out/prof_layer.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/prof_layer.log; ls -la gpurun_out/
