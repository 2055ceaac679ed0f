#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/decode_phases.py 5 > gpurun_out/r2_decode_phases_bar_c1.txt 2>&1; grep -E "whole|group barriers" gpurun_out/r2_decode_phases_bar_c1.txt | cut -c1-900
OTB_DG_CLUSTER=0 timeout 300 python tools/decode_phases.py 5 > gpurun_out/r2_decode_phases_bar_c0.txt 2>&1; echo "software barrier:"; grep -E "whole|group barriers" gpurun_out/r2_decode_phases_bar_c0.txt | cut -c1-900
