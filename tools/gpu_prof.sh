#!/bin/bash
mkdir -p gpurun_out
# full capture of the dominant GEMM shapes of one encoder layer (second pass = warm code paths)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_tc_kernel" -s 9 -c 5 \
   -o gpurun_out/prof_layer_v4 -f python tools/prof_layer.py 3 > gpurun_out/prof_layer.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/prof_layer.log
for L in 6 8 12; do
timeout 900 python bench.py --steps 48 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err; echo "bench lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_l$L.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['roofline'])"; tail -3 gpurun_out/bench_l$L.err
done
