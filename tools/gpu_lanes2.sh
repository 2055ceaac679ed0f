#!/bin/bash
mkdir -p gpurun_out
for L in 16 24 32; do
timeout 300 python bench.py --steps 96 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_c32_l$L.json 2> gpurun_out/bench_c32_l$L.err; echo "conn=32 lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_c32_l$L.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'])"; tail -2 gpurun_out/bench_c32_l$L.err
done
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 300 python bench.py --steps 96 --warmup 3 --lanes 16 --no-cpu-baseline > gpurun_out/bench_c8_l16.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_c8_l16.json'));print('conn=8 lanes=16', {k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'])"
