#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 96 --warmup 3 --lanes 16 --tile-policy latency --no-cpu-baseline > gpurun_out/bench_lat_l16.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_lat_l16.json'));print('latency policy lanes=16', {k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'])"
for L in 12 20; do
timeout 300 python bench.py --steps 96 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_c32_l$L.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_c32_l$L.json'));print('throughput policy lanes=$L', {k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'])"
done
