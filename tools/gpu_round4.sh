#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_all.log 2>&1; echo "all gpu tests rc=$?"; tail -n 3 gpurun_out/test_all.log | cut -c1-200
for L in 8 1 12; do
timeout 300 python bench.py --steps 48 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err; echo "bench lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_l$L.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown']['single_lane_step_ms'], d['breakdown']['encoder_fwd_ms'])"; tail -2 gpurun_out/bench_l$L.err
done
OTB_LN_CLUSTER=1 timeout 300 python bench.py --steps 48 --warmup 3 --lanes 8 --no-cpu-baseline > gpurun_out/bench_l8_cluster.json 2> gpurun_out/bench_l8_cluster.err; echo "bench lanes=8 LN cluster rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_l8_cluster.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown']['single_lane_step_ms'])"
