#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_all.log 2>&1; echo "all gpu tests rc=$?"; tail -n 2 gpurun_out/test_all.log | cut -c1-200
for CV in 1 0; do
OTB_CARVEOUT=$CV timeout 300 python bench.py --steps 48 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cv$CV.json 2> gpurun_out/bench_cv$CV.err; echo "bench carveout=$CV rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_cv$CV.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown']['single_lane_step_ms'], d['breakdown']['encoder_fwd_ms'], d['breakdown']['single_lane_step_ms_under_lane_policy'])"; tail -2 gpurun_out/bench_cv$CV.err
done
OTB_CARVEOUT=1 timeout 300 python bench.py --steps 48 --warmup 3 --lanes 12 --no-cpu-baseline > gpurun_out/bench_cv1_l12.json 2> /dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_cv1_l12.json'));print('lanes 12 carveout=1', {k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'])"
