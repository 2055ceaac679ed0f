#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_all.log 2>&1; echo "all gpu tests rc=$?"; tail -n 3 gpurun_out/test_all.log | cut -c1-200; grep -E "^FAILED|^E  |Error" gpurun_out/test_all.log | head -30
timeout 300 python bench.py --workload train --steps 8 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_train.json'));print({k:d[k] for k in ['value','ms_per_step','final_loss','gpu_launches']}, d['e2e']['value'])"; tail -3 gpurun_out/bench_train.err
timeout 400 python bench.py --steps 48 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown'], d['config']['tile_policy'])"; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --steps 48 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/bench_l1.json 2> gpurun_out/bench_l1.err; echo "bench l1 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_l1.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown']['single_lane_step_ms'], d['breakdown']['encoder_fwd_ms'])"
