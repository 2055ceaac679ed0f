#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_all.log 2>&1; echo "all gpu tests rc=$?"; tail -n 2 gpurun_out/test_all.log | cut -c1-160; grep -E "^FAILED|^E  " gpurun_out/test_all.log | head
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_last.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown']['single_lane_step_ms'])"
