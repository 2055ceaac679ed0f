#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/test_gpu_*.log | tail -4; grep -E "^FAILED|^E  " gpurun_out/test_gpu_*.log | head -30
for L in 6; do
timeout 900 python bench.py --steps 48 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err; echo "bench lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_l$L.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown'], d['roofline'])"; tail -3 gpurun_out/bench_l$L.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_enc.csv \
    python tools/prof_layer.py 6 > gpurun_out/ncu_enc.log 2>&1; echo "ncu rc=$?"
