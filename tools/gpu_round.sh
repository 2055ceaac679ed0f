#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/test_gpu_*.log | tail -4; grep -E "^FAILED|^E  " gpurun_out/test_gpu_*.log | head -30
for L in 4 6; do
timeout 900 python bench.py --steps 48 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err; echo "bench lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_l$L.json'));print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['breakdown'], d['roofline']['achieved'], d['clocks'])"; tail -3 gpurun_out/bench_l$L.err
done
timeout 600 python bench.py --workload conformer --steps 10 --warmup 3 > gpurun_out/bench_conformer.json 2> gpurun_out/bench_conformer.err; echo "conformer rc=$?"; cat gpurun_out/bench_conformer.json | cut -c1-400; tail -3 gpurun_out/bench_conformer.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/launches_conformer.csv \
    python bench.py --workload conformer --steps 1 --warmup 3 > gpurun_out/ncu_conf.log 2>&1; echo "ncu rc=$?"
