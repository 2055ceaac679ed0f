#!/bin/bash
# One GPU visit: tests, smoke, bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/test_gpu_*.log | tail -4; grep -E "^FAILED|^E  " gpurun_out/test_gpu_*.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python tools/gemm_phases.py > gpurun_out/gemm_phases.txt 2>&1; grep -E "events|epi_done|tfull_seen" gpurun_out/gemm_phases.txt
for L in 4 1; do
timeout 900 python bench.py --steps 16 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err; echo "bench lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_l$L.json'));print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['breakdown'], d['roofline']['achieved'], d['clocks'])"; tail -3 gpurun_out/bench_l$L.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
