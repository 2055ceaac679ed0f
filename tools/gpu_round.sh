#!/bin/bash
# One GPU visit: tests, smoke, bench (b200 + reference), ncu launch list.  Logs under gpurun_out/.
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/test_gpu_*.log | tail -4
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
nproc
