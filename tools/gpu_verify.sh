#!/bin/bash
# One-call verification of the current tree on the B200 box: GPU test suites, headline bench, reference arm,
# ncu launch list of one bench step.  Everything lands under gpurun_out/.
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/test_gpu_*.log | tail -4; grep -E "^FAILED|^E  " gpurun_out/test_gpu_*.log | head -30
timeout 600 python bench.py --steps 32 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 8 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/bench_l1.json 2> gpurun_out/bench_l1.err; echo "bench l1 rc=$?"; cat gpurun_out/bench_l1.json
timeout 300 python bench.py --workload conformer --steps 16 --warmup 3 > gpurun_out/bench_conf.json 2> gpurun_out/bench_conf.err; echo "conf rc=$?"; cat gpurun_out/bench_conf.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/launches.csv
