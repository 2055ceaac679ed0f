#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -s --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/test_mega.log 2>&1; echo "mega tests rc=$?"
tail -n 60 gpurun_out/test_mega.log
bash tools/run_gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/test_gpu_*.log | tail -4; grep -E "^FAILED|^E  " gpurun_out/test_gpu_*.log | head -20
for L in 1 2 6; do
timeout 300 python bench.py --steps 16 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_mega_l$L.json 2> gpurun_out/bench_mega_l$L.err; echo "bench lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_mega_l$L.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown'])"; tail -3 gpurun_out/bench_mega_l$L.err
done
