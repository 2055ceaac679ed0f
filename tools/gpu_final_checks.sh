#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_all.log 2>&1; echo "all gpu tests rc=$?"; tail -n 2 gpurun_out/test_all.log | cut -c1-200; grep -E "^FAILED|^E  " gpurun_out/test_all.log | head -20
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'));print({k:d[k] for k in ['value','ms_per_step','gpu_launches','steps']}, d['e2e'], d['breakdown'], d['clocks']); r=d['roofline']; print(r['bound'], r['achieved'], r['peak'], r['frac'], r['traffic']); print(r['kernel']); print(d.get('cpu_baseline'), d['config']['lanes'], d['config']['tile_policy'])"; tail -3 gpurun_out/bench_final.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_k20.json 2> /dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_k20.json'));print('K=20:', {k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'])"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
timeout 300 python bench.py --workload train --steps 8 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_train.json'));print({k:d[k] for k in ['value','ms_per_step','final_loss']}, d['e2e']['value'], d['roofline']['achieved'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_final.csv
