#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_all.log 2>&1; echo "all gpu tests rc=$?"; tail -n 8 gpurun_out/test_all.log | cut -c1-300
timeout 600 python bench.py --steps 48 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['breakdown']); r=d['roofline']; print(r['kernel']); print(r['frac'], r['traffic'], r['all_gemm_shapes_ms_per_pass']); print(d.get('cpu_baseline'), d['config']['lanes'])"; tail -3 gpurun_out/bench.err
for L in 1 12; do
timeout 300 python bench.py --steps 48 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err; echo "bench lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_l$L.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown'])"; tail -2 gpurun_out/bench_l$L.err
done
timeout 600 python bench.py --workload train --steps 8 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_train.json'));print({k:d[k] for k in ['value','ms_per_step','gpu_launches','final_loss']}, d['e2e']['value'], d['roofline']['achieved'])"; tail -5 gpurun_out/bench_train.err
timeout 300 python tools/gemm_phases.py > gpurun_out/gemm_phases.txt 2>&1; echo "phases rc=$?"; grep -A9 "encoder w1 glu\|encoder out-proj\|encoder qkv \|BIAS M=7968 N=2048" gpurun_out/gemm_phases.txt | head -60
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
