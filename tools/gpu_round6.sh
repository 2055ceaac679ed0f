#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_all.log 2>&1; echo "all gpu tests (PDL on) rc=$?"; tail -n 3 gpurun_out/test_all.log | cut -c1-200; grep -E "^FAILED|^E  " gpurun_out/test_all.log | head -20
for L in 8 1; do
timeout 300 python bench.py --steps 48 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_pdl_l$L.json 2> gpurun_out/bench_pdl_l$L.err; echo "bench PDL lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_pdl_l$L.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown'])"; tail -2 gpurun_out/bench_pdl_l$L.err
done
OTB_PDL=0 timeout 300 python bench.py --steps 48 --warmup 3 --lanes 8 --no-cpu-baseline > gpurun_out/bench_nopdl_l8.json 2> gpurun_out/bench_nopdl_l8.err; echo "bench no-PDL lanes=8 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_nopdl_l8.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown'])"
timeout 300 python bench.py --workload train --steps 8 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_train.json'));print({k:d[k] for k in ['value','ms_per_step','final_loss']}, d['e2e']['value'])"; tail -3 gpurun_out/bench_train.err
timeout 300 python bench.py --workload conformer --steps 16 --warmup 3 > gpurun_out/bench_conf.json 2> gpurun_out/bench_conf.err; echo "conf rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_conf.json'));print({k:d[k] for k in ['value','ms_per_step']})"
