#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_all.log 2>&1; echo "all gpu tests rc=$?"; tail -n 5 gpurun_out/test_all.log
timeout 600 python bench.py --steps 32 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['breakdown']); print(json.dumps(d['roofline'], indent=1)); print(d.get('cpu_baseline'))"; tail -3 gpurun_out/bench.err
for L in 8 12; do
timeout 300 python bench.py --steps 48 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err; echo "bench lanes=$L rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_l$L.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'])"; tail -2 gpurun_out/bench_l$L.err
done
timeout 600 python bench.py --workload train --steps 8 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench rc=$?"; cat gpurun_out/bench_train.json; tail -5 gpurun_out/bench_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"; wc -l gpurun_out/launches_train.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 74 -c 37 \
   -o gpurun_out/prof_decode_gemms -f python tools/prof_decode.py 4 > gpurun_out/prof_decode.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/prof_decode.log; ls -la gpurun_out/*.ncu-rep
