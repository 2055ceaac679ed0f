#!/bin/bash
# End-of-round verification: every GPU test, the benchmark contract (b200 + reference arms), the secondary workloads,
# the ncu launch list of one lone recognize pass and ncu --set full captures of the forward / backward kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_all.log 2>&1; echo "all gpu tests rc=$?"; tail -n 3 gpurun_out/test_all.log | cut -c1-200; grep -E "^FAILED|^E  " gpurun_out/test_all.log | head -20
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'));print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e'], d['breakdown'], d['clocks']); r=d['roofline']; print(r['bound'], r['achieved'], r['peak'], r['frac'], r['traffic']); print(r['kernel']); print(d.get('cpu_baseline'))"; tail -3 gpurun_out/bench_final.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/bench_ref.json
timeout 300 python bench.py --workload train --steps 8 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_train.json'));print({k:d[k] for k in ['value','ms_per_step','final_loss']}, d['e2e']['value'], d['roofline']['achieved'])"
timeout 300 python bench.py --workload conformer --steps 16 --warmup 3 > gpurun_out/bench_conf.json 2> gpurun_out/bench_conf.err; echo "conf rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_conf.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['roofline']['achieved'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_final.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_tc_kernel|conv1|layernorm" -s 9 -c 9 \
   -o gpurun_out/prof_layer_final -f python tools/prof_layer.py 3 > gpurun_out/prof_layer.log 2>&1; echo "ncu layer rc=$?"; tail -1 gpurun_out/prof_layer.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd|layernorm_bwd|glu_bwd|colsum|im2col|col2im|conv1_wgrad|gemm_tc_kernel<128, 0, 1>|ls_ce" -s 60 -c 16 \
   -o gpurun_out/prof_train_final -f python tools/prof_train.py 2 > gpurun_out/prof_train.log 2>&1; echo "ncu train rc=$?"; tail -1 gpurun_out/prof_train.log
ls -la gpurun_out/*.ncu-rep
