#!/bin/bash
# round 2: persistent decode kernel -- parity, phase stamps and lone-batch time with the hardware cluster barrier vs the software
# barrier (OTB_DG_CLUSTER=0), then the benchmark with each decode path.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_edges.py tests/test_gpu_bench_config.py -q -m gpu -x -s --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_persistent_c1.log 2>&1
echo "persistent tests (cluster barrier) rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_persistent_c1.log | cut -c1-260 | head -12
OTB_DG_CLUSTER=0 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_config.py -q -m gpu -x -s --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_persistent_c0.log 2>&1
echo "persistent tests (software barrier) rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_persistent_c0.log | cut -c1-260 | head -12
timeout 300 python tools/decode_phases.py 5 50 > gpurun_out/r2_decode_phases_c1.txt 2>&1; echo "phases(cluster) rc=$?"; grep -vE "^\s+layer 0" gpurun_out/r2_decode_phases_c1.txt | tail -40; grep "layer 0" gpurun_out/r2_decode_phases_c1.txt | cut -c1-900
OTB_DG_CLUSTER=0 timeout 300 python tools/decode_phases.py 30 > gpurun_out/r2_decode_phases_c0.txt 2>&1; echo "phases(software) rc=$?"; grep -vE "^\s+layer 0" gpurun_out/r2_decode_phases_c0.txt | tail -20
timeout 600 python bench.py --steps 12 --min-ms 300 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_p.json 2> gpurun_out/r2_bench_p.err; echo "bench(auto) rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_p.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'lanes', d['config']['lanes'], d['config']['decode_path'])
print('probe', d['config']['persistent_probe']); print('breakdown', d['breakdown']); print('validation', d['validation'])
print('roofline frac', d['roofline']['frac'], d['roofline']['achieved'], d['roofline'].get('whole_step'))
PY
tail -3 gpurun_out/r2_bench_p.err
for L in 2 4; do timeout 600 python bench.py --steps 12 --min-ms 300 --no-extras --no-cpu-baseline --lanes $L > gpurun_out/r2_bench_p_l$L.json 2> gpurun_out/r2_bench_p_l$L.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_p_l$L.json').read().strip().splitlines()[-1]); print('lanes $L', d['value'], d['e2e']['value'], d['config']['decode_path'])"; done
