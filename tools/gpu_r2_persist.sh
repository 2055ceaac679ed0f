#!/bin/bash
# round 2: persistent decode kernel v3 -- parity (both barrier kinds), phase stamps, lone-batch time, benchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_edges.py tests/test_gpu_bench_config.py -q -m gpu -x -s --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_persistent_v4c1.log 2>&1
echo "persistent tests (cluster barrier) rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_persistent_v4c1.log | cut -c1-260 | head -12
OTB_DG_CLUSTER=0 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_config.py -q -m gpu -x -s --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_persistent_v4c0.log 2>&1
echo "persistent tests (software barrier) rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_persistent_v4c0.log | cut -c1-260 | head -12
timeout 300 python tools/decode_phases.py 5 50 > gpurun_out/r2_decode_phases_v4.txt 2>&1; echo "phases(cluster) rc=$?"; grep -vE "^\s+layer 0" gpurun_out/r2_decode_phases_v4.txt | tail -40; grep -E "layer 0|layer 2|tail," gpurun_out/r2_decode_phases_v4.txt | cut -c1-1400
for L in 2 3; do timeout 600 python bench.py --steps 12 --min-ms 300 --no-extras --no-cpu-baseline --lanes $L > gpurun_out/r2_bench_v4_l$L.json 2> gpurun_out/r2_bench_v4_l$L.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_v4_l$L.json').read().strip().splitlines()[-1]); print('lanes $L', round(d['value']), round(d['e2e']['value']), d['config']['decode_path'], d['config']['persistent_probe'][:120], d['breakdown']['single_lane_step_ms'], d['breakdown']['persistent_decode_kernel_ms'], d['validation']['ids_sha1'])"; done
OTB_DG_CLUSTER=0 timeout 600 python bench.py --steps 12 --min-ms 300 --no-extras --no-cpu-baseline --lanes 3 > gpurun_out/r2_bench_v4_sw3.json 2> gpurun_out/r2_bench_v4_sw3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_v4_sw3.json').read().strip().splitlines()[-1]); print('software barrier, lanes 3', round(d['value']), round(d['e2e']['value']), d['breakdown']['single_lane_step_ms'])"
