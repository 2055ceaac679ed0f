#!/bin/bash
# software-barrier A/B: OTB_DG_FLAGS=256 restores the __threadfence in front of red.release.gpu
mkdir -p gpurun_out
OTB_DG_CLUSTER=0 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_config.py -q -m gpu -x --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_swbar_tests.log 2>&1; echo "persistent tests (software barrier) rc=$?"; tail -2 gpurun_out/r2_swbar_tests.log | cut -c1-200
for F in 0 256 0 256; do OTB_DG_CLUSTER=0 OTB_DG_FLAGS=$F timeout 300 python tools/decode_phases.py 5 > gpurun_out/r2_swbar_f$F.txt 2>&1; echo "flags $F: $(grep -m1 whole gpurun_out/r2_swbar_f$F.txt | cut -c1-90)"; grep -m1 "group barriers" gpurun_out/r2_swbar_f$F.txt | cut -c100-700; done
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 48 > gpurun_out/r2_swbar_bench.json 2> gpurun_out/r2_swbar_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_swbar_bench.json').read().strip().splitlines()[-1]); print('bench', round(d['value']), round(d['e2e']['value']), d['validation']['match'], d['validation']['one_best_equal_to_bf16_policy_oracle'])"
OTB_DG_FLAGS=256 timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 48 > gpurun_out/r2_swbar_bench256.json 2> gpurun_out/r2_swbar_bench256.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_swbar_bench256.json').read().strip().splitlines()[-1]); print('bench with threadfence', round(d['value']), round(d['e2e']['value']), d['validation']['match'])"
