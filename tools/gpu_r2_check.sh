#!/bin/bash
# quick round-2 check of a kernel change: persistent-path tests, phase stamps (and the previous build on the same box), default bench line
TAG=${TAG:-vx}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_edges.py tests/test_gpu_bench_config.py -q -m gpu -x --tb=short -p no:cacheprovider > gpurun_out/r2_tests_${TAG}.log 2>&1; echo "model/edge/bench-config tests rc=$?"; tail -2 gpurun_out/r2_tests_${TAG}.log | cut -c1-200
timeout 300 python tools/decode_phases.py 5 50 > gpurun_out/r2_decode_phases_${TAG}.txt 2>&1; grep -E "whole|layer 2, per|tail, per" gpurun_out/r2_decode_phases_${TAG}.txt | awk '!s[$0]++' | cut -c1-1250
if [ -f opentransformer_b200/libotb200_prev.so ]; then OTB_LIB_PATH=$PWD/opentransformer_b200/libotb200_prev.so timeout 300 python tools/decode_phases.py 5 50 > gpurun_out/r2_decode_phases_${TAG}_prevlib.txt 2>&1; echo "previous build, same box:"; grep -E "whole|layer 2, per" gpurun_out/r2_decode_phases_${TAG}_prevlib.txt | awk '!s[$0]++' | cut -c1-700; fi
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_bench_${TAG}.json 2> gpurun_out/r2_bench_${TAG}.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_${TAG}.err | cut -c1-200
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_${TAG}.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['config'].get('lanes'), d['config'].get('decode_path'), d['config'].get('group_barrier'))
print('breakdown', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d['breakdown'].items() if k != 'timed_region_ms'})
print('validation', d['validation']['ids_sha1'], d['validation']['match'], d['validation']['one_best_equal_to_bf16_policy_oracle'])
print('roofline', round(d['roofline']['achieved'], 1), d['roofline']['frac'], d['clocks'])
PY
