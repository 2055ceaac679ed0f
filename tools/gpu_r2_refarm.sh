#!/bin/bash
# the reference arm and the cpu_baseline leg on the GPU box with oracle/_ref travelling along
mkdir -p gpurun_out
ls oracle/_ref/BUILD_INFO && cat oracle/_ref/BUILD_INFO
timeout 600 python -m pytest tests/test_oracle_ref.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref_arm.json 2> gpurun_out/r2_ref_arm.err; echo "reference arm rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_ref_arm.json').read().strip().splitlines()[-1]); print(round(d['value'],3), d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])"
timeout 900 python bench.py --no-extras --steps 24 > gpurun_out/r2_bench_with_ref.json 2> gpurun_out/r2_bench_with_ref.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_with_ref.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['cpu_baseline'], d['validation']['oracle_check'])"
