#!/bin/bash
# insurance run of the bench contract at N = 8 (one box): the driver's scaling run uses exactly this launch line
mkdir -p gpurun_out
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 12 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "N=$N rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
    print(round(d['value']), d['unit'], 'e2e', round(d['e2e']['value']), 'n_gpus', d['n_gpus'], d['config'].get('decode_path'), d['config'].get('group_barrier'), 'comm', json.dumps(d.get('comm'))[:300], d['validation']['match'])
except Exception as e:
    print('parse failed', e)
PY
