#!/bin/bash
# round 2, 2 GPUs: the bench contract at N = 2 (recognize: utterance batches sharded over ranks, no data-path collective;
# training step: one NCCL all-reduce of the flat gradient per optimizer step) with NCCL_DEBUG=INFO captured by bench.py (`comm`)
mkdir -p gpurun_out
for W in transformer train; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 12 --warmup 3 --workload $W --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n2_$W.json 2> gpurun_out/r2_bench_n2_$W.err
  echo "$W N=2 rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_n2_$W.json').read().strip().splitlines()[-1])
    print(round(d['value'], 1), d['unit'], 'e2e', round(d['e2e']['value'], 1), 'n_gpus', d['n_gpus'], 'comm', json.dumps(d.get('comm'))[:600])
except Exception as e:
    print('parse failed', e)
PY
done
for W in transformer train; do
  timeout 600 python bench.py --steps 12 --warmup 3 --workload $W --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n1_$W.json 2> gpurun_out/r2_bench_n1_$W.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_$W.json').read().strip().splitlines()[-1]); print('$W N=1', round(d['value'],1), d['unit'])"
done
