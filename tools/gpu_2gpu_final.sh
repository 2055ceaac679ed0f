#!/bin/bash
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_k20.json 2> /dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_k20.json'));print('N=1 K=20:', {k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['config']['lanes'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 96 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_n2.json'));print({k:d[k] for k in ['value','ms_per_step','n_gpus']}, d['e2e']['value'])"; tail -2 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train --steps 8 --warmup 3 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "train n2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_train_n2.json'));print({k:d[k] for k in ['value','ms_per_step','n_gpus','final_loss']}, d['e2e']['value'])"; tail -2 gpurun_out/bench_train_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload conformer --steps 16 --warmup 3 > gpurun_out/bench_conf_n2.json 2> /dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_conf_n2.json'));print('conformer n2', {k:d[k] for k in ['value','ms_per_step','n_gpus']})"
