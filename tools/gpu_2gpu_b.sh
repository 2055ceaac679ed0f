#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; wc -l gpurun_out/bench_n2.json; head -c 120 gpurun_out/bench_n2.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload conformer --steps 16 --warmup 3 > gpurun_out/bench_conf_n2.json 2> gpurun_out/bench_conf_n2.err; echo "conf n2 rc=$?"; wc -l gpurun_out/bench_conf_n2.json; python -c "
import json;d=json.load(open('gpurun_out/bench_conf_n2.json'));print('conformer n2', {k:d[k] for k in ['value','ms_per_step','n_gpus']})"; tail -2 gpurun_out/bench_conf_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload conformer --steps 16 --warmup 3 > gpurun_out/bench_conf.json 2> gpurun_out/bench_conf.err; python -c "
import json;d=json.load(open('gpurun_out/bench_conf.json'));print('conformer n1', {k:d[k] for k in ['value','ms_per_step','n_gpus']})"; tail -2 gpurun_out/bench_conf.err
