#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
CUDA_VISIBLE_DEVICES=0 timeout 120 python tools/launch_rate.py 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train --steps 8 --warmup 3 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "train n2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_train_n2.json'));print({k:d[k] for k in ['value','ms_per_step','n_gpus','final_loss']}, d['e2e']['value'])"; tail -3 gpurun_out/bench_train_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_n2.json'));print({k:d[k] for k in ['value','ms_per_step','n_gpus']}, d['e2e']['value'])"; tail -3 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/bench_ref_n2.json
