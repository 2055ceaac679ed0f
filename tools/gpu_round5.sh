#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_phases.py > gpurun_out/gemm_phases.txt 2>&1; echo "phases rc=$?"; grep -B1 -A12 "encoder w1 glu\|encoder out-proj\|encoder qkv \|encoder w2 \|BIAS M=7968 N=2048" gpurun_out/gemm_phases.txt | grep -v "CTA0" | head -90
timeout 400 python bench.py --steps 48 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['breakdown'], d['config']['tile_policy'])"; tail -3 gpurun_out/bench.err
