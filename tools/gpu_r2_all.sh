#!/bin/bash
# round 2: one GPU call = everything that needs validating (GPU slots on the pod are scarce): persistent-kernel parity first,
# then the whole -m gpu suite (all failures listed), the phase stamps of the persistent kernel and a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_edges.py -q -m gpu -x -s --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_persistent.log 2>&1
echo "persistent tests rc=$?"
grep -E "persistent|passed|failed|^FAILED|^E  |rror" gpurun_out/r2_persistent.log | cut -c1-260 | head -30
timeout 2400 python -m pytest tests -q -m gpu -s --tb=short -p no:cacheprovider --durations=12 > gpurun_out/r2_tests_all.log 2>&1
echo "all gpu tests rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2_tests_all.log | cut -c1-300 | head -80
timeout 300 python tools/decode_phases.py 5 50 > gpurun_out/r2_decode_phases.txt 2>&1; echo "phases rc=$?"; tail -40 gpurun_out/r2_decode_phases.txt
timeout 900 python bench.py --steps 12 --min-ms 300 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"; cat gpurun_out/r2_bench_a.json | cut -c1-3000; tail -5 gpurun_out/r2_bench_a.err
