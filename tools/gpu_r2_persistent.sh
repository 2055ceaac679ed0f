#!/bin/bash
# round 2: first runs of the persistent group kernel (csrc/decode_group.cu): lock-step parity, then the phase stamps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_edges.py tests/test_gpu_bench_config.py -q -m gpu -x -s --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_persistent.log 2>&1
echo "persistent tests rc=$?"
grep -E "persistent|passed|failed|^FAILED|^E  |rror" gpurun_out/r2_persistent.log | cut -c1-260 | head -40
timeout 300 python tools/decode_phases.py 5 50 > gpurun_out/r2_decode_phases.txt 2>&1; echo "phases rc=$?"; tail -45 gpurun_out/r2_decode_phases.txt
