#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_edges.py -q -m gpu -s --tb=short -p no:cacheprovider -k "persistent or ragged or extreme" > gpurun_out/test_mega.log 2>&1; echo "mega tests rc=$?"
grep -E "persistent decode|passed|failed|^FAILED|^E  " gpurun_out/test_mega.log | cut -c1-220 | head -20
timeout 300 python tools/mega_phases.py 5 50 > gpurun_out/mega_phases.txt 2>&1; echo "phases rc=$?"; cat gpurun_out/mega_phases.txt | tail -60
