#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x -s --tb=short -p no:cacheprovider -k "full_row_groups or barrier_kinds" > gpurun_out/r2_onetest.log 2>&1; echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  |identical" gpurun_out/r2_onetest.log | cut -c1-300 | head -12
