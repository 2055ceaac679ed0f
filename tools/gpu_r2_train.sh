#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -x -s --tb=short -p no:cacheprovider > gpurun_out/r2_train_tests.log 2>&1; echo "train tests rc=$?"; grep -E "passed|failed|^FAILED|^E  |pre-norm" gpurun_out/r2_train_tests.log | cut -c1-300 | head -12
