#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -q -m gpu -s --tb=short -p no:cacheprovider > gpurun_out/test_train_ops.log 2>&1; echo "train ops rc=$?"
tail -n 80 gpurun_out/test_train_ops.log
