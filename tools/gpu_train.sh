#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -s --tb=short -p no:cacheprovider -x > gpurun_out/test_train.log 2>&1; echo "train rc=$?"
tail -n 60 gpurun_out/test_train.log
