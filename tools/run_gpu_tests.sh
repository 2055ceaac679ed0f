#!/bin/bash
# Run the GPU suites in separate processes (a trapped kernel poisons only its own process) with
# bounded time, collecting logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in tests/test_gpu_ops.py tests/test_gpu_model.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s --tb=short -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?; echo "$f exit $r"; tail -n 40 gpurun_out/$n.log; [ $r -ne 0 ] && rc=$r
done
exit $rc
