#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short -p no:cacheprovider > gpurun_out/r2_tests_final.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_tests_final.log | cut -c1-300
