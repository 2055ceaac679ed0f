#!/bin/bash
# software-barrier A/B: OTB_DG_FLAGS=512 polls with relaxed loads + one acquire fence (256 restores the __threadfence in front of red.release.gpu)
mkdir -p gpurun_out
for F in 0 512 0 512; do OTB_DG_CLUSTER=0 OTB_DG_FLAGS=$F timeout 300 python tools/decode_phases.py 5 > gpurun_out/r2_swbar_f$F.txt 2>&1; echo "flags $F: $(grep -m1 whole gpurun_out/r2_swbar_f$F.txt | cut -c1-90)"; grep -m1 "group barriers" gpurun_out/r2_swbar_f$F.txt | cut -c100-700; done
OTB_DG_CLUSTER=0 OTB_DG_FLAGS=512 timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_swbar_tests.log 2>&1; echo "persistent tests (software barrier, flags 512) rc=$?"; tail -2 gpurun_out/r2_swbar_tests.log | cut -c1-200
