#!/bin/bash
# round 2 iteration run for the persistent decode kernel: TAG=<version tag>; micro-benchmarks, parity, phase stamps, bench
TAG=${TAG:-v5}
mkdir -p gpurun_out
if [ -n "$MICRO" ]; then
  for m in $MICRO; do
    nvcc -arch=sm_100a -O3 -o /tmp/$m tools/micro/$m.cu -lcuda 2> gpurun_out/r2_micro_$m.err && timeout 120 /tmp/$m > gpurun_out/r2_micro_$m.txt 2>&1
    echo "micro $m rc=$?"; head -60 gpurun_out/r2_micro_$m.txt
  done
fi
[ -z "$SKIP_TESTS" ] && { timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_edges.py tests/test_gpu_bench_config.py -q -m gpu -x -s --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_persistent_${TAG}c1.log 2>&1; }
[ -z "$SKIP_TESTS" ] && { echo "persistent tests (cluster barrier) rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_persistent_${TAG}c1.log | cut -c1-260 | head -12; }
[ -z "$SKIP_TESTS" ] && { OTB_DG_CLUSTER=0 timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x -s --tb=short -p no:cacheprovider -k "persistent" > gpurun_out/r2_persistent_${TAG}c0.log 2>&1; }
[ -z "$SKIP_TESTS" ] && { echo "persistent tests (software barrier) rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_persistent_${TAG}c0.log | cut -c1-260 | head -12; }
timeout 300 python tools/decode_phases.py 5 50 > gpurun_out/r2_decode_phases_${TAG}.txt 2>&1; echo "phases(cluster) rc=$?"; grep -vE "^\s+layer 0" gpurun_out/r2_decode_phases_${TAG}.txt | tail -40; grep -E "layer 0|layer 2|tail," gpurun_out/r2_decode_phases_${TAG}.txt | cut -c1-1400
if [ -z "$SKIP_BENCH" ]; then for L in ${LANES:-3}; do timeout 600 python bench.py --steps 12 --min-ms 300 --no-extras --no-cpu-baseline --lanes $L > gpurun_out/r2_bench_${TAG}_l$L.json 2> gpurun_out/r2_bench_${TAG}_l$L.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_${TAG}_l$L.json').read().strip().splitlines()[-1]); print('lanes $L', round(d['value']), round(d['e2e']['value']), d['config']['decode_path'], d['config']['persistent_probe'][:120], d['breakdown']['single_lane_step_ms'], d['breakdown']['persistent_decode_kernel_ms'], d['validation']['ids_sha1'])"; done; fi
if [ -z "$SKIP_BENCH" ]; then OTB_DG_CLUSTER=0 timeout 600 python bench.py --steps 12 --min-ms 300 --no-extras --no-cpu-baseline --lanes 3 > gpurun_out/r2_bench_${TAG}_sw3.json 2> gpurun_out/r2_bench_${TAG}_sw3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_${TAG}_sw3.json').read().strip().splitlines()[-1]); print('software barrier, lanes 3', round(d['value']), round(d['e2e']['value']), d['breakdown']['single_lane_step_ms'])"; fi
if [ -n "$FLAGS_AB" ]; then for F in $FLAGS_AB; do OTB_DG_FLAGS=$F timeout 300 python tools/decode_phases.py 5 50 > gpurun_out/r2_decode_phases_${TAG}_f$F.txt 2>&1; echo "flags $F"; grep -E "whole|layer 2, per" gpurun_out/r2_decode_phases_${TAG}_f$F.txt | cut -c1-420; done; fi
if [ -f opentransformer_b200/libotb200_prev.so ] && [ -n "$PREV_AB" ]; then OTB_LIB_PATH=$PWD/opentransformer_b200/libotb200_prev.so timeout 300 python tools/decode_phases.py 5 50 > gpurun_out/r2_decode_phases_${TAG}_prevlib.txt 2>&1; echo "previous build on the same box:"; grep -E "whole|layer 2, per|tail" gpurun_out/r2_decode_phases_${TAG}_prevlib.txt | cut -c1-1300; fi
