#!/bin/bash
# same-box A/B of the default bench line: current build vs opentransformer_b200/libotb200_prev.so
mkdir -p gpurun_out
for rep in 1 2; do
  for which in cur prev; do
    if [ $which = prev ]; then export OTB_LIB_PATH=$PWD/opentransformer_b200/libotb200_prev.so; else unset OTB_LIB_PATH; fi
    timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 48 > gpurun_out/r2_ab_${which}_$rep.json 2> gpurun_out/r2_ab_${which}_$rep.err
    python -c "
import json; d=json.loads(open('gpurun_out/r2_ab_${which}_$rep.json').read().strip().splitlines()[-1]); print('$which $rep', round(d['value']), round(d['e2e']['value']), d['breakdown']['single_lane_step_ms'], d['config']['group_barrier'])"
  done
done
