#!/bin/bash
# Prints the n-best id digests of bench.py's input batch 0 for both decode paths on a B200; the values are recorded by hand in
# tests/golden/bench_digest.json (keys gpu_graph / gpu_persistent), where bench.py's `validation.match` reads them.
mkdir -p gpurun_out
for path in graph persistent; do
  python bench.py --steps 4 --min-ms 100 --no-extras --no-cpu-baseline --decode $path > gpurun_out/r2_digest_$path.json 2> gpurun_out/r2_digest_$path.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_digest_$path.json').read().strip().splitlines()[-1]); print('$path', d['config']['decode_path'], d['validation'])"
done
