#!/bin/bash
# round 2 validation run: the whole GPU test suite, the default bench line, the graph-path bench line, smoke()
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short -p no:cacheprovider > gpurun_out/r2_tests_final.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_tests_final.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
keep = {k: d[k] for k in ('value', 'e2e', 'ms_per_step', 'gpu_launches', 'roofline', 'validation', 'clocks') if k in d}
keep['config'] = {k: d['config'][k] for k in ('lanes', 'decode_path', 'group_barrier', 'persistent_probe') if k in d['config']}
keep['breakdown'] = d.get('breakdown')
keep['cpu_baseline'] = d.get('cpu_baseline')
keep['extras'] = d.get('extras')
print(json.dumps(keep)[:3000])
PY
timeout 600 python bench.py --decode graph --no-extras --no-cpu-baseline > gpurun_out/r2_bench_graph.json 2> gpurun_out/r2_bench_graph.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_graph.json').read().strip().splitlines()[-1]); print('graph path', round(d['value']), round(d['e2e']['value']), d['config']['lanes'], d['validation']['ids_sha1'], d['validation']['match'])"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi_final.txt
