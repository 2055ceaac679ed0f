#!/bin/bash
# round 2 profiles: (1) launch list of the default bench command (per-launch durations, cold-cache and serialised: only the
# SHARES are comparable with bench.py's live numbers), (2) one `--set full` capture of the dominant kernel
# (decode_group_kernel, one launch = the whole 60-step beam search of a 32-utterance batch).  Numbers printed by bench.py
# under ncu are never bench values.
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --min-ms 1 --no-extras --no-cpu-baseline --lanes 1 > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_bench.csv
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:decode_group_kernel -s 1 -c 1 -f -o gpurun_out/r2_decode_group \
    python tools/decode_phases.py 5 > gpurun_out/r2_ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/r2_decode_group.ncu-rep
ncu -i gpurun_out/r2_decode_group.ncu-rep --page raw --csv > gpurun_out/r2_decode_group_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2_decode_group_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum', 'launch__grid_size', 'launch__cluster_size', 'launch__registers_per_thread',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
got = {}
for i, h in enumerate(hdr):
    if h in want:
        print(h, units[i], vals[i])
        got[h] = (units[i], vals[i])
import json
def num(k):
    u, v = got[k]
    x = float(v.replace(',', ''))
    return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
try:
    rd, wr = num('dram__bytes_read.sum'), num('dram__bytes_write.sum')
    json.dump({'decode_group_kernel': rd + wr, 'decode_group_kernel_detail': {'dram_read_bytes': rd, 'dram_write_bytes': wr,
               'source': 'ncu --set full --clock-control none -k regex:decode_group_kernel -s 1 -c 1 python tools/decode_phases.py 5 (one launch = 60 steps, 32 utterances)',
               'metrics': {k: list(v) for k, v in got.items()}}}, open('gpurun_out/r2_traffic.json', 'w'), indent=1)
    print('traffic per launch: %.1f MB read + %.1f MB written' % (rd / 1e6, wr / 1e6))
except Exception as e:
    print('traffic json not written:', e)
PY
