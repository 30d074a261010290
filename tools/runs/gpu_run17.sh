#!/bin/bash
# round-2 GPU call 17 (1 GPU): validation of the final state -- full GPU suite, smoke, default bench line, ncu launch list + full capture
# of the kernels changed since the last capture, graph timeline
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r17_pytest_all.log 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/r17_smoke.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r17_bench_n1.json 2> gpurun_out/r17_bench_n1.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r17_bench_n1_again.json 2> gpurun_out/r17_bench_n1_again.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python tools/prof_step.py --steps 2 > gpurun_out/r17_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'k_passC2|k_pool_scatter|k_pool_count|k_passC1' -o gpurun_out/r02_final_full python tools/prof_step.py --steps 1 > gpurun_out/r17_ncu.log 2>&1
ncu -i gpurun_out/r02_final_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_final_raw.csv 2>/dev/null
timeout 300 python tools/timeline.py > gpurun_out/r17_timeline_n1.txt 2> gpurun_out/r17_timeline_n1.err
tail -3 gpurun_out/r17_pytest_all.log; tail -2 gpurun_out/r17_smoke.log
for f in gpurun_out/r17_bench_n1*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); r=d['roofline']; print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel'], r['frac'], d['kernel_launches_per_step'])
except Exception as e: print('$f', 'ERR', e)
"; done
