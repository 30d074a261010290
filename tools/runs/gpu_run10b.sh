#!/bin/bash
# round-2 GPU call 10b (1 GPU): full GPU suite, smoke, bench (+ A/B: no-pipeline, 4-row TMA ring), ncu launch list, timeline
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r10_pytest_all.log 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/r10_smoke.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r10_bench_n1.json 2> gpurun_out/r10_bench_n1.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu --no-pipeline > gpurun_out/r10_bench_n1_nopipe.json 2> gpurun_out/r10_bench_n1_nopipe.err
VHAP_B200_SO=$PWD/vhap_b200/variants/ns4.so timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r10_bench_n1_ns4.json 2> gpurun_out/r10_bench_n1_ns4.err
VHAP_B200_TEXFOLD_PAD_KB=50 timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r10_bench_n1_pad50.json 2> gpurun_out/r10_bench_n1_pad50.err
VHAP_B200_TEXFOLD_PAD_KB=150 timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r10_bench_n1_pad150.json 2> gpurun_out/r10_bench_n1_pad150.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r10_bench_n1_again.json 2> gpurun_out/r10_bench_n1_again.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python tools/prof_step.py --steps 2 > gpurun_out/r10_ncu_launches.log 2>&1
timeout 300 python tools/timeline.py > gpurun_out/r10_timeline_n1.txt 2> gpurun_out/r10_timeline_n1.err
tail -3 gpurun_out/r10_pytest_all.log; tail -2 gpurun_out/r10_smoke.log
for f in gpurun_out/r10_bench_n1*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['kernel_launches_per_step'])
except Exception as e: print('$f', 'ERR', e)
"; done
