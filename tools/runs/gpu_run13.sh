#!/bin/bash
# round-2 GPU call 13 (1 GPU): pools with 4 pixels per thread, dynamic work queue in the backward passes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r13_pytest_all.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r13_bench_n1.json 2> gpurun_out/r13_bench_n1.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r13_bench_n1_again.json 2> gpurun_out/r13_bench_n1_again.err
timeout 300 python tools/timeline.py > gpurun_out/r13_timeline_n1.txt 2> gpurun_out/r13_timeline_n1.err
tail -3 gpurun_out/r13_pytest_all.log
for f in gpurun_out/r13_bench_n1*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); r=d['roofline']; print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel'], r['frac'], r['avg_launch_ms'], {k:v['avg_launch_ms'] for k,v in r['other_kernels'].items()})
except Exception as e: print('$f', 'ERR', e)
"; done
