#!/bin/bash
# round-2 GPU call 14 (1 GPU): fewer launches (self-clearing scan, merged clears, camera set-up folded), pass C2 block-shape variants
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r14_pytest_all.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r14_bench_n1.json 2> gpurun_out/r14_bench_n1.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r14_bench_n1_again.json 2> gpurun_out/r14_bench_n1_again.err
VHAP_B200_SO=$PWD/vhap_b200/variants/c2pb128.so timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r14_bench_n1_c2pb128.json 2> gpurun_out/r14_bench_n1_c2pb128.err
VHAP_B200_SO=$PWD/vhap_b200/variants/c2mb3.so timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r14_bench_n1_c2mb3.json 2> gpurun_out/r14_bench_n1_c2mb3.err
timeout 300 python tools/timeline.py > gpurun_out/r14_timeline_n1.txt 2> gpurun_out/r14_timeline_n1.err
tail -3 gpurun_out/r14_pytest_all.log
for f in gpurun_out/r14_bench_n1*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); r=d['roofline']; print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel'], r['frac'], r['avg_launch_ms'], {k:v['avg_launch_ms'] for k,v in r['other_kernels'].items()})
except Exception as e: print('$f', 'ERR', e)
"; done
