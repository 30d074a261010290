#!/bin/bash
# round-2 GPU call 11 (1 GPU): programmatic dependent launch on / off
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r11_pytest_all.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r11_bench_n1_pdl.json 2> gpurun_out/r11_bench_n1_pdl.err
VHAP_B200_PDL=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r11_bench_n1_nopdl.json 2> gpurun_out/r11_bench_n1_nopdl.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r11_bench_n1_pdl2.json 2> gpurun_out/r11_bench_n1_pdl2.err
timeout 300 python bench.py --config nersemble --steps 30 --warmup 5 --no-extra --no-cpu > gpurun_out/r11_bench_n1_ners.json 2> gpurun_out/r11_bench_n1_ners.err
timeout 300 python tools/timeline.py > gpurun_out/r11_timeline_n1.txt 2> gpurun_out/r11_timeline_n1.err
tail -3 gpurun_out/r11_pytest_all.log
for f in gpurun_out/r11_bench_n1*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['kernel_launches_per_step'])
except Exception as e: print('$f', 'ERR', e)
"; done
