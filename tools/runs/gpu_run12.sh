#!/bin/bash
# round-2 GPU call 12 (1 GPU): graph node latency micro-benchmark, pass C2 software prefetch variant
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 120 tools/micro/graph_gap > gpurun_out/r12_graph_gap.txt 2>&1
VHAP_B200_SO=$PWD/vhap_b200/variants/c2pf.so timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r12_bench_n1_c2pf.json 2> gpurun_out/r12_bench_n1_c2pf.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r12_bench_n1_base.json 2> gpurun_out/r12_bench_n1_base.err
VHAP_B200_SO=$PWD/vhap_b200/variants/c2pf.so timeout 300 python -m pytest tests/test_gpu_bench_configs.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r12_pytest_c2pf.log 2>&1
cat gpurun_out/r12_graph_gap.txt; tail -2 gpurun_out/r12_pytest_c2pf.log
for f in gpurun_out/r12_bench_n1*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); r=d['roofline']; print('$f', d['value'], d['ms_per_step'], r['kernel'], r['avg_launch_ms'], {k:v['avg_launch_ms'] for k,v in r['other_kernels'].items()})
except Exception as e: print('$f', 'ERR', e)
"; done
