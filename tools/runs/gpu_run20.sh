#!/bin/bash
# round-2 GPU call 20 (1 GPU): last validation of the committed tree -- full GPU suite, smoke, default bench line, reference arm
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r20_pytest_all.log 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/r20_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r20_bench_default.json 2> gpurun_out/r20_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r20_bench_reference.json 2> gpurun_out/r20_bench_reference.err
tail -3 gpurun_out/r20_pytest_all.log; tail -1 gpurun_out/r20_smoke.log; tail -c 600 gpurun_out/r20_bench_reference.json
python -c "
import json
s=[l for l in open('gpurun_out/r20_bench_default.json') if l.startswith('{')][-1]
d=json.loads(s); r=d['roofline']; print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], d['e2e']['value'], r['kernel'], r['frac'], d['kernel_launches_per_step'], d['clocks'])
"
