#!/bin/bash
# round-2 GPU call 8 (8 GPUs): scaling at N=8 for the three data-parallel texture-update variants (+ the extra configs once)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r8_ngpu.txt
b8() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 40 --warmup 5 $2 \
     > gpurun_out/r8_bench_n8_$1.json 2> gpurun_out/r8_bench_n8_$1.err; }
b8 peer "--dp-texture peer --no-extra"
b8 allreduce "--dp-texture allreduce --no-extra"
b8 shard "--dp-texture shard --no-extra"
b8 auto_extras ""
for f in gpurun_out/r8_bench_n8_*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['parallelism'][230:300])
    for k,v in (d.get('extra_configs') or {}).items(): print('   ', k, v['value'], v['ms_per_step'], v['e2e']['value'])
except Exception as e: print('$f', 'ERR', e)
"; done
