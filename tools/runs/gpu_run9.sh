#!/bin/bash
# round-2 GPU call 9 (4 GPUs): N=4 and N=2 for the peer / allreduce texture updates (auto policy)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
bn() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $1 --steps 40 --warmup 5 --no-extra --dp-texture $2 \
     > gpurun_out/r9_bench_n$1_$2.json 2> gpurun_out/r9_bench_n$1_$2.err; }
bn 4 peer
bn 4 allreduce
bn 2 peer
bn 2 allreduce
for f in gpurun_out/r9_bench_n*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('$f', 'ERR', e)
"; done
