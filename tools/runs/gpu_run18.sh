#!/bin/bash
# round-2 GPU call 18 (8 GPUs): scaling points of the final state -- N = 8 (auto = peer+nvls, with the extra configurations), N = 8 allreduce,
# N = 4 and N = 2 (auto), N = 2 peer
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
bn() { n=$1; tag=$2; shift; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --steps 40 --warmup 5 "$@" \
     > gpurun_out/r18_bench_n${n}_$tag.json 2> gpurun_out/r18_bench_n${n}_$tag.err; }
bn 8 auto_extras
bn 8 allreduce --no-extra --dp-texture allreduce
bn 4 auto --no-extra
bn 2 auto --no-extra
bn 2 peer --no-extra --dp-texture peer
for f in gpurun_out/r18_bench_n*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('dp_peer_wait'), {k:(v['value'],v['ms_per_step']) for k,v in d.get('extra_configs',{}).items()})
except Exception as e: print('$f', 'ERR', e)
"; done
