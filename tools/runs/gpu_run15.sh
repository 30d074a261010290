#!/bin/bash
# round-2 GPU call 15 (2 GPUs): DP parity, N=2 bench (allreduce, peer), per-rank timelines of both
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py -q -p no:cacheprovider -s > gpurun_out/r15_pytest_shard.log 2>&1
bn() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $1 --steps 40 --warmup 5 --no-extra --dp-texture $2 \
     > gpurun_out/r15_bench_n$1_$2.json 2> gpurun_out/r15_bench_n$1_$2.err; }
bn 2 allreduce
bn 2 peer
tl() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29644 tools/timeline.py --dp-texture $1 --per-rank gpurun_out/r15_timeline_n2_$1 > /dev/null 2> gpurun_out/r15_timeline_n2_$1.err; }
tl allreduce
tl peer
tail -3 gpurun_out/r15_pytest_shard.log
for f in gpurun_out/r15_bench_n*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('$f', 'ERR', e)
"; done
