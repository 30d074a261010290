#!/bin/bash
# round-2 GPU call 3 (2 GPUs): data-parallel parity on real NCCL + scaling at N=2 for both texture-update variants
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r3_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_shard.py -q -p no:cacheprovider -s > gpurun_out/r3_pytest_shard.log 2>&1
for tex in shard allreduce; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --steps 40 --warmup 5 --no-extra --dp-texture $tex \
     > gpurun_out/r3_bench_n2_$tex.json 2> gpurun_out/r3_bench_n2_$tex.err
done
timeout 400 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err
tail -4 gpurun_out/r3_pytest_shard.log
for f in gpurun_out/r3_bench_n*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('$f', 'ERR', e)
"; done
