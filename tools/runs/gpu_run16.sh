#!/bin/bash
# round-2 GPU call 16 (2 GPUs): full GPU suite (pool kernels changed), N=2 bench variants with per-rank peer-wait diagnostics
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r16_pytest_all.log 2>&1
bn() { tag=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 40 --warmup 5 --no-extra "$@" \
     > gpurun_out/r16_bench_n2_$tag.json 2> gpurun_out/r16_bench_n2_$tag.err; }
bn allreduce --dp-texture allreduce
VHAP_B200_PDL=0 bn allreduce_nopdl --dp-texture allreduce
bn peer --dp-texture peer
bn allreduce_nopipe --dp-texture allreduce --no-pipeline
bn allreduce_ncclslab --dp-texture allreduce --dp-slab nccl
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r16_bench_n1.json 2> gpurun_out/r16_bench_n1.err
tail -3 gpurun_out/r16_pytest_all.log
for f in gpurun_out/r16_bench_n*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('dp_peer_wait'))
except Exception as e: print('$f', 'ERR', e)
"; done
