#!/bin/bash
# round-2 GPU call 4 (2 GPUs): DP parity (default peer+shard and the NCCL baseline), N=2 bench for the variants, single-GPU tests + bench + ncu of the TMA fold
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 900 python -m pytest tests/test_gpu_shard.py -q -p no:cacheprovider -s > gpurun_out/r4_pytest_shard.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_shard.py > gpurun_out/r4_pytest_all.log 2>&1
b2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --steps 40 --warmup 5 --no-extra $2 \
     > gpurun_out/r4_bench_n2_$1.json 2> gpurun_out/r4_bench_n2_$1.err; }
b2 default ""
b2 nccl_slab "--dp-slab nccl"
b2 r1_baseline "--dp-slab nccl --dp-texture allreduce"
timeout 400 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r4_bench_n1.json 2> gpurun_out/r4_bench_n1.err
VHAP_B200_TEXFOLD=reg timeout 400 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r4_bench_n1_regfold.json 2> gpurun_out/r4_bench_n1_regfold.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'k_tex_fold3' -o gpurun_out/r02_fold3_full python tools/prof_step.py --steps 1 > gpurun_out/r4_ncu.log 2>&1
tail -3 gpurun_out/r4_pytest_shard.log gpurun_out/r4_pytest_all.log
for f in gpurun_out/r4_bench_n*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernels'].get('tex_fold_reg_adam'))
except Exception as e: print('$f', 'ERR', e)
"; done
