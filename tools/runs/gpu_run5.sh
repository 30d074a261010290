#!/bin/bash
# round-2 GPU call 5 (2 GPUs): peer-memory / NVLS texture update parity + N=2 bench variants; view-sharing + bench-config tests; N=1 bench with extras
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 900 python -m pytest tests/test_gpu_shard.py -q -p no:cacheprovider -s > gpurun_out/r5_pytest_shard.log 2>&1
timeout 1800 python -m pytest tests/test_gpu_views.py tests/test_gpu_bench_configs.py tests/test_gpu_staging.py tests/test_gpu_parity.py -q -p no:cacheprovider > gpurun_out/r5_pytest_sel.log 2>&1
b2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --steps 40 --warmup 5 --no-extra $2 \
     > gpurun_out/r5_bench_n2_$1.json 2> gpurun_out/r5_bench_n2_$1.err; }
b2 default ""
b2 peer_allreduce "--dp-texture allreduce"
b2 shard "--dp-texture shard"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r5_bench_n1.json 2> gpurun_out/r5_bench_n1.err
tail -3 gpurun_out/r5_pytest_shard.log; tail -3 gpurun_out/r5_pytest_sel.log
for f in gpurun_out/r5_bench_n*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['parallelism'][:140])
    for k,v in (d.get('extra_configs') or {}).items(): print('   ', k, v['value'], v['ms_per_step'], v['e2e']['value'])
    print('   standin', d.get('gpu_eager_standin'))
except Exception as e: print('$f', 'ERR', e)
"; done
