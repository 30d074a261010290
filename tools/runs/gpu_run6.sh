#!/bin/bash
# round-2 GPU call 6 (2 GPUs): DP parity after the barrier merge / lean gradient fold, N=2 variants, N=1 A/B of the C2 split
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_modular.py tests/test_gpu_parity.py -q -p no:cacheprovider -s > gpurun_out/r6_pytest.log 2>&1
b2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --steps 40 --warmup 5 --no-extra $2 \
     > gpurun_out/r6_bench_n2_$1.json 2> gpurun_out/r6_bench_n2_$1.err; }
b2 default ""
b2 peer_allreduce "--dp-texture allreduce"
b2 shard "--dp-texture shard"
timeout 400 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r6_bench_n1.json 2> gpurun_out/r6_bench_n1.err
VHAP_B200_SO=$PWD/vhap_b200/variants/nosplit.so timeout 400 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r6_bench_n1_nosplit.json 2> gpurun_out/r6_bench_n1_nosplit.err
tail -3 gpurun_out/r6_pytest.log
for f in gpurun_out/r6_bench_n*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], {k[:10]:v['ms_per_step'] for k,v in list(d['kernels'].items())[:6]})
except Exception as e: print('$f', 'ERR', e)
"; done
