#!/bin/bash
# round-2 GPU call 2: GPU tests (full log), new bench incl. extra_configs, A/B of the texture fold (v1/v2) and the C2 vertex aggregation
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_bench_configs.py --deselect tests/test_gpu_dropin_tracker.py > gpurun_out/r2_pytest_old.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_bench_configs.py tests/test_gpu_dropin_tracker.py tests/test_gpu_staging.py tests/test_gpu_shard.py -q -p no:cacheprovider > gpurun_out/r2_pytest_new.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
ab() { timeout 300 env $2 python bench.py --steps 40 --warmup 5 --no-cpu --no-extra 2>>gpurun_out/r2_ab.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], {n[:10]:v['ms_per_step'] for n,v in k.items()})" >> gpurun_out/r2_ab.txt; }
ab main X=1
ab texfold_v1 VHAP_B200_TEXFOLD=v1
ab noagg VHAP_B200_SO=$PWD/vhap_b200/variants/noagg.so
ab main_again X=1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'k_passC2|k_tex_fold2' -o gpurun_out/r02_v2_full python tools/prof_step.py --steps 1 > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_pytest_old.log gpurun_out/r2_pytest_new.log; cat gpurun_out/r2_ab.txt
