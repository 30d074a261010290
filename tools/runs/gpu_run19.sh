#!/bin/bash
# round-2 GPU call 19 (1 GPU): occupancy variants of passes A / C1
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for v in base c1m3 c1m5 c1m6 am3 am6; do
  if [ $v = base ]; then unset VHAP_B200_SO; else export VHAP_B200_SO=$PWD/vhap_b200/variants/$v.so; fi
  timeout 200 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu > gpurun_out/r19_bench_n1_$v.json 2> gpurun_out/r19_bench_n1_$v.err
done
for f in gpurun_out/r19_bench_n1*.json; do python -c "
import json,sys
try:
    s=[l for l in open('$f') if l.startswith('{')][-1]
    d=json.loads(s); k=d['kernels']; print('$f', d['value'], d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('passC1_color_adjoint','passA_shade','passC_backward') if n in k})
except Exception as e: print('$f', 'ERR', e)
"; done
