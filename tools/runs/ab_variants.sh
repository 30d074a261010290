#!/bin/bash
# dev aid: A/B the experiment variants built into vhap_b200/variants/ (build_ext.py with VH_SO_OUT / VH_EXTRA_FLAGS) on the GPU box
cd "$(dirname "$0")/../.."
run() { timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], {n[:8]:v['ms_per_step'] for n,v in k.items()})"; }
run main
for so in vhap_b200/variants/*.so; do VHAP_B200_SO=$PWD/$so run $(basename $so .so); done
for e in $VH_AB_ENVS; do env $e bash -c "$(declare -f run); run $e"; done
