#!/bin/bash
# round-2 GPU call 1: all GPU tests (incl. the new bench-configuration parity tests), baseline bench, ncu captures of the hot kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
nproc >> gpurun_out/r1_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r1_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/r1_bench512.json 2> gpurun_out/r1_bench512.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'k_passC1|k_passC2|k_tex_fold|k_blend_tc|k_fine|k_passB|k_passA' -o gpurun_out/r02_base_full python tools/prof_step.py --steps 1 > gpurun_out/r1_ncu.log 2>&1
ls -la gpurun_out | tail -20
tail -5 gpurun_out/r1_pytest.log
