#!/bin/bash
# round-2 GPU call 21 (1 GPU): ncu --set full of every hot kernel in the final state (one launch each)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 700 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'k_passA|k_passB|k_passC1|k_passC2|k_fine|k_tex_fold3|k_blend_tc|k_pool_scatter|k_pool_count|k_aa_pairs|k_skin_fwd|k_skin_bwd|k_bin' -o gpurun_out/r02_final_all python tools/prof_step.py --steps 1 > gpurun_out/r21_ncu.log 2>&1
ncu -i gpurun_out/r02_final_all.ncu-rep --page raw --csv > gpurun_out/r02_ncu_final_all_raw.csv 2>/dev/null
tail -3 gpurun_out/r21_ncu.log; ls -la gpurun_out/r02_ncu_final_all_raw.csv
