#!/bin/bash
# round-2 GPU call 22 (1 GPU): the opt-in switches still work -- three-launch scan fallback (VHAP_B200_SCAN3) and programmatic dependent launch
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
VHAP_B200_SCAN3=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modular.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r22_pytest_scan3.log 2>&1
VHAP_B200_PDL=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r22_pytest_pdl.log 2>&1
tail -2 gpurun_out/r22_pytest_scan3.log; tail -2 gpurun_out/r22_pytest_pdl.log
