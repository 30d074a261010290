#!/bin/bash
# round-2 GPU call 23: the reference arm (CPU port) on the GPU box's host cores after the oracle rasteriser was batched
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r23_bench_reference.json 2> gpurun_out/r23_bench_reference.err
tail -c 900 gpurun_out/r23_bench_reference.json
