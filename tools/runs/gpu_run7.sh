#!/bin/bash
# round-2 GPU call 7 (2 GPUs): graph timelines of the data-parallel step for the three texture-update variants
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for tex in peer shard allreduce; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29644 tools/timeline.py --dp-texture $tex > gpurun_out/r7_timeline_n2_$tex.txt 2> gpurun_out/r7_timeline_n2_$tex.err
done
timeout 300 python tools/timeline.py > gpurun_out/r7_timeline_n1.txt 2> gpurun_out/r7_timeline_n1.err
wc -l gpurun_out/r7_timeline_*.txt
