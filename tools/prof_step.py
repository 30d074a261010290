#!/usr/bin/env python
"""dev aid for ncu: builds the bench workload, warms up, then runs `--steps` eager optimisation steps between cudaProfilerStart/Stop
(use with `ncu --profile-from-start off ...`).  Never a bench number."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--config", default="monocular")
    a = ap.parse_args()
    eng, batches, fg = bench.build_workload(a.size, a.batch, 2, 0, 1, a.config)
    resident = bench.stage_all(eng, batches).batches
    for i in range(4):
        eng.step(resident[i % 2])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(a.steps):
        eng.step(resident[i % 2])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", a.steps, "steps; fg", fg)


if __name__ == "__main__":
    main()
