"""dev aid: device timeline of ONE CUDA-graph replay of the optimisation step (all streams of this library), from event-record nodes
captured into the graph (vhap_profile_enable(ctx, 2) + vhap_profile_timeline).  Shows which kernels are on the critical path of the
overlapped step; under torchrun (data parallel) the gaps are where the collectives / peer barriers sit.
Usage (GPU box): python tools/timeline.py [--size 512] [--batch 16] [--dp-texture auto|peer|shard|allreduce] > gpurun_out/timeline.txt"""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from vhap_b200.parallel import DataParallelStep

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--config", default="monocular")
ap.add_argument("--dp-texture", default="auto")
ap.add_argument("--dp-slab", default="peer")
ap.add_argument("--json", default=None)
ap.add_argument("--per-rank", default=None, help="prefix: every rank writes <prefix>.rank<r>.txt (rank 0 also prints)")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
eng, batches, fg = bench.build_workload(a.size, a.batch, 1, rank, world, a.config)
resident = bench.stage_all(eng, batches).batches
dp = DataParallelStep(eng, texture=a.dp_texture, slab=a.dp_slab)
for i in range(4):
    dp.step(resident[0])
torch.cuda.synchronize()
eng.L.vhap_profile_enable(eng.ctx, 2)
dp.graph_begin(resident)                  # two graphs (texture parity 0/1): slots [0, n/2) belong to the first
for i in range(10):
    eng.graph_step(0)
torch.cuda.synchronize()
eng.L.vhap_profile_enable(eng.ctx, -1)
N = 4096
kid = (C.c_int32 * N)(); t0 = (C.c_float * N)(); t1 = (C.c_float * N)()
n = eng.L.vhap_profile_timeline(eng.ctx, kid, t0, t1, N)
names = [eng.L.vhap_profile_kernel_name(k).decode() for k in range(eng.L.vhap_profile_kernel_count())]
rows = [(t0[i], t1[i], names[kid[i]]) for i in range(n)]
# per kernel id the slots of graph 0 come first: split by recording order
per = {}
for r in rows:
    per.setdefault(r[2], []).append(r)
g0, g1 = [], []
for k, v in per.items():
    h = len(v) // 2
    g0 += v[:h]; g1 += v[h:]
out = open(f"{a.per_rank}.rank{rank}.txt", "w") if a.per_rank else None
if rank == 0 or out:
    def print(*args, _p=print):            # noqa: A001  rank 0 -> stdout, every rank -> its file
        if rank == 0:
            _p(*args)
        if out:
            _p(*args, file=out)
    print(f"# world {world}, rank {rank}, texture mode {dp.texture_mode}, {a.config} {a.size} B={a.batch}")
    for name, g in (("graph parity A", g0), ("graph parity B", g1)):
        g.sort()
        if not g:
            continue
        base = g[0][0]
        end = max(r[1] for r in g)
        print(f"--- {name}: span {end - base:.4f} ms, {len(g)} launches")
        for s, e, k in g:
            print(f"{s - base:9.4f} {e - base:9.4f} {e - s:8.4f}  {k}")
    if a.json:
        json.dump({"rows": rows}, open(a.json, "w"))
eng.graph_end()
if world > 1:
    dist.destroy_process_group()
