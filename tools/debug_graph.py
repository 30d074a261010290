"""dev aid: characterise graph-replay vs eager differences (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from scene import make_scene
from vhap_b200.engine import Engine

sc = make_scene(B=3, H=96, W=96, T=256, n_t=4, timesteps=[1, 2, 1])
e = Engine(sc["m"], sc["cfg"], 4, tex_painted=sc["tex_painted"])
batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
nsteps = int(os.environ.get("NSTEPS", "4"))


def run(mode):
    e.load_params(sc["params"])
    e.set_stage("rgb_global_tracking")
    e.inject_random(None, None, None)
    e.global_step = 5
    traj = []
    if mode == "graph":
        e.graph_begin([batch])
    for i in range(nsteps):
        e.graph_step(0) if mode == "graph" else e.step(batch)
        torch.cuda.synchronize()
        traj.append({k: v.copy() for k, v in e.get_params().items()})
    if mode == "graph":
        e.graph_end()
    torch.cuda.synchronize()
    return traj


def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


ref = run("eager")
for trial in range(int(os.environ.get("TRIALS", "12"))):
    mode = "graph" if trial % 3 else "eager"
    t = run(mode)
    worst = {k: max(rel(t[i][k], ref[i][k]) for i in range(nsteps)) for k in ref[0]}
    bad = {k: round(v, 5) for k, v in worst.items() if v > 3e-3}
    print(trial, mode, "bad:", bad, flush=True)
    if bad:
        for i in range(nsteps):
            print("  step", i, {k: round(rel(t[i][k], ref[i][k]), 5) for k in bad})
        k = "static_offset"
        if k in bad:
            for i in range(nsteps):
                d = (t[i][k] - ref[i][k]).reshape(-1, 3)
                nz = np.abs(d).max(1) > 1e-7
                print("  step", i, "n verts differing", int(nz.sum()), "max|d|", float(np.abs(d).max()), "max|ref|", float(np.abs(ref[i][k]).max()),
                      "first ids", np.nonzero(nz)[0][:12].tolist())
