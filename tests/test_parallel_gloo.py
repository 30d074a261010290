"""CPU test of the data-parallel host logic with gloo, world_size 2: forward-slab reduction semantics, frame sharding,
and that sharded oracle gradients summed over ranks (with batch-independent terms scaled by 1/world, the engine's
`shared_scale` rule) equal the single-process gradient."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vhap_b200.parallel import reduce_forward_slab, allreduce_sum, shard_frames
    # ---- slab reduction: sums for [0..2], max for [3], local for the rest
    local = torch.tensor([1.0 + rank, 10.0 + rank, 100.0 + rank, 0.5 + rank, 7.0 + rank, 0.5 + rank, 3.0 + rank, 0.0])
    out = torch.zeros(8)
    reduce_forward_slab(local, out)
    exp = torch.tensor([3.0, 21.0, 201.0, 1.5, 7.0 + rank, 0.5 + rank, 3.0 + rank, 0.0])
    ok1 = torch.allclose(out, exp)
    # ---- sharding covers all frames exactly once
    lo, hi = shard_frames(5, rank, world)
    cover = torch.zeros(5)
    cover[lo:hi] = 1
    allreduce_sum(cover)
    ok2 = bool((cover == 1).all())
    # ---- gradient equality on a small landmark-only oracle energy: mean over the GLOBAL batch + shape regulariser
    from oracle import energy as E, lbs as L
    from tests.scene import make_scene
    from vhap_b200.config import STAGES
    torch.set_num_threads(1)
    sc = make_scene(B=4, H=32, W=32, T=8, n_t=4, timesteps=[0, 1, 2, 3])
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc["params"].items()}
    ts = sc["ts"][lo:hi] if False else sc["ts"][rank * 2:(rank + 1) * 2]
    sample = dict(rgb=sc["rgb16"][rank * 2:(rank + 1) * 2].double(), lmk2d=torch.tensor(sc["lmk2d"][rank * 2:(rank + 1) * 2]), timestep_index=ts)
    verts, _, lm = L.flame_forward(sc["model"], P["shape"][None].expand(2, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts],
                                   P["jaw_pose"][ts], P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
    K, RT = E.fill_cam_params(P, 2, 32, 32)
    w = sc["cfg"].w
    # local share of the global mean (global batch 4 = 2 per rank) + batch-independent term scaled by 1/world
    e = w.landmark * E.lmk_energy(lm, sample["lmk2d"], K, RT, (32, 32)) * (2 / 4) + (1.0 / world) * w.reg_shape * (P["shape"] ** 2).mean()
    e.backward()
    g = torch.cat([P["shape"].grad.reshape(-1), P["rotation"].grad.reshape(-1), P["focal_length"].grad.reshape(-1)])
    allreduce_sum(g)
    # ---- uneven shards (5 frames over 2 ranks = 3 + 2): the global batch is the all-reduced sum of the local sizes, not B * world
    from vhap_b200.parallel import global_batch_size
    sc5 = make_scene(B=5, H=32, W=32, T=8, n_t=5, timesteps=[0, 1, 2, 3, 4])
    lo5, hi5 = shard_frames(5, rank, world)
    gB = global_batch_size(hi5 - lo5)
    P5 = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc5["params"].items()}
    ts5 = sc5["ts"][lo5:hi5]
    nb = hi5 - lo5
    _, _, lm5 = L.flame_forward(sc5["model"], P5["shape"][None].expand(nb, -1), P5["expr"][ts5], P5["rotation"][ts5], P5["neck_pose"][ts5],
                                P5["jaw_pose"][ts5], P5["eyes_pose"][ts5], P5["translation"][ts5], static_offset=P5["static_offset"])
    K5, RT5 = E.fill_cam_params(P5, nb, 32, 32)
    e5 = w.landmark * E.lmk_energy(lm5, torch.tensor(sc5["lmk2d"][lo5:hi5]), K5, RT5, (32, 32)) * (nb / gB) \
        + w.reg_expr * (P5["expr"][ts5] ** 2).sum() / (gB * P5["expr"].shape[1]) + (1.0 / world) * w.reg_shape * (P5["shape"] ** 2).mean()
    e5.backward()
    g5 = torch.cat([P5["shape"].grad.reshape(-1), P5["expr"].grad.reshape(-1), P5["rotation"].grad.reshape(-1)])
    allreduce_sum(g5)
    q.put((rank, ok1, ok2, g.numpy(), (gB, hi5 - lo5), g5.numpy()))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] and r[2] for r in res)
    # single-process reference
    from oracle import energy as E, lbs as L
    from tests.scene import make_scene
    sc = make_scene(B=4, H=32, W=32, T=8, n_t=4, timesteps=[0, 1, 2, 3])
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc["params"].items()}
    ts = sc["ts"]
    verts, _, lm = L.flame_forward(sc["model"], P["shape"][None].expand(4, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts],
                                   P["jaw_pose"][ts], P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
    K, RT = E.fill_cam_params(P, 4, 32, 32)
    w = sc["cfg"].w
    e = w.landmark * E.lmk_energy(lm, torch.tensor(sc["lmk2d"]), K, RT, (32, 32)) + w.reg_shape * (P["shape"] ** 2).mean()
    e.backward()
    ref = torch.cat([P["shape"].grad.reshape(-1), P["rotation"].grad.reshape(-1), P["focal_length"].grad.reshape(-1)]).numpy()
    for r in res:
        np.testing.assert_allclose(r[3], ref, rtol=1e-9, atol=1e-12)
    # uneven shards: global batch 5 on both ranks, shard sizes 3 + 2, summed gradients == single process
    assert sorted(r[4][1] for r in res) == [2, 3] and all(r[4][0] == 5 for r in res)
    sc5 = make_scene(B=5, H=32, W=32, T=8, n_t=5, timesteps=[0, 1, 2, 3, 4])
    P5 = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc5["params"].items()}
    ts5 = sc5["ts"]
    _, _, lm5 = L.flame_forward(sc5["model"], P5["shape"][None].expand(5, -1), P5["expr"][ts5], P5["rotation"][ts5], P5["neck_pose"][ts5],
                                P5["jaw_pose"][ts5], P5["eyes_pose"][ts5], P5["translation"][ts5], static_offset=P5["static_offset"])
    K5, RT5 = E.fill_cam_params(P5, 5, 32, 32)
    e5 = w.landmark * E.lmk_energy(lm5, torch.tensor(sc5["lmk2d"]), K5, RT5, (32, 32)) + w.reg_expr * (P5["expr"][ts5] ** 2).mean() \
        + w.reg_shape * (P5["shape"] ** 2).mean()
    e5.backward()
    ref5 = torch.cat([P5["shape"].grad.reshape(-1), P5["expr"].grad.reshape(-1), P5["rotation"].grad.reshape(-1)]).numpy()
    for r in res:
        np.testing.assert_allclose(r[5], ref5, rtol=1e-9, atol=1e-12)


def test_shard_frames_rejects_empty_shards():
    import pytest
    from vhap_b200.parallel import shard_frames
    with pytest.raises(ValueError):
        shard_frames(3, 0, 4)
    assert [shard_frames(5, r, 4) for r in range(4)] == [(0, 2), (2, 3), (3, 4), (4, 5)]
