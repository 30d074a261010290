"""Input staging: uint8 RGB targets normalised inside the kernels, and the ring of device slots refilled from
pinned host memory on a copy stream (vhap_b200/staging.py)."""
import numpy as np
import pytest
import torch

from tests.scene import make_scene

pytestmark = pytest.mark.gpu


def test_uint8_targets_are_to_tensor_exact():
    """background pixels of the rendered plane (background = target) are the target itself: with uint8 input they must equal
    F.to_tensor's u8 / 255 in fp32 bit for bit (video_dataset.py:256-260), and the energy must equal the one computed from the same
    values staged as floats up to the fp16 staging error of that path"""
    from vhap_b200.engine import Engine
    sc = make_scene(B=2, H=96, W=96, T=128, n_t=3, timesteps=[0, 2])
    e = Engine(sc["m"], sc["cfg"], 3, tex_painted=sc["tex_painted"])
    try:
        e.load_params(sc["params"])
        e.set_stage("rgb_global_tracking")
        e.inject_random(sc["w_fg"], sc["w_bg"], sc["u_rand"])
        u8 = (sc["rgb16"].to(torch.float32).permute(0, 2, 3, 1) * 255).round().clamp(0, 255).to(torch.uint8).contiguous()
        yy, xx = torch.meshgrid(torch.arange(96), torch.arange(96), indexing="ij")
        u8[1] = ((xx + 3 * yy)[..., None] + torch.tensor([0, 85, 170])).remainder(256).to(torch.uint8)     # frame 1: every byte value occurs
        b8 = e.stage_sample(u8, sc["lmk2d"], sc["ts"])
        assert b8.c.target_format == 1 and b8.target.dtype == torch.uint8
        planes = e.render_planes(b8, training=False)
        from tests.test_gpu_bench_configs import dilate
        bg = torch.as_tensor(~dilate((planes["cid"][..., 1] != 0).cpu().numpy()), device=e.dev)     # away from the silhouette (antialias blends)
        assert bg.float().mean() > 0.3
        ref = u8.to(torch.float32).div(255).to(e.dev)
        assert torch.equal(planes["rgba"][..., :3][bg], ref[bg])
        assert len(torch.unique(u8[1][bg[1].cpu()])) == 256
        l8 = e.energy(b8, backward=False, training=True).clone()
        bf = e.stage_sample(u8.to(torch.float32).div(255).permute(0, 3, 1, 2), sc["lmk2d"], sc["ts"])      # fp16-staged floats
        assert bf.c.target_format == 0
        lf = e.energy(bf, backward=False, training=True).clone()
        assert abs(float(l8[2]) - float(lf[2])) < 2e-3 * abs(float(lf[2]))
        with pytest.raises(ValueError):
            e.stage_sample(u8[..., :2], sc["lmk2d"], sc["ts"])
    finally:
        e.close()


def test_input_ring_refill_and_graph_replay():
    """slots keep their addresses: a step graph captured on slot j sees the data prefetched into it later"""
    from vhap_b200.engine import Engine
    from vhap_b200.staging import InputRing, pin_sample
    sc = make_scene(B=2, H=64, W=64, T=128, n_t=4, timesteps=[0, 1])
    e = Engine(sc["m"], sc["cfg"], 4, tex_painted=sc["tex_painted"])
    try:
        e.load_params(sc["params"])
        u8 = (sc["rgb16"].to(torch.float32).permute(0, 2, 3, 1) * 255).round().clamp(0, 255).to(torch.uint8).contiguous()
        sA = pin_sample(u8, sc["lmk2d"], [0, 1])
        sB = pin_sample(255 - u8, sc["lmk2d"], [2, 3])
        ring = InputRing(e, [sA, sA])
        assert ring.bytes_per_step() == u8.numel() + 2 * 68 * 3 * 4 + 2 * 4
        e.set_stage("rgb_global_tracking")
        ptr = ring.batches[1].target.data_ptr()
        ring.prefetch(1, sB)
        bt = ring.acquire(1)
        assert bt.target.data_ptr() == ptr
        lB = e.energy(bt, backward=False, training=False).clone()
        ring.release(1)
        torch.cuda.synchronize()
        assert torch.equal(bt.target.cpu(), sB["rgb"]) and bt.timesteps.cpu().tolist() == [2, 3]
        direct = e.stage_sample(sB["rgb"], sB["lmk2d"], sB["ts"].numpy())
        lD = e.energy(direct, backward=False, training=False).clone()
        assert torch.allclose(lB, lD, rtol=1e-6, atol=1e-7)
        with pytest.raises(ValueError):
            ring.prefetch(0, dict(sA, rgb=sA["rgb"][:, :32]))
    finally:
        e.close()


def test_graph_replay_follows_lr_scale():
    """with live step graphs the learning rates used to be frozen at capture time.  The scale now lives in device memory:
    a replay after `engine.lr_scale = s` must equal an eager step at that scale (ExponentialLR between epochs, tracker.py:1407-1412)"""
    from vhap_b200.engine import Engine
    sc = make_scene(B=2, H=64, W=64, T=128, n_t=3, timesteps=[0, 2])
    e = Engine(sc["m"], sc["cfg"], 3, tex_painted=sc["tex_painted"])
    try:
        batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
        res = {}
        for mode in ("eager", "graph", "graph_frozen"):
            e.load_params(sc["params"])
            e.set_stage("lmk_init_all", lr_scale=1.0)          # smooth stage: trajectories are comparable to fp32 rounding
            if mode != "eager":
                e.graph_begin([batch], pipelined=False)
            for i in range(3):
                if mode != "graph_frozen":
                    e.lr_scale = 0.5 ** i
                e.graph_step(0) if mode != "eager" else e.step(batch)
            if mode != "eager":
                e.graph_end()
            torch.cuda.synchronize()
            res[mode] = e.get_params()["expr"].copy()
        d_ok = np.abs(res["graph"] - res["eager"]).max()
        d_frozen = np.abs(res["graph_frozen"] - res["eager"]).max()
        assert d_ok < 1e-5, d_ok
        assert d_frozen > 1e-2, d_frozen                        # the schedule matters at this scale: the check above is not vacuous
    finally:
        e.close()
