"""Shared small-scene builder for the tests (CPU and GPU)."""
import numpy as np
import torch

from vhap_b200 import synth
from vhap_b200.config import EngineConfig
from vhap_b200.flame_model import FlameModelData

_MODEL = None


def get_model():
    global _MODEL
    if _MODEL is None:
        _MODEL = FlameModelData.synthetic()
    return _MODEL


def make_scene(B=2, H=96, W=96, T=256, n_t=4, seed=0, timesteps=None, dtype=torch.float64, views=False):
    """Returns dict with model data, oracle model tensors, params (numpy f32), sample tensors, injected randomness.
    views=True: a calibrated multi-view rig (per-frame extrinsic [B,3,4] / intrinsic [B,3,3] in sc['RT'] / sc['K'], cfg.calibrated)."""
    from oracle import lbs as L, energy as E, camera as C
    m = get_model()
    model = L.model_tensors(m, dtype)
    p = synth.init_params(m, n_t, T, seed=seed)
    ts = np.asarray(timesteps if timesteps is not None else (np.arange(B) + 1) % n_t, np.int64)
    rgb16 = torch.tensor(synth.procedural_image(B, H, W, seed)).to(torch.float16)
    params64 = {k: torch.tensor(v, dtype=dtype) for k, v in p.items()}
    with torch.no_grad():
        _, _, lm = L.flame_forward(model, params64["shape"][None].expand(B, -1), params64["expr"][ts], params64["rotation"][ts],
                                   params64["neck_pose"][ts], params64["jaw_pose"][ts], params64["eyes_pose"][ts],
                                   params64["translation"][ts], static_offset=params64["static_offset"])
        RTv = Kv = None
        if views:
            centre = lm.mean((0, 1)).numpy()
            RTv, Kv = synth.ring_cameras(m, B, H, W, centre, seed=seed)
            K, RT = torch.tensor(Kv, dtype=dtype), torch.tensor(RTv, dtype=dtype)
        else:
            K, RT = E.fill_cam_params(params64, B, H, W)
        ndc = C.world_to_ndc(lm, RT, K, (H, W), flip_y=True)
    lmk2d = synth.landmarks_px(ndc.numpy(), H, W, seed)
    g = torch.Generator().manual_seed(seed + 11)
    w_fg = torch.rand(B, H, W, generator=g) < 0.5
    w_bg = torch.rand(B, H, W, generator=g) < 0.5
    u = torch.rand(B, H, W, generator=g, dtype=torch.float32)
    tex_painted = synth.procedural_texture(T, seed)
    cfg = EngineConfig(tex_resolution=T, calibrated=bool(views))
    return dict(RT=RTv, K=Kv, m=m, model=model, params=p, ts=ts, rgb16=rgb16, lmk2d=lmk2d, w_fg=w_fg, w_bg=w_bg, u_rand=u,
                tex_painted=tex_painted, cfg=cfg, B=B, H=H, W=W, T=T)


def mip_offsets(T):
    off, o, s = [], 0, T
    while True:
        off.append(o)
        o += s * s
        if s == 1:
            break
        s //= 2
    return off, o
