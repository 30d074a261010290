"""oracle/energy.py:compute_energy's photometric glue against the reference's own FlameTracker.compute_photometric_energy
(tracker.py:391-478 + get_background_color :288-304 + the render_rgba wrapper :306-338), recorded with a fake renderer by
tests/golden/make_photo_golden.py: background per mode, v-flipped UVs, align_*_except region look-ups per stage, disturbance only
when a stage is given, and the L1 normalisation sum|gt - pred| / sum(alpha > 0) with its gradient.  (The renderer itself --
nvdiffrast -- stays unpinned.)"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import energy as OE
from oracle import raster as RA
from oracle import render as RE
from tests.scene import make_scene
from vhap_b200.config import STAGES

G = dict(np.load(Path(__file__).parent / "golden" / "photo_golden.npz"))


@pytest.fixture(scope="module")
def scene():
    B, _, H, W = G["gt_rgb"].shape
    return make_scene(B=B, H=H, W=W, T=8, n_t=3, timesteps=[0, 2][:B])


@pytest.mark.parametrize("mode", ["target", "white", "black"])
@pytest.mark.parametrize("stage_name", ["rgb_init_all", "rgb_global_tracking", None])
def test_photometric_glue(scene, monkeypatch, mode, stage_name):
    sc = scene
    key = f"{mode}/{stage_name}"
    rgba = torch.tensor(G["rgba"], dtype=torch.float64, requires_grad=True)
    calls = []

    def fake_render(rast, rast_db, verts, clip, faces, verts_uv, faces_uv, tex, lights, bg, adj, fid2cid=None, tex_exc=None, bnd_exc=None,
                    disturbance=None, **kw):
        calls.append(dict(verts_uv=verts_uv, tex=tex, lights=lights, bg=bg, tex_exc=tex_exc, bnd_exc=bnd_exc, disturbance=disturbance))
        return {"rgba": rgba, "diffuse_detach_normal": torch.zeros_like(rgba[..., :3])}

    monkeypatch.setattr(RE, "render_rgba", fake_render)
    monkeypatch.setattr(RA, "rasterize", lambda clip, faces, size: (None, None))
    cfg = sc["cfg"]
    monkeypatch.setattr(cfg.render, "background_train", mode)
    monkeypatch.setattr(cfg.render, "background_eval", mode)
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in sc["params"].items()}
    gt = torch.tensor(G["gt_rgb"], dtype=torch.float64)
    sample = dict(rgb=gt, lmk2d=torch.tensor(sc["lmk2d"]), timestep_index=sc["ts"])
    tp = torch.tensor(sc["tex_painted"], dtype=torch.float64)
    stage = STAGES[stage_name] if stage_name else None
    dist = {"w_fg": None, "w_bg": None, "u_rand": None}
    Et, log = OE.compute_energy(P, sample, stage, cfg, sc["m"], sc["model"], disturbance=dist, tex_painted=tp)
    c = calls[0]
    # ---- what the tracker hands to the renderer
    assert np.allclose(c["verts_uv"].numpy(), G[key + "/verts_uv_arg"], atol=1e-6)                       # v flipped (tracker.py:315-316)
    assert torch.equal(c["tex"], tp + P["tex_extra"])                                                    # get_albedo (tracker.py:247-258)
    if bool(G[key + "/bg_is_tensor"]):
        assert torch.is_tensor(c["bg"]) and np.allclose(c["bg"].numpy(), G[key + "/bg"], atol=1e-7)      # gt image as background, [B,H,W,3]
    else:
        assert list(c["bg"]) == list(G[key + "/bg"])
    if stage is None:
        assert c["tex_exc"] is None and c["bnd_exc"] is None and c["disturbance"] is None                # evaluation mode
    else:
        assert np.array_equal(np.asarray(c["tex_exc"]), G[key + "/tex_exc"]) and np.array_equal(np.asarray(c["bnd_exc"]), G[key + "/bnd_exc"])
        assert c["disturbance"] is dist and bool(G[key + "/enable_disturbance"])
    # ---- loss normalisation and its gradient
    ref = float(G[key + "/loss"])
    assert abs(float(log["photo"]) / cfg.w.photo - ref) <= 1e-6 * ref
    log["photo"].backward()
    g_ref = G[key + "/g_rgba"] * cfg.w.photo
    assert np.abs(rgba.grad.numpy() - g_ref).max() <= 1e-6 * np.abs(g_ref).max()
