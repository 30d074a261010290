#!/usr/bin/env python
"""Generate tests/golden/energy_golden.npz by calling the REFERENCE's own tracker methods (vhap/model/tracker.py, imported
unmodified): FlameTracker.compute_lmk_energy (:347-389) and FlameTracker.compute_regularization_energy (:480-605) with the
helpers they call (pose / joint / expression smoothness :616-648, joint prior :650-680, Laplacian :682-690, region weights
:607-614), on a bare instance (object.__new__, no __init__) that carries seeded parameters.  The module's absent third-party
imports (nvdiffrast, pytorch3d, matplotlib) are never reached by these methods and are replaced by empty stub modules; inside
this process `.cuda()` is the identity (no GPU here).  Region vertex sets, the uniform Laplacian and the residual-cluster UV
mask come from this repo's FlameModelData (the licensed FLAME masks are not available), so what is pinned is the ENERGY
ARITHMETIC: terms, weights, means, detaches -- values and autograd gradients.

    PYTHONPATH=/root/reference python tests/golden/make_energy_golden.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_stub("nvdiffrast").torch = _stub("nvdiffrast.torch", RasterizeCudaContext=lambda *a, **k: object(), RasterizeGLContext=lambda *a, **k: object())
_stub("pytorch3d"); _stub("pytorch3d.io", load_obj=None); _stub("pytorch3d.structures"); _stub("pytorch3d.structures.meshes", Meshes=None)
_stub("matplotlib", cm=None); _stub("matplotlib.pyplot")
torch.Tensor.cuda = lambda self, *a, **k: self
import vhap.model.tracker as RT                         # noqa: E402
from vhap.config import base as RB                      # noqa: E402
from vhap.util.render_nvdiffrast import NVDiffRenderer  # noqa: E402
from vhap_b200.flame_model import FlameModelData        # noqa: E402
from oracle import energy as OE                         # noqa: E402  (only for the dense Laplacian builder, data not arithmetic)

STAGE_CLASSES = {"rgb_global_tracking": RB.StageRgbGlobalTrackingConfig, "rgb_init_offset": RB.StageRgbInitOffsetConfig,
                 "lmk_init_all": RB.StageLmkInitAllConfig}
PARAMS = ("shape", "expr", "rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose", "lights", "tex_extra", "static_offset")


def main():
    g = torch.Generator().manual_seed(23)
    m = FlameModelData.synthetic()
    V = m.v_template.shape[0]
    n_t, B, T, H, W = 5, 3, 16, 40, 56
    ts = np.array([0, 2, 3])
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc)
    P = {"shape": rn(300, sc=0.5), "expr": rn(n_t, 100, sc=0.3), "rotation": rn(n_t, 3, sc=0.1), "translation": rn(n_t, 3, sc=0.02),
         "neck_pose": rn(n_t, 3, sc=0.05), "jaw_pose": rn(n_t, 3, sc=0.08), "eyes_pose": rn(n_t, 6, sc=0.05), "lights": rn(9, 3, sc=0.2),
         "tex_extra": rn(3, T, T, sc=0.05), "static_offset": rn(1, V, 3, sc=2e-4)}
    P["lights"][0] += float(np.sqrt(4 * np.pi))
    tex_painted = torch.rand(3, T, T, generator=g)
    verts_cano = rn(B, V, 3, sc=0.1)
    diffuse = (1.0 + rn(B, 3, H, W, sc=0.3))             # some values above 1 -> the relu(max - 1) branch is active
    uvmask = (torch.rand(T, T, generator=g) > 0.7).float()
    lmks = rn(B, 70, 3, sc=0.05); lmks[..., 2] -= 0.1
    lmk2d = torch.cat([torch.rand(B, 68, 2, generator=g) * torch.tensor([W, H]), torch.rand(B, 68, 1, generator=g)], -1)
    out = {"ts": ts, "n_t": n_t, "tex_painted": tex_painted.numpy(), "verts_cano": verts_cano.numpy(), "diffuse": diffuse.numpy(),
           "uvmask": uvmask.numpy(), "lmks": lmks.numpy(), "lmk2d": lmk2d.numpy(), "image_size": np.array([H, W])}
    for k, v in P.items():
        out["p_" + k] = v.numpy()
    lap = OE.laplacian_dense(m, torch.float32)

    for stage_name, cls in STAGE_CLASSES.items():
        stage = cls()
        trk = object.__new__(RT.FlameTracker)             # no __init__: datasets, FLAME pickle, GPU are not needed by the methods below
        trk.device = "cpu"
        trk.cfg = types.SimpleNamespace(w=RB.LossWeightConfig(), model=RB.ModelConfig(), data=types.SimpleNamespace(scale_factor=1.0, n_downsample_rgb=None))
        trk.n_timesteps = n_t
        for k in PARAMS:
            setattr(trk, k, P[k].clone().requires_grad_(True))
        trk.dynamic_offset = None
        trk.lights_uniform = torch.zeros(9, 3); trk.lights_uniform[0] = float(np.sqrt(4 * np.pi))
        trk.opt_dict = {k: (k in stage.optimizable_params) for k in ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset", "dynamic_offset")}
        trk.flame = types.SimpleNamespace(laplacian_matrix=lap, laplacian_matrix_negate_diag=None,
                                          mask=types.SimpleNamespace(get_vid_by_region=lambda regions: torch.as_tensor(m.get_vid_by_region(list(regions)))))
        trk.flame_uvmask = types.SimpleNamespace(get_uvmask_by_region=lambda regions: uvmask)
        trk.flame_tex_painted = lambda: tex_painted[None]
        trk.render = NVDiffRenderer(use_opengl=False, lighting_type="SH", lighting_space="world")
        log = trk.compute_regularization_energy({"diffuse_detach_normal": diffuse}, verts_cano + trk.static_offset, verts_cano + trk.static_offset,
                                                lmks, None, ts, stage_name)
        # landmark energy with the tracker's own uncalibrated camera (tracker.py:141-157)
        f = torch.tensor([1.5]) * max(H, W)
        sample = {"rgb": torch.zeros(B, 3, H, W), "lmk2d": lmk2d, "intrinsic": torch.stack([f, f, torch.tensor([0.5 * W]), torch.tensor([0.5 * H])], dim=1),
                  "extrinsic": torch.cat([torch.eye(3), torch.tensor([[0.0], [0.0], [-1.0]])], 1)[None].expand(B, -1, -1)}
        lm_in = lmks.clone().requires_grad_(True)
        e_lmk, _ = trk.compute_lmk_energy(sample, lm_in, getattr(stage, "disable_jawline_landmarks", False))
        total = sum(log.values()) + e_lmk
        total.backward()
        for k, v in log.items():
            out[f"{stage_name}/{k}"] = np.float64(v.item())
        out[f"{stage_name}/lmk_unweighted"] = np.float64(e_lmk.item())
        out[f"{stage_name}/g_lmks"] = lm_in.grad.numpy()
        for k in PARAMS:
            gr = getattr(trk, k).grad
            out[f"{stage_name}/g_{k}"] = (gr if gr is not None else torch.zeros_like(P[k])).numpy()
        print(stage_name, {k: round(float(v), 6) for k, v in log.items()}, "lmk", float(e_lmk))
    path = Path(__file__).with_name("energy_golden.npz")
    np.savez_compressed(path, **out)
    print(path)


if __name__ == "__main__":
    main()
