#!/usr/bin/env python
"""Generate tests/golden/flame_golden.npz by calling the REFERENCE's own FlameHead.forward (vhap/model/flame.py:571-646) and
FlameTracker.forward_flame (vhap/model/tracker.py:213-235), imported unmodified, on bare instances (no __init__: the licensed FLAME
pickle, pytorch3d and the GPU are only needed by the constructors) that carry this repo's synthetic FLAME buffers.
Pins the composition around the (already pinned) lbs.py functions: offsets added before skinning, verts_cano = shaped + offsets,
translation after skinning, landmarks from the translated vertices, zero_centered_at_root_node, the per-timestep gathers and the
albedo expand of forward_flame -- values and gradients.

    PYTHONPATH=/root/reference python tests/golden/make_flame_golden.py
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
import vhap.model.tracker as RT                       # noqa: E402
from vhap.model.flame import FlameHead                # noqa: E402
from vhap.config import base as RB                    # noqa: E402
from vhap_b200.flame_model import FlameModelData      # noqa: E402
from oracle import lbs as OL                          # noqa: E402  (model_tensors: data conversion only)


def main():
    g = torch.Generator().manual_seed(41)
    m = FlameModelData.synthetic()
    mt = OL.model_tensors(m, torch.float32)
    V = m.v_template.shape[0]
    fh = object.__new__(FlameHead)
    torch.nn.Module.__init__(fh)
    fh.dtype = torch.float32
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "faces"):
        setattr(fh, k, mt[k])
    fh.parents = mt["parents"].long()
    fh.full_lmk_faces_idx = mt["lmk_faces_idx"].reshape(1, -1)
    fh.full_lmk_bary_coords = mt["lmk_bary"].reshape(1, -1, 3)
    n_t, T = 4, 8
    ts = np.array([3, 0, 3])
    rn = lambda *s, sc=1.0: torch.randn(*s, generator=g) * sc
    P = {"shape": rn(300, sc=0.5), "expr": rn(n_t, 100, sc=0.3), "rotation": rn(n_t, 3, sc=0.2), "neck_pose": rn(n_t, 3, sc=0.1), "jaw_pose": rn(n_t, 3, sc=0.1),
         "eyes_pose": rn(n_t, 6, sc=0.1), "translation": rn(n_t, 3, sc=0.05), "static_offset": rn(1, V, 3, sc=1e-3), "tex_extra": rn(3, T, T, sc=0.05)}
    tex_painted = torch.rand(1, 3, T, T, generator=g)
    out = {"ts": ts, "tex_painted": tex_painted.numpy()}
    for k, v in P.items():
        out["p_" + k] = v.numpy()
    # ---- FlameTracker.forward_flame on a bare tracker holding the bare FlameHead
    trk = object.__new__(RT.FlameTracker)
    trk.cfg = types.SimpleNamespace(model=RB.ModelConfig())
    trk.flame = fh
    trk.flame_tex_painted = lambda: tex_painted
    trk.dynamic_offset = None
    for k, v in P.items():
        setattr(trk, k, v.clone().requires_grad_(True))
    verts, verts_cano, lmks, albedos = trk.forward_flame(ts)
    wv, wc, wl, wa = rn(*verts.shape), rn(*verts_cano.shape), rn(*lmks.shape), rn(*albedos.shape)
    ((verts * wv).sum() + (verts_cano * wc).sum() + (lmks * wl).sum() + (albedos * wa).sum()).backward()
    out.update(verts=verts.detach().numpy(), verts_cano=verts_cano.detach().numpy(), lmks=lmks.detach().numpy(), albedos=albedos.detach().numpy(),
               w_verts=wv.numpy(), w_cano=wc.numpy(), w_lmks=wl.numpy(), w_alb=wa.numpy())
    for k in P:
        out["g_" + k] = getattr(trk, k).grad.numpy()
    # ---- FlameHead.forward with zero_centered_at_root_node (flame.py:617-619), no offsets
    B = len(ts)
    with torch.no_grad():
        vz, lz = fh(P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts], P["jaw_pose"][ts], P["eyes_pose"][ts],
                    P["translation"][ts], zero_centered_at_root_node=True)
    out.update(verts_zero_centered=vz.numpy(), lmks_zero_centered=lz.numpy())
    path = Path(__file__).with_name("flame_golden.npz")
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, verts.shape, lmks.shape, albedos.shape)


if __name__ == "__main__":
    main()
