#!/usr/bin/env python
"""Generate tests/golden/e2e_golden.npz by running the REFERENCE's own FlameTracker.compute_energy (vhap/model/tracker.py:692-750)
END TO END on CPU: forward_flame -> FlameHead.forward -> lbs; compute_lmk_energy; rasterize_flame -> NVDiffRenderer.rasterize;
compute_photometric_energy -> NVDiffRenderer.render_rgba; compute_regularization_energy; the sum -- all the reference's code, imported
unmodified, on bare instances (tracker, FlameHead) that carry this repo's synthetic FLAME buffers and seeded parameters.  Only the
four nvdiffrast entry points (dr.rasterize / interpolate / texture / antialias) are served by the oracle's restatements of those ops,
and the disturbance draws are injected (see make_rgba_golden.py).  Stored: every log term, the total and the gradients w.r.t. ALL
parameters (incl. the focal length) for four stages -- the strongest statement about the oracle that can be made without nvdiffrast.

    PYTHONPATH=/root/reference python tests/golden/make_e2e_golden.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).parent))
import make_rgba_golden as MR                          # noqa: E402  registers the nvdiffrast stub (oracle-served ops) + CPU patches
from oracle import energy as OE, lbs as OL, raster as RA   # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


sys.modules["nvdiffrast.torch"].rasterize = lambda ctx, pos, tri, resolution: RA.rasterize(pos, tri.long(), tuple(resolution))
_stub("pytorch3d"); _stub("pytorch3d.io", load_obj=None); _stub("pytorch3d.structures"); _stub("pytorch3d.structures.meshes", Meshes=None)
_stub("matplotlib", cm=None); _stub("matplotlib.pyplot")
import vhap.model.tracker as RT                        # noqa: E402
from vhap.model.flame import FlameHead                 # noqa: E402
from vhap.config import base as RB                     # noqa: E402
from vhap.util.render_nvdiffrast import NVDiffRenderer  # noqa: E402
from tests.scene import make_scene                     # noqa: E402

PARAMS = ("shape", "expr", "rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose", "lights", "tex_extra", "static_offset", "focal_length")
STAGES = {"rgb_global_tracking": RB.StageRgbGlobalTrackingConfig, "rgb_init_all": RB.StageRgbInitAllConfig, "lmk_init_all": RB.StageLmkInitAllConfig}


def main():
    sc = make_scene(B=2, H=28, W=36, T=32, n_t=3, timesteps=[1, 2], dtype=torch.float32)
    m = sc["m"]
    MR.STATE["adj_opp"] = m.face_adjacency_opposite()
    mt = OL.model_tensors(m, torch.float32)
    B, H, W, T = sc["B"], sc["H"], sc["W"], sc["T"]
    fh = object.__new__(FlameHead)
    torch.nn.Module.__init__(fh)
    fh.dtype = torch.float32
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "faces"):
        setattr(fh, k, mt[k])
    fh.parents = mt["parents"].long()
    fh.full_lmk_faces_idx = mt["lmk_faces_idx"].reshape(1, -1)
    fh.full_lmk_bary_coords = mt["lmk_bary"].reshape(1, -1, 3)
    fh.textures_idx = mt["faces_uv"]; fh.verts_uvs = mt["verts_uv"]
    fh.laplacian_matrix = OE.laplacian_dense(m, torch.float32)
    fh.laplacian_matrix_negate_diag = None
    fh.mask = types.SimpleNamespace(get_fid_by_region=lambda r: torch.as_tensor(m.get_fid_by_region(list(r))),
                                    get_vid_by_region=lambda r: torch.as_tensor(m.get_vid_by_region(list(r))))
    tex_painted = torch.tensor(sc["tex_painted"], dtype=torch.float32)
    uvmask = torch.as_tensor(np.asarray(m.uvmask_res), dtype=torch.float32)
    if uvmask.shape[-1] != T:
        uvmask = uvmask[:: uvmask.shape[0] // T, :: uvmask.shape[1] // T]
    P0 = {k: torch.tensor(v, dtype=torch.float32) for k, v in sc["params"].items()}
    out = {"uvmask": uvmask.numpy(), "ts": np.asarray(sc["ts"]), "n_t": 3, "tex_painted": tex_painted.numpy(), "rgb": sc["rgb16"].to(torch.float32).numpy(),
           "lmk2d": np.asarray(sc["lmk2d"], np.float32), "w_fg": sc["w_fg"].numpy(), "w_bg": sc["w_bg"].numpy(), "u_rand": sc["u_rand"].numpy(),
           "image_size": np.array([H, W])}
    for k, v in P0.items():
        out["p_" + k] = v.numpy()
    pipeline = {name: cls() for name, cls in STAGES.items()}
    fid2cid = torch.as_tensor(m.fid2cid(sc["cfg"].tex_clusters))
    for stage in list(STAGES) + [None]:
        trk = object.__new__(RT.FlameTracker)
        trk.device = "cpu"
        trk.calibrated = False
        trk.cfg = types.SimpleNamespace(w=RB.LossWeightConfig(), model=RB.ModelConfig(), render=RB.RenderConfig(), pipeline=pipeline,
                                        data=types.SimpleNamespace(scale_factor=1.0, n_downsample_rgb=None))
        trk.n_timesteps = 3
        trk.image_size = (H, W)
        for k in PARAMS:
            v = P0[k].clone()
            setattr(trk, k, v.requires_grad_(True))
        trk.dynamic_offset = None
        trk.lights_uniform = torch.zeros(9, 3); trk.lights_uniform[0] = float(np.sqrt(4 * np.pi))
        trk.RT = torch.eye(3, 4); trk.RT[2, 3] = -1
        st = pipeline[stage] if stage else None
        trk.opt_dict = {k: (st is not None and k in st.optimizable_params) for k in ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset", "dynamic_offset")}
        trk.flame = fh
        trk.flame_uvmask = types.SimpleNamespace(get_uvmask_by_region=lambda regions: uvmask)
        trk.flame_tex_painted = lambda: tex_painted[None]
        trk.render = NVDiffRenderer(use_opengl=False, lighting_type="SH", lighting_space="world", disturb_rate_fg=0.5, disturb_rate_bg=0.5, fid2cid=fid2cid)
        sample = {"rgb": torch.tensor(out["rgb"]), "lmk2d": torch.tensor(out["lmk2d"]), "timestep_index": out["ts"]}
        trk.fill_cam_params_into_sample(sample)
        with MR.InjectedRandom(sc["w_fg"], sc["w_bg"], sc["u_rand"]):
            E, log, *_ = trk.compute_energy(sample, stage=stage)
        E.backward()
        key = str(stage)
        for k, v in log.items():
            out[f"{key}/{k}"] = np.float64(v.item())
        for k in PARAMS:
            gr = getattr(trk, k).grad
            out[f"{key}/g_{k}"] = (gr if gr is not None else torch.zeros_like(P0[k])).numpy()
        print(key, {k: round(float(v), 5) for k, v in log.items()})
    path = Path(__file__).with_name("e2e_golden.npz")
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size)


if __name__ == "__main__":
    main()
