#!/usr/bin/env python
"""Generate tests/golden/vis_golden.npz by running the REFERENCE's own NVDiffRenderer.render_rgba_vis
(vhap/util/render_nvdiffrast.py:486-567, imported unmodified) with its nvdiffrast calls (dr.rasterize / interpolate / texture /
antialias) served by the oracle's restatements of those ops (same arrangement as make_rgba_golden.py).  Pins, against the reference's
code, what render_rgba_vis adds to render_rgba: the camera chain inside the call, albedo = 1 without a texture, lighting_type
'constant' without lights, normal / diffuse filled with the (flipped) background outside the mesh, antialias of RGB and alpha.

    PYTHONPATH=/root/reference python tests/golden/make_vis_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).parent))
import make_rgba_golden as MR                          # noqa: E402  registers the nvdiffrast stub (oracle-served ops) + CPU patches
from oracle import energy as OE, lbs as OL, raster as RA   # noqa: E402
from tests.scene import make_scene                     # noqa: E402

sys.modules["nvdiffrast.torch"].rasterize = lambda ctx, pos, tri, resolution: RA.rasterize(pos, tri.long(), tuple(resolution))
NVDiffRenderer = MR.NVDiffRenderer

CASES = {"tex_sh_white": dict(tex=True, lights=True, bg=[1.0, 1.0, 1.0]), "tex_sh_img": dict(tex=True, lights=True, bg="img"),
         "bare_black": dict(tex=False, lights=False, bg=[0.0, 0.0, 0.0])}


def main():
    sc = make_scene(B=2, H=28, W=36, T=32, n_t=3, timesteps=[0, 2])
    m, model = sc["m"], sc["model"]
    MR.STATE["adj_opp"] = m.face_adjacency_opposite()
    B, H, W = sc["B"], sc["H"], sc["W"]
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in sc["params"].items()}
    ts = torch.as_tensor(sc["ts"]).long()
    with torch.no_grad():
        v64, _, _ = OL.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts], P["jaw_pose"][ts],
                                     P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
        K, RTm = OE.fill_cam_params(P, B, H, W)
    k4 = K.reshape(B, 4)
    K = torch.zeros(B, 3, 3, dtype=k4.dtype)          # the [N,4] branch of projection_from_intrinsics only broadcasts for N = 1 (:148)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = k4[:, 0], k4[:, 1], k4[:, 2], k4[:, 3], 1.0
    f32 = torch.float32
    verts = v64.to(f32)
    faces, faces_uv = model["faces"], model["faces_uv"]
    verts_uv = model["verts_uv"].to(f32).clone(); verts_uv[:, 1] = 1 - verts_uv[:, 1]
    tex = torch.tensor(sc["tex_painted"] + sc["params"]["tex_extra"], dtype=f32)
    lights = P["lights"].to(f32)
    bg_img = sc["rgb16"].to(f32).permute(0, 2, 3, 1).contiguous()
    out = dict(verts=verts.numpy(), RT=RTm.to(f32).numpy(), K=K.to(f32).numpy(), tex=tex.numpy(), lights=lights.numpy(), bg_img=bg_img.numpy(),
               verts_uv=verts_uv.numpy(), image_size=np.array([H, W]))
    for name, c in CASES.items():
        rnd = NVDiffRenderer(use_opengl=False, lighting_type="SH" if c["lights"] else "constant", lighting_space="world")
        bg = bg_img if c["bg"] == "img" else c["bg"]
        kw = dict(verts_uv=verts_uv, faces_uv=faces_uv, tex=tex[None].expand(B, -1, -1, -1)) if c["tex"] else {}
        res = rnd.render_rgba_vis(verts, faces, RTm.to(f32), K.to(f32), (H, W), bg, lights=lights[None] if c["lights"] else None, **kw)
        for k in ("albedo", "normal", "diffuse", "rgba"):
            out[f"{name}/{k}"] = res[k].detach().numpy()
        out["verts_clip"] = res["verts_clip"].numpy()          # the same for every case
        print(name, {k: tuple(res[k].shape) for k in res}, "alpha sum", float(res["rgba"][..., 3].sum()))
    path = Path(__file__).with_name("vis_golden.npz")
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size)


if __name__ == "__main__":
    main()
