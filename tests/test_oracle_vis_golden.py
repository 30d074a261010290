"""oracle/render.py:render_rgba_vis against the reference's own NVDiffRenderer.render_rgba_vis (render_nvdiffrast.py:486-567) run
unmodified with its nvdiffrast calls served by the oracle's restatements (tests/golden/make_vis_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import camera as Cm, raster as RA, render as RE
from tests.scene import get_model

G = dict(np.load(Path(__file__).parent / "golden" / "vis_golden.npz"))
CASES = {"tex_sh_white": dict(tex=True, lights=True, bg=[1.0, 1.0, 1.0]), "tex_sh_img": dict(tex=True, lights=True, bg="img"),
         "bare_black": dict(tex=False, lights=False, bg=[0.0, 0.0, 0.0])}


@pytest.mark.parametrize("name", list(CASES))
def test_render_rgba_vis_matches_reference_code(name):
    c = CASES[name]
    m = get_model()
    T = lambda k: torch.tensor(G[k])
    faces = torch.as_tensor(m.faces.astype(np.int64)); faces_uv = torch.as_tensor(m.faces_uv.astype(np.int64))
    H, W = (int(v) for v in G["image_size"])
    verts = T("verts")
    clip = Cm.world_to_clip(verts, T("RT"), T("K"), (H, W))
    assert np.abs(clip.numpy() - G["verts_clip"]).max() <= 2e-6 * np.abs(G["verts_clip"]).max()
    clip = T("verts_clip")                                                  # identical ids on both sides
    rast, rast_db = RA.rasterize(clip, faces, (H, W))
    kw = dict(verts_uv=T("verts_uv"), faces_uv=faces_uv, tex=T("tex")) if c["tex"] else {}
    out = RE.render_rgba_vis(rast, rast_db, verts, clip, faces, m.face_adjacency_opposite(), T("bg_img") if c["bg"] == "img" else c["bg"],
                             lights=T("lights") if c["lights"] else None, **kw)
    for k in ("albedo", "normal", "diffuse", "rgba"):
        ref = G[f"{name}/{k}"]
        assert out[k].shape == ref.shape, k
        err = np.abs(out[k].numpy().astype(np.float64) - ref).max()
        assert err <= 2e-6 * max(1.0, np.abs(ref).max()), (k, err)
    if not c["tex"]:
        assert float(out["albedo"].min()) == 1.0 and float(out["diffuse"][out["rgba"][..., 3] == 1].min()) == 1.0
