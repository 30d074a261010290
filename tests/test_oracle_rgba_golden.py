"""oracle/render.py:render_rgba against the reference's own NVDiffRenderer.render_rgba run end to end
(tests/golden/make_rgba_golden.py: the reference's code with its nvdiffrast calls served by the oracle's restatements of those ops
and its random draws injected).  Values and gradients of everything in render_rgba that is not nvdiffrast: normals, texc detach
mask, SH shading twice, composite, background flip, the cluster disturbance, detach_by_indices, aa mask, output flips."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import render as RE
from tests.scene import get_model

G = dict(np.load(Path(__file__).parent / "golden" / "rgba_golden.npz"))
CASES = {"plain_imgbg": dict(bg="img", disturb=False, exc=False), "disturb_exc_imgbg": dict(bg="img", disturb=True, exc=True),
         "disturb_white": dict(bg=[1.0, 1.0, 1.0], disturb=True, exc=False)}


def close(a, b, tol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b).max()
    assert err <= tol * max(1.0, np.abs(b).max()), (what, err)


@pytest.mark.parametrize("name", list(CASES))
def test_render_rgba_matches_reference_code(name):
    c = CASES[name]
    m = get_model()
    T = lambda k: torch.tensor(G[k])
    faces = torch.as_tensor(m.faces.astype(np.int64))
    faces_uv = torch.as_tensor(m.faces_uv.astype(np.int64))
    verts = T("verts").requires_grad_(True)
    clip = T("clip").requires_grad_(True)
    tex = T("tex").requires_grad_(True)
    lights = T("lights").requires_grad_(True)
    fid2cid_padded = torch.cat([torch.zeros(1, dtype=torch.long), torch.as_tensor(G["fid2cid"]).long()])        # render_nvdiffrast.py:77-79
    dist = dict(w_fg=T("w_fg"), w_bg=T("w_bg"), u_rand=T("u_rand")) if c["disturb"] else None
    bg = T("bg_img") if c["bg"] == "img" else c["bg"]
    out = RE.render_rgba(T("rast"), T("rast_db"), verts, clip, faces, T("verts_uv"), faces_uv, tex, lights, bg, m.face_adjacency_opposite(),
                         fid2cid_padded, G["tex_exc"] if c["exc"] else None, G["bnd_exc"] if c["exc"] else None, dist)
    for k in ("rgba", "albedo", "normal", "diffuse", "diffuse_detach_normal"):
        close(out[k].detach().numpy(), G[f"{name}/{k}"], 2e-6, k)
    aa = ((out["rgba_pre"] - out["rgba"]) != 0).any(-1).numpy()
    assert np.array_equal(aa, G[f"{name}/aa"][..., 0].astype(bool))
    if c["disturb"]:
        assert np.array_equal(out["cid"].numpy(), G[f"{name}/cid"])
    (out["rgba"] * T("w_out")).sum().backward()
    close(lights.grad.numpy(), G[f"{name}/g_lights"], 2e-5, "g_lights")
    close(tex.grad.numpy(), G[f"{name}/g_tex"], 2e-5, "g_tex")
    close(verts.grad.numpy(), G[f"{name}/g_verts"], 2e-5, "g_verts")
    close(clip.grad.numpy(), G[f"{name}/g_clip"], 2e-5, "g_clip")
