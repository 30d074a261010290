"""oracle/camera.py, oracle/render.py (vertex normals, SH shading) against outputs AND autograd gradients of the reference's own
renderer methods (vhap/util/render_nvdiffrast.py imported unmodified with a stub for its absent `nvdiffrast` dependency,
tests/golden/make_render_golden.py).  Pins SURVEY.md 8(a) rows a6, a9, a11 of the oracle; the dr.* calls stay unpinned."""
from pathlib import Path

import numpy as np
import torch

from oracle import camera as Cm
from oracle import render as RE

G = dict(np.load(Path(__file__).parent / "golden" / "render_golden.npz"))
T = lambda k, **kw: torch.tensor(G[k], **kw)
H, W = (int(v) for v in G["image_size"])


def close(a, b, tol=2e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), np.abs(a - b).max()


def test_projection_both_intrinsic_forms():
    close(Cm.projection_from_intrinsics(T("K3"), (H, W)).numpy(), G["proj_K3"])
    close(Cm.projection_from_intrinsics(T("K4"), (H, W)).numpy(), G["proj_K4"])


def test_camera_chain_and_gradient():
    verts = T("verts").requires_grad_(True)
    cam = Cm.world_to_camera(verts, T("RT"))
    clip = Cm.camera_to_clip(cam, T("K3"), (H, W))
    close(cam.detach().numpy(), G["cam"]); close(clip.detach().numpy(), G["clip"])
    close(Cm.world_to_clip(verts, T("RT"), T("K4"), (H, W)).detach().numpy(), G["clip_w2c"])
    close(Cm.world_to_ndc(verts, T("RT"), T("K3"), (H, W), flip_y=True).detach().numpy(), G["ndc_flip"], 1e-5)
    close(Cm.world_to_ndc(verts, T("RT"), T("K3"), (H, W), flip_y=False).detach().numpy(), G["ndc"], 1e-5)
    (clip * T("w_clip")).sum().backward()
    close(verts.grad.numpy(), G["g_verts_from_clip"], 1e-5)
    # float64 restatement agrees with the fp32 reference to fp32 rounding
    close(Cm.world_to_clip(T("verts", dtype=torch.float64), T("RT", dtype=torch.float64), T("K3", dtype=torch.float64), (H, W)).numpy(), G["clip"], 1e-5)


def test_vertex_normals_fallback_and_gradient():
    verts = T("verts").requires_grad_(True)
    vn = RE.compute_v_normals(verts, T("faces"))
    close(vn.detach().numpy(), G["v_normals"], 1e-5)
    for v in G["lonely"]:                                   # vertices of no face: the (0,0,1) fallback (render_nvdiffrast.py:312)
        assert np.allclose(vn.detach().numpy()[:, int(v)], [0, 0, 1])
    (vn * T("w_vn")).sum().backward()
    close(verts.grad.numpy(), G["g_verts_from_vn"], 2e-4)   # 1/|n| amplifies fp32 rounding of the reference's own backward


def test_sh_shading_and_gradients():
    close(np.asarray(RE.SH_CONST, np.float32), G["sh_const"], 1e-7)
    normal = T("sh_normal").requires_grad_(True)
    lights = T("sh_lights").requires_grad_(True)
    d = RE.sh_shading(normal, lights[0])
    close(d.detach().numpy(), G["sh_diffuse"], 1e-5)
    (d * T("w_sh")).sum().backward()
    close(normal.grad.numpy(), G["g_normal"], 1e-5); close(lights.grad.numpy(), G["g_lights"], 1e-5)


def test_detach_by_indices_semantics():
    """the reference zeroes the gradient of the listed rows (render_nvdiffrast.py:349-352); the engine's vert_flags do the same"""
    g = G["dbi_grad"]
    idx = G["dbi_idx"]
    assert np.all(g[:, idx] == 0) and np.all(np.delete(g, idx, axis=1) == 1)
