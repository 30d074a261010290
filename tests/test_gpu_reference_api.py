"""GPU tests of the drop-in classes (vhap_b200/reference_api.py): same call signatures / return dictionaries as the
reference's FlameHead and NVDiffRenderer, differentiable through autograd, checked against the oracle."""
import numpy as np
import pytest
import torch

from tests.scene import make_scene

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def l2rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_flame_head_module_matches_oracle():
    from oracle import lbs as L
    from vhap_b200.reference_api import B200FlameHead
    sc = make_scene(B=2, H=64, W=64, T=64, n_t=3, timesteps=[0, 2])
    m, model, ts = sc["m"], sc["model"], sc["ts"]
    head = B200FlameHead(300, 100, model=m)
    dev = head.eng.dev
    P64 = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc["params"].items()}
    P32 = {k: torch.tensor(v, device=dev, requires_grad=True) for k, v in sc["params"].items()}
    B = 2
    v64, c64, l64 = L.flame_forward(model, P64["shape"][None].expand(B, -1), P64["expr"][ts], P64["rotation"][ts], P64["neck_pose"][ts],
                                    P64["jaw_pose"][ts], P64["eyes_pose"][ts], P64["translation"][ts], static_offset=P64["static_offset"])
    tsd = torch.as_tensor(ts, device=dev)
    verts, cano, lmks = head(P32["shape"][None].expand(B, -1), P32["expr"][tsd], P32["rotation"][tsd], P32["neck_pose"][tsd], P32["jaw_pose"][tsd],
                             P32["eyes_pose"][tsd], P32["translation"][tsd], return_verts_cano=True, static_offset=P32["static_offset"])
    assert verts.shape == (B, 5143, 3) and lmks.shape == (B, 70, 3)
    assert rel(verts.detach().cpu().numpy(), v64.detach().numpy()) < 1e-5
    g = torch.Generator().manual_seed(0)
    wv = torch.randn(B, 5143, 3, generator=g, dtype=torch.float64); wl = torch.randn(B, 70, 3, generator=g, dtype=torch.float64)
    ((v64 * wv).sum() + (l64 * wl).sum()).backward()
    ((verts * wv.to(torch.float32).to(dev)).sum() + (lmks * wl.to(torch.float32).to(dev)).sum()).backward()
    for k in ("shape", "expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation", "static_offset"):
        assert rel(P32[k].grad.cpu().numpy(), P64[k].grad.numpy()) < 3e-4, k
    with pytest.raises(NotImplementedError):
        head(P32["shape"][None].expand(B, -1), P32["expr"][tsd], P32["rotation"][tsd], P32["neck_pose"][tsd], P32["jaw_pose"][tsd],
             P32["eyes_pose"][tsd], P32["translation"][tsd], dynamic_offset=torch.zeros(B, 5143, 3, device=dev))


def test_renderer_rasterize_and_render_rgba_autograd():
    from oracle import lbs as L, energy as E, camera as Cm, raster as RA, render as RE
    from vhap_b200.reference_api import B200Renderer
    sc = make_scene(B=2, H=96, W=96, T=256, n_t=3, timesteps=[0, 2])
    m, model, ts = sc["m"], sc["model"], sc["ts"]
    B, H, W, T = 2, 96, 96, 256
    rnd = B200Renderer(lighting_type="SH", model=m, tex_size=T)
    dev = rnd.eng.dev
    with pytest.raises(NotImplementedError):
        B200Renderer(lighting_type="front", model=m, tex_size=T)
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in sc["params"].items()}
    with torch.no_grad():
        v64, _, _ = L.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts], P["jaw_pose"][ts],
                                    P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
    verts32 = v64.to(torch.float32)
    K, RT = E.fill_cam_params(P, B, H, W)
    # ---- engine through the reference-style API
    verts = verts32.to(dev).requires_grad_(True)
    tex = torch.tensor(sc["tex_painted"] + sc["params"]["tex_extra"], device=dev)[None].requires_grad_(True)
    lights = torch.tensor(sc["params"]["lights"], device=dev).requires_grad_(True)
    Kd, RTd = K.to(torch.float32).to(dev), RT.to(torch.float32).to(dev)
    faces = torch.as_tensor(m.faces.astype(np.int64), device=dev)
    rast = rnd.rasterize(verts, faces, RTd, Kd, (H, W), False, True)
    assert set(rast) == {"rast_out", "rast_out_db", "verts", "verts_camera", "verts_clip"} and rast["rast_out"].shape == (B, H, W, 4)
    bg = sc["rgb16"].to(torch.float32).permute(0, 2, 3, 1).to(dev)
    vuv = torch.as_tensor(m.verts_uv, device=dev).clone(); vuv[:, 1] = 1 - vuv[:, 1]
    out = rnd.render_rgba(rast, verts, faces, vuv, torch.as_tensor(m.faces_uv.astype(np.int64), device=dev), tex, lights[None], bg)
    assert {"albedo", "normal", "diffuse", "diffuse_detach_normal", "rgba", "aa"} <= set(out)
    gt = sc["rgb16"].to(torch.float32).to(dev)
    loss = (gt - out["rgba"].permute(0, 3, 1, 2)[:, :3]).abs().sum() / 1000.0
    loss.backward()
    # ---- oracle on the same fp32 vertices
    vo = verts32.to(torch.float64).requires_grad_(True)
    texo = tex.detach().cpu()[0].to(torch.float64).requires_grad_(True)
    lo = lights.detach().cpu().to(torch.float64).requires_grad_(True)
    clip = Cm.world_to_clip(vo, RT, K, (H, W))
    r, rdb = RA.rasterize(clip, model["faces"], (H, W))
    vuv64 = model["verts_uv"].clone(); vuv64[:, 1] = 1 - vuv64[:, 1]
    oo = RE.render_rgba(r, rdb, vo, clip, model["faces"], vuv64.to(torch.float32).to(torch.float64), model["faces_uv"], texo, lo, bg.cpu().to(torch.float64),
                        m.face_adjacency_opposite())
    lo_ = (sc["rgb16"].to(torch.float64) - oo["rgba"].permute(0, 3, 1, 2)[:, :3]).abs().sum() / 1000.0
    lo_.backward()
    ids_ref = r[..., 3].detach().numpy()
    ids_got = rast["rast_out"][..., 3].cpu().numpy()
    assert (ids_ref != ids_got).sum() <= 4
    bad = (np.abs(out["rgba"].detach().cpu().numpy() - oo["rgba"].detach().numpy()).max(-1) > 2e-4)
    assert bad.mean() < 0.005
    # 'aa' = pixels changed by the antialias (render_nvdiffrast.py:466), 3 identical float channels
    aa_ref = ((oo["rgba_pre"] - oo["rgba"]) != 0).any(-1).numpy()
    aa_got = out["aa"].cpu().numpy()
    assert aa_got.shape == aa_ref.shape + (3,) and set(np.unique(aa_got)) <= {0.0, 1.0}
    assert (aa_got[..., 0].astype(bool) != aa_ref).mean() < 0.005 and aa_ref.any()
    assert abs(loss.item() - lo_.item()) < 2e-3 * abs(lo_.item())
    assert rel(lights.grad.cpu().numpy(), lo.grad.numpy()) < 2e-3
    assert l2rel(tex.grad.cpu().numpy()[0], texo.grad.numpy()) < 2e-2
    assert l2rel(verts.grad.cpu().numpy(), vo.grad.numpy()) < 0.2       # sliver-pixel conditioning w.r.t. fp32 clip positions, see test_gpu_parity


@pytest.mark.parametrize("case", ["tex_sh_img", "bare_black", "tex_sh_grey"])
def test_render_rgba_vis_matches_oracle(case):
    """B200Renderer.render_rgba_vis vs oracle/render.py:render_rgba_vis (pinned against the reference's code by
    tests/golden/make_vis_golden.py) on the same fp32 vertices: planes to 2e-4 outside the pixels whose face id differs"""
    from oracle import lbs as L, energy as E, camera as Cm, raster as RA, render as RE
    from vhap_b200.reference_api import B200Renderer
    sc = make_scene(B=2, H=96, W=128, T=128, n_t=3, timesteps=[0, 2])
    m, model = sc["m"], sc["model"]
    B, H, W = 2, 96, 128
    dev = "cuda:0"
    rnd = B200Renderer(model=m, tex_size=128)
    ts = torch.as_tensor(sc["ts"]).long()
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in sc["params"].items()}
    with torch.no_grad():
        v64, _, _ = L.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts], P["jaw_pose"][ts],
                                    P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
    verts32 = v64.to(torch.float32)
    K, RT = E.fill_cam_params(P, B, H, W)
    faces = torch.as_tensor(m.faces.astype(np.int64), device=dev)
    vuv = torch.as_tensor(m.verts_uv, device=dev).clone(); vuv[:, 1] = 1 - vuv[:, 1]
    tex = torch.tensor(sc["tex_painted"] + sc["params"]["tex_extra"], device=dev)
    lights = torch.tensor(sc["params"]["lights"], device=dev)
    bg_img = sc["rgb16"].to(torch.float32).permute(0, 2, 3, 1).contiguous()
    textured = case != "bare_black"
    bg = {"tex_sh_img": bg_img.to(dev), "bare_black": [0.0, 0.0, 0.0], "tex_sh_grey": [0.5, 0.25, 0.75]}[case]
    kw = dict(verts_uv=vuv, faces_uv=torch.as_tensor(m.faces_uv.astype(np.int64), device=dev), tex=tex[None], lights=lights[None]) if textured else {}
    out = rnd.render_rgba_vis(verts32.to(dev), faces, RT.to(torch.float32).to(dev), K.to(torch.float32).to(dev), (H, W), bg, **kw)
    assert set(out) == {"albedo", "normal", "diffuse", "rgba", "verts_clip"}
    vo = verts32.to(torch.float64)
    clip = Cm.world_to_clip(vo, RT, K, (H, W))
    r, rdb = RA.rasterize(clip, model["faces"], (H, W))
    vuv64 = model["verts_uv"].clone(); vuv64[:, 1] = 1 - vuv64[:, 1]
    okw = dict(verts_uv=vuv64.to(torch.float32).to(torch.float64), faces_uv=model["faces_uv"], tex=tex.cpu().to(torch.float64), lights=lights.cpu().to(torch.float64)) if textured else {}
    oo = RE.render_rgba_vis(r, rdb, vo, clip, model["faces"], m.face_adjacency_opposite(), bg_img.to(torch.float64) if case == "tex_sh_img" else bg, **okw)
    assert np.abs(out["verts_clip"].cpu().numpy() - clip.numpy()).max() < 1e-5
    fg = (r[..., 3] > 0).flip(1).numpy()
    assert fg.mean() > 0.1
    for k, tol in (("albedo", 2e-4), ("normal", 2e-4), ("diffuse", 2e-4), ("rgba", 2e-4)):
        got, ref = out[k].cpu().numpy(), oo[k].numpy()
        assert got.shape == ref.shape, k
        d = np.abs(got - ref).max(-1)
        if k == "albedo":
            d = d[fg]                     # outside the mesh nvdiffrast samples texel (0,0); the engine's albedo plane is only defined on the mesh
        assert (d > tol).mean() < 0.005, (k, (d > tol).mean(), d.max())
    if not textured:
        on = out["rgba"][..., 3].cpu().numpy() == 1.0
        assert np.abs(out["rgba"][..., :3].cpu().numpy()[on] - 1.0).max() < 1e-6            # albedo 1 x constant lighting
    with pytest.raises(NotImplementedError):
        rnd.render_rgba_vis(verts32.to(dev), faces, RT.to(torch.float32).to(dev), K.to(torch.float32).to(dev), (H, W), v_color=torch.ones(5143, 3))
