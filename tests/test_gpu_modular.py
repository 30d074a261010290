"""GPU parity tests of the modular entry points on IDENTICAL inputs (north_star: ids bit-exact, RGB / gradients 1e-4):
vhap_render_photometric (rasterise + render_rgba + photometric energy and its backward) from fp32 clip positions and
vertex normals that are also what the float64 oracle consumes; vhap_vertex_normals(_backward); vhap_project(_backward)."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.scene import make_scene, mip_offsets

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def setup():
    from vhap_b200.engine import Engine
    from oracle import lbs as L, energy as E, camera as Cm, render as RE
    sc = make_scene(B=2, H=128, W=128, T=256, n_t=3, timesteps=[0, 2])
    e = Engine(sc["m"], sc["cfg"], 3, tex_painted=sc["tex_painted"])
    e.load_params(sc["params"])
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in sc["params"].items()}
    ts, B, H, W = sc["ts"], sc["B"], sc["H"], sc["W"]
    with torch.no_grad():
        verts, _, _ = L.flame_forward(sc["model"], P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts],
                                      P["jaw_pose"][ts], P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
        K, RT = E.fill_cam_params(P, B, H, W)
        clip32 = Cm.world_to_clip(verts, RT, K, (H, W)).to(torch.float32)
        vn32 = RE.compute_v_normals(verts, sc["model"]["faces"]).to(torch.float32)
    yield e, sc, verts.to(torch.float32), clip32, vn32
    e.close()


@pytest.mark.parametrize("stage_name", ["rgb_init_all", None])
def test_render_photometric_identical_inputs(setup, stage_name):
    from oracle import raster as RA, render as RE
    from vhap_b200.config import STAGES
    e, sc, verts32, clip32, vn32 = setup
    m, model = sc["m"], sc["model"]
    B, H, W, T = sc["B"], sc["H"], sc["W"], sc["T"]
    dt = torch.float64
    stage = STAGES[stage_name] if stage_name else None
    w_photo, w_regdiff = sc["cfg"].w.photo, sc["cfg"].w.reg_diffuse
    # ---- oracle on the fp32 inputs
    clip = clip32.to(dt).requires_grad_(True)
    vn = vn32.to(dt).requires_grad_(True)
    lights = torch.tensor(sc["params"]["lights"], dtype=dt, requires_grad=True)
    tex = torch.tensor(sc["tex_painted"] + sc["params"]["tex_extra"]).to(torch.float32).to(dt).requires_grad_(True)
    ids, _ = RA.rasterize_ids(clip32.numpy(), m.faces, H, W)
    rast, rast_db = RA.shade_pass(clip, model["faces"], ids)
    vuv = model["verts_uv"].clone().to(torch.float32)
    vuv[:, 1] = 1 - vuv[:, 1]
    vuv = vuv.to(dt)
    bg = sc["rgb16"].to(dt).permute(0, 2, 3, 1)
    fid2cid = torch.as_tensor(np.concatenate([[0], m.fid2cid(sc["cfg"].tex_clusters)]).astype(np.int64))
    tex_exc = m.get_fid_by_region(list(stage.align_texture_except)) if stage else None
    bnd_exc = m.get_vid_by_region(list(stage.align_boundary_except)) if stage else None
    dist = dict(w_fg=sc["w_fg"], w_bg=sc["w_bg"], u_rand=sc["u_rand"]) if stage else None
    out = RE.render_rgba(rast, rast_db, None, clip, model["faces"], vuv, model["faces_uv"], tex, lights, bg, m.face_adjacency_opposite(),
                         fid2cid, tex_exc, bnd_exc, dist, v_normal=vn)
    pred = out["rgba"].permute(0, 3, 1, 2)
    n_fg = (pred[:, [3]].detach() > 0).sum()
    photo = w_photo * (sc["rgb16"].to(dt) - pred[:, :3]).abs().sum() / (3 * n_fg)
    total = photo
    if stage is not None:
        dd = out["diffuse_detach_normal"].permute(0, 3, 1, 2)
        regd = w_regdiff * (torch.relu(dd.max() - 1) + dd.var(dim=1).mean())
        total = total + regd
    total.backward()
    # ---- engine
    e.load_params(sc["params"])
    e.set_stage(stage)
    e.inject_random(sc["w_fg"], sc["w_bg"], sc["u_rand"])
    batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
    cs, cp = e._c_stage(True), e._c_params()
    dclip, dvn = clip32.to(e.dev).contiguous(), vn32.to(e.dev).contiguous()
    V = e.V
    g_clip = torch.zeros(B, V, 4, device=e.dev); g_vn = torch.zeros(B, V, 3, device=e.dev); g_l = torch.zeros(27, device=e.dev)
    gtex = e.L.vhap_tex_grad_ptr(e.ctx)
    e.L.vhap_set_want_planes(e.ctx, 1)
    e._ck(e.L.vhap_render_photometric(e.ctx, C.byref(cp), C.byref(batch.c), C.byref(cs), dclip.data_ptr(), dvn.data_ptr(), e.losses.data_ptr(),
                                      g_clip.data_ptr(), g_vn.data_ptr(), g_l.data_ptr(), gtex, e._stream()))
    e.L.vhap_set_want_planes(e.ctx, 0)
    planes = {}
    for which, name in ((0, "rgba"), (5, "cid")):
        t = torch.empty(B, H, W, 4, device=e.dev)
        e._ck(e.L.vhap_get_plane(e.ctx, which, t.data_ptr(), e._stream()))
        planes[name] = t.cpu().numpy()
    got = e.loss_dict()
    # texel gradient: fold (no regularisers) into a dense map
    import copy
    cfg_save = e.cfg
    e.cfg = copy.deepcopy(e.cfg); e.cfg.w.reg_tex_tv = None; e.cfg.w.reg_tex_res_clusters = None
    tex_g = e.texture_grad_dense().cpu().numpy()
    e.cfg = cfg_save
    torch.cuda.synchronize()
    # ---- checks
    ids_got = planes["cid"][..., 1].astype(np.int32)[:, ::-1]
    assert np.array_equal(ids_got, ids), "rasterised ids must be bit-exact"
    assert rel(planes["rgba"], out["rgba"].detach().numpy()) < 1e-4
    assert abs(got["n_fg"] - float(n_fg)) < 0.5
    assert abs(got["photo"] - float(photo)) < 1e-5 * float(photo)
    if stage is not None:
        assert abs(got["reg_diffuse"] - float(regd)) < 1e-4 * float(regd)
    errs = dict(clip=rel(g_clip.cpu().numpy()[..., [0, 1, 3]], clip.grad.numpy()[..., [0, 1, 3]]), vnorm=rel(g_vn.cpu().numpy(), vn.grad.numpy()),
                lights=rel(g_l.cpu().numpy().reshape(9, 3), lights.grad.numpy()), tex=rel(tex_g, tex.grad.numpy()))
    print("render_photometric gradient rel errors", stage_name, {k: float("%.3g" % v) for k, v in errs.items()})
    assert all(v < 1e-4 for v in errs.values()), errs


def test_vertex_normals_forward_backward(setup):
    from oracle import render as RE
    e, sc, verts32, clip32, vn32 = setup
    B, V = verts32.shape[:2]
    v = verts32.to(torch.float64).requires_grad_(True)
    n = RE.compute_v_normals(v, sc["model"]["faces"])
    g = torch.randn(B, V, 3, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    (n * g).sum().backward()
    dv = verts32.to(e.dev).contiguous()
    dn = torch.empty(B, V, 3, device=e.dev)
    e.reserve(B, 128, 128)
    e._ck(e.L.vhap_vertex_normals(e.ctx, dv.data_ptr(), B, dn.data_ptr(), e._stream()))
    assert rel(dn.cpu().numpy(), n.detach().numpy()) < 1e-4
    gv = torch.zeros(B, V, 3, device=e.dev)
    dg = g.to(torch.float32).to(e.dev).contiguous()
    e._ck(e.L.vhap_vertex_normals_backward(e.ctx, dv.data_ptr(), dg.data_ptr(), B, gv.data_ptr(), e._stream()))
    assert rel(gv.cpu().numpy(), v.grad.numpy()) < 1e-4


def test_project_forward_backward(setup):
    from oracle import energy as E, camera as Cm
    e, sc, verts32, clip32, vn32 = setup
    B, V = verts32.shape[:2]
    H, W = sc["H"], sc["W"]
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in sc["params"].items()}
    f = P["focal_length"].clone().requires_grad_(True)
    P["focal_length"] = f
    v = verts32.to(torch.float64).requires_grad_(True)
    K, RT = E.fill_cam_params(P, B, H, W)
    clip = Cm.world_to_clip(v, RT, K, (H, W))
    g = torch.randn(B, V, 4, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    (clip * g).sum().backward()
    batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
    cp = e._c_params()
    dv = verts32.to(e.dev).contiguous()
    dc = torch.empty(B, V, 4, device=e.dev)
    e._ck(e.L.vhap_project(e.ctx, C.byref(cp), C.byref(batch.c), dv.data_ptr(), dc.data_ptr(), e._stream()))
    assert rel(dc.cpu().numpy(), clip.detach().numpy()) < 1e-5
    gv = torch.zeros(B, V, 3, device=e.dev); gf = torch.zeros(1, device=e.dev)
    dg = g.to(torch.float32).to(e.dev).contiguous()
    e._ck(e.L.vhap_project_backward(e.ctx, C.byref(cp), C.byref(batch.c), dv.data_ptr(), dg.data_ptr(), gv.data_ptr(), gf.data_ptr(), e._stream()))
    assert rel(gv.cpu().numpy(), v.grad.numpy()) < 1e-5
    assert abs(gf.item() - f.grad.item()) < 1e-4 * abs(f.grad.item())


def test_fused_path_equals_modular_chain(setup):
    """The fused compute_energy backward must equal the chain of separately oracle-verified modular pieces evaluated on
    the engine's OWN fp32 geometry: render_photometric -> project_backward + vertex_normals_backward -> flame_backward."""
    import copy
    from vhap_b200.config import STAGES
    e, sc, _, _, _ = setup
    B, H, W, V = sc["B"], sc["H"], sc["W"], e.V
    cfg_save = e.cfg
    e.cfg = copy.deepcopy(cfg_save)
    w = e.cfg.w
    w.landmark = None; w.reg_tex_tv = None; w.reg_tex_res_clusters = None
    w.reg_shape = w.reg_expr = w.reg_neck = w.reg_jaw = w.reg_eyes = 0.0
    w.reg_offset = w.reg_offset_lap = w.reg_offset_rigid = None
    w.smooth_trans = w.smooth_rot = w.smooth_neck = w.smooth_jaw = w.smooth_eyes = w.smooth_expr = 0.0
    e.load_params(sc["params"])
    e.set_stage(STAGES["rgb_global_tracking"])
    e.inject_random(sc["w_fg"], sc["w_bg"], sc["u_rand"])
    batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
    e.zero_grad()
    e.energy(batch, backward=True, training=True)
    fused = {k: v.clone() for k, v in e.g.items()}
    geo = []
    for which in range(3):
        t = torch.empty(B, V, 4, device=e.dev)
        e._ck(e.L.vhap_get_geometry(e.ctx, which, t.data_ptr(), e._stream()))
        geo.append(t)
    verts3 = geo[0][..., :3].contiguous(); clip = geo[1].contiguous(); vn3 = geo[2][..., :3].contiguous()
    e.texture_grad_dense()         # drain the texel-gradient pyramid
    cs, cp = e._c_stage(True), e._c_params()
    g_clip = torch.zeros(B, V, 4, device=e.dev); g_vn = torch.zeros(B, V, 3, device=e.dev); g_l = torch.zeros(27, device=e.dev)
    e._ck(e.L.vhap_render_photometric(e.ctx, C.byref(cp), C.byref(batch.c), C.byref(cs), clip.data_ptr(), vn3.data_ptr(), e.losses.data_ptr(),
                                      g_clip.data_ptr(), g_vn.data_ptr(), g_l.data_ptr(), None, e._stream()))
    g_verts = torch.zeros(B, V, 3, device=e.dev); g_f = torch.zeros(1, device=e.dev)
    e._ck(e.L.vhap_project_backward(e.ctx, C.byref(cp), C.byref(batch.c), verts3.data_ptr(), g_clip.data_ptr(), g_verts.data_ptr(), g_f.data_ptr(), e._stream()))
    e._ck(e.L.vhap_vertex_normals_backward(e.ctx, verts3.data_ptr(), g_vn.data_ptr(), B, g_verts.data_ptr(), e._stream()))
    e.zero_grad()
    opt = {k: True for k in ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset")}
    opt["texture"] = False
    cg = e._c_grads(opt)
    # flame_forward must be re-run: render_photometric reused the ctx geometry buffers
    e._ck(e.L.vhap_flame_forward(e.ctx, C.byref(cp), C.byref(batch.c), None, None, None, e._stream()))
    e._ck(e.L.vhap_flame_backward(e.ctx, C.byref(cp), C.byref(batch.c), g_verts.data_ptr(), None, C.byref(cg), e._stream()))
    torch.cuda.synchronize()
    errs = {k: rel(fused[k].cpu().numpy(), e.g[k].cpu().numpy()) for k in ("shape", "expr", "rotation", "translation", "neck_pose", "jaw_pose",
                                                                            "eyes_pose", "static_offset")}
    errs["lights"] = rel(fused["lights"].cpu().numpy(), g_l.cpu().numpy())
    errs["focal_length"] = abs(fused["focal_length"].item() - g_f.item()) / abs(g_f.item())
    print("fused vs modular chain", {k: float("%.3g" % v) for k, v in errs.items()})
    e.cfg = cfg_save
    assert all(v < 1e-4 for v in errs.values()), errs
