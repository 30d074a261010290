"""CPU tests: the device math headers (vhap_b200/csrc/*.cuh), compiled for the host by tests/hostcheck, against the
oracle's float64 autograd.  This pins the analytic forward/backward formulas without a GPU; the kernels that wrap
the same functions are tested on the GPU in tests/test_gpu_*.py."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import camera as C, energy as E, lbs as L, raster as RA, render as RE
from tests.hostcheck.build import build
from tests.scene import make_scene, mip_offsets


@pytest.fixture(scope="module")
def lib():
    return ctypes.CDLL(str(build()))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_pose_chain_forward_backward(lib):
    rng = np.random.default_rng(0)
    for trial in range(3):
        pose = (rng.normal(size=15) * 0.4).astype(np.float32)
        if trial == 1:
            pose[3:6] = 0          # zero rotation exercises the 1e-8 quirk
        J = (rng.normal(size=(5, 3)) * 0.1).astype(np.float32)
        gA = rng.normal(size=(5, 12)).astype(np.float32)
        gpf = rng.normal(size=36).astype(np.float32)
        A = np.zeros((5, 12), np.float32); pf = np.zeros(36, np.float32)
        g_pose = np.zeros(15, np.float32); g_J = np.zeros((5, 3), np.float32)
        lib.hc_pose(_ptr(pose), _ptr(J), _ptr(gA), _ptr(gpf), _ptr(A), _ptr(pf), _ptr(g_pose), _ptr(g_J))
        p = torch.tensor(pose, dtype=torch.float64, requires_grad=True)
        Jt = torch.tensor(J, dtype=torch.float64, requires_grad=True)
        rot = L.batch_rodrigues(p.view(-1, 3)).view(1, 5, 3, 3)
        _, At = L.batch_rigid_transform(rot, Jt[None], torch.tensor([-1, 0, 1, 1, 1]))
        A_ref = At[0, :, :3, :].reshape(5, 12)
        pf_ref = (rot[0, 1:] - torch.eye(3, dtype=torch.float64)).reshape(-1)
        loss = (A_ref * torch.tensor(gA, dtype=torch.float64)).sum() + (pf_ref * torch.tensor(gpf, dtype=torch.float64)).sum()
        loss.backward()
        np.testing.assert_allclose(A, A_ref.detach().numpy(), rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(pf, pf_ref.detach().numpy(), rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(g_J, Jt.grad.numpy(), rtol=1e-4, atol=1e-5)
        if trial != 1:      # at exactly zero rotation fp32 and fp64 differ by O(1) in the ill-conditioned 1e-8 term
            np.testing.assert_allclose(g_pose, p.grad.numpy(), rtol=2e-4, atol=2e-5)


def _render_setup(disturb, seed=0, B=2, H=64, W=64, T=128):
    sc = make_scene(B=B, H=H, W=W, T=T, seed=seed)
    m, model = sc["m"], sc["model"]
    dt = torch.float64
    P = {k: torch.tensor(v, dtype=dt) for k, v in sc["params"].items()}
    ts = sc["ts"]
    with torch.no_grad():
        verts, _, _ = L.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts],
                                      P["jaw_pose"][ts], P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
        K, RT = E.fill_cam_params(P, B, H, W)
        clip32 = C.world_to_clip(verts, RT, K, (H, W)).to(torch.float32)
        vn32 = RE.compute_v_normals(verts, model["faces"]).to(torch.float32)
    return sc, clip32, vn32


@pytest.mark.parametrize("disturb,size", [(0, (64, 64)), (1, (64, 64)), (1, (48, 72))])
def test_render_passes_match_oracle(lib, disturb, size):
    """(64, 64) takes the shift/mask pixel-unflatten path of power-of-two images, (48, 72) the general one"""
    sc, clip32, vn32 = _render_setup(disturb, H=size[0], W=size[1])
    m, model = sc["m"], sc["model"]
    B, H, W, T = sc["B"], sc["H"], sc["W"], sc["T"]
    dt = torch.float64
    V, F = m.v_template.shape[0], m.faces.shape[0]
    w_photo, w_regdiff = 30.0, 100.0
    tex_exc = m.get_fid_by_region(["hair", "boundary", "neck"])
    bnd_exc = m.get_vid_by_region(["hair", "bottomline"])
    fid2cid = np.concatenate([[0], m.fid2cid(sc["cfg"].tex_clusters)]).astype(np.uint8)
    adj = m.face_adjacency_opposite()

    # ---------------- oracle (float64 autograd on the float32-rounded inputs) ----------------
    clip = clip32.to(dt).requires_grad_(True)
    vn = vn32.to(dt).requires_grad_(True)
    lights = torch.tensor(sc["params"]["lights"], dtype=dt, requires_grad=True)
    tex = torch.tensor(sc["tex_painted"] + sc["params"]["tex_extra"]).to(torch.float32).to(dt).requires_grad_(True)
    ids, _ = RA.rasterize_ids(clip32.numpy(), m.faces, H, W)
    rast, rast_db = RA.shade_pass(clip, model["faces"], ids)
    verts_uv = model["verts_uv"].clone().to(torch.float32).to(dt)
    verts_uv[:, 1] = 1 - verts_uv[:, 1]
    verts_uv = verts_uv.to(torch.float32).to(dt)
    bg = sc["rgb16"].to(dt).permute(0, 2, 3, 1)
    dist = dict(w_fg=sc["w_fg"], w_bg=sc["w_bg"], u_rand=sc["u_rand"]) if disturb else None
    out = RE.render_rgba(rast, rast_db, None, clip, model["faces"], verts_uv, model["faces_uv"], tex, lights, bg, adj,
                         torch.as_tensor(fid2cid.astype(np.int64)), tex_exc, bnd_exc, dist, v_normal=vn)
    pred = out["rgba"].permute(0, 3, 1, 2)
    n_fg = (pred[:, [3]].detach() > 0).sum()
    abs_err = (sc["rgb16"].to(dt) - pred[:, :3]).abs().sum()
    photo = w_photo * abs_err / (3 * n_fg)
    dd = out["diffuse_detach_normal"].permute(0, 3, 1, 2)
    regd = w_regdiff * (torch.relu(dd.max() - 1) + dd.var(dim=1).mean())
    (photo + regd).backward()

    # ---------------- host-compiled device math ----------------
    faces4 = np.zeros((F, 4), np.int32); faces4[:, :3] = m.faces
    fuv4 = np.zeros((F, 4), np.int32); fuv4[:, :3] = m.faces_uv
    adj4 = np.zeros((F, 4), np.int32); adj4[:, :3] = adj
    vuv = verts_uv.to(torch.float32).numpy().copy()
    clipf = np.ascontiguousarray(clip32.numpy())
    vn4 = np.zeros((B, V, 4), np.float32); vn4[..., :3] = vn32.numpy()
    off, total = mip_offsets(T)
    mips = RE.build_mips(tex.detach().to(torch.float32).permute(1, 2, 0))
    pyr = np.zeros((total, 4), np.float32)
    for l, mm in enumerate(mips):
        pyr[off[l]:off[l] + mm.shape[0] ** 2, :3] = mm.reshape(-1, 3).numpy()
    mip_off = np.asarray(off + [0] * (16 - len(off)), np.int32)
    face_flags = np.zeros(F, np.uint8); face_flags[tex_exc] = 1
    vert_flags = np.zeros(V, np.uint8); vert_flags[bnd_exc] = 1
    target = np.zeros((B, H, W, 4), np.float16); target[..., :3] = sc["rgb16"].permute(0, 2, 3, 1).numpy()
    inj_w = (sc["w_fg"].numpy().astype(np.uint8) | (sc["w_bg"].numpy().astype(np.uint8) << 1))
    inj_u = sc["u_rand"].numpy()
    ids32 = np.ascontiguousarray(ids.astype(np.int32))
    lights32 = sc["params"]["lights"].astype(np.float32)
    pre = np.zeros((B, H, W, 4), np.float32); fin = np.zeros((B, H, W, 4), np.float32); sums = np.zeros(8, np.float32)
    g_clip = np.zeros((B, V, 4), np.float32); g_vn = np.zeros((B, V, 4), np.float32)
    g_tex = np.zeros((total, 4), np.float32); g_l = np.zeros(27, np.float32)
    lib.hc_render.argtypes = [ctypes.c_int] * 7 + [ctypes.c_void_p] * 13 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                                              ctypes.c_float, ctypes.c_float, ctypes.c_int] + [ctypes.c_void_p] * 7
    lib.hc_render(B, H, W, V, F, T, len(off) - 1, _ptr(faces4), _ptr(fuv4), _ptr(vuv), _ptr(clipf), _ptr(vn4), _ptr(lights32),
                  _ptr(pyr), _ptr(mip_off), _ptr(ids32), _ptr(face_flags), _ptr(vert_flags), _ptr(fid2cid), _ptr(adj4),
                  _ptr(target), disturb, _ptr(inj_w), _ptr(inj_u), w_photo, w_regdiff, 1,
                  _ptr(pre), _ptr(fin), _ptr(sums), _ptr(g_clip), _ptr(g_vn), _ptr(g_tex), _ptr(g_l))

    # forward parity (raster orientation on the C side -> flip)
    ref_rgba = out["rgba"].detach().flip(1).numpy()
    np.testing.assert_allclose(fin, ref_rgba, rtol=1e-4, atol=2e-5)
    assert abs(sums[3] - float(n_fg)) < 0.5
    np.testing.assert_allclose(sums[2], float(abs_err), rtol=1e-5)
    np.testing.assert_allclose(sums[0] / (B * H * W), float(dd.var(dim=1).mean()), rtol=1e-4)
    np.testing.assert_allclose(sums[4], float(dd.max()), rtol=1e-5)

    # gradient parity: 1e-4 relative to the gradient's own scale (north_star tolerance)
    def rel(a, b):
        return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert rel(g_clip[..., [0, 1, 3]], clip.grad.numpy()[..., [0, 1, 3]]) < 1e-4
    assert rel(g_vn[..., :3], vn.grad.numpy()) < 1e-4
    assert rel(g_l.reshape(9, 3), lights.grad.numpy()) < 1e-4
    # fold the texel-gradient pyramid to level 0 and compare with autograd's d/d tex
    g0 = np.zeros((T, T, 3), np.float64)
    for l in range(len(off)):
        s = T >> l
        gl = g_tex[off[l]:off[l] + s * s, :3].reshape(s, s, 3).astype(np.float64)
        g0 += np.repeat(np.repeat(gl, 1 << l, 0), 1 << l, 1) / (4.0 ** l)
    assert rel(g0, tex.grad.permute(1, 2, 0).numpy()) < 1e-4
    # sanity: the compared quantities are not trivially zero
    assert np.abs(g_clip).max() > 0 and np.abs(g_vn).max() > 0 and np.abs(g0).max() > 0 and np.abs(g_l).max() > 0
    assert (ids > 0).mean() > 0.05
