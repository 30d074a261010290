"""GPU parity at the BENCHED configurations (BASELINE.json configs[1..4]): 2048^2 texture (k_tex_fold with 8 strips per row, a
12-level pyramid), 512^2 / 1024^2 power-of-two images, the 802x550 calibrated multi-view rig with per-view extrinsics and
intrinsics, the fused energy + backward through the C-ABI against the float64 oracle FROM THE PARAMETERS.

End to end the fp32 engine and the fp64 oracle disagree on a handful of discrete per-pixel decisions (which triangle covers a
sample that lies within 1e-7 of an edge, whether an antialias crossing exists): those pixels are found by comparing the
rendered planes, dilated by one pixel (their colour enters the neighbours through the antialias blend) and removed from the L1
sum on BOTH sides (vhap_set_loss_mask / sample['loss_mask']); everything else must agree to `TOL`.  The cluster disturbance is
switched off here -- one differing id shifts every later entry of a cluster pool, i.e. changes the sample drawn by thousands of
pixels; its forward and adjoint are held to 1e-4 on identical inputs by test_render_identical_inputs_bench_size below and
tests/test_gpu_modular.py.  Every per-parameter error is appended to gpurun_out/parity_r02.jsonl (committed as
profiles/r02_parity.txt)."""
import copy
import ctypes as C
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.scene import make_scene

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
# Tolerances.  On IDENTICAL inputs the engine holds 1e-4 at T <= 256 (tests/test_gpu_modular.py).  At the bench sizes the fp32 pixel math
# itself (the engine's, and nvdiffrast's) is conditioned worse: a texture coordinate carries ~1e-7 of rounding from the fp32 barycentrics,
# i.e. 2e-4 texel at T = 2048, which enters the bilinear weights and, multiplied by the texel pitch, the uv gradient; barycentrics of
# small triangles cancel.  Measured against the float64 oracle (profiles/r02_parity.txt): identical inputs 512^2 / T=2048: rgba 1e-4,
# d/d clip 6e-4; from the parameters: interior pixels 2e-4 .. 1e-3, all pixels 1e-3 .. 1.6e-2 in max norm (the tail is carried by the
# antialias adjoint of silhouette pairs whose edge is almost parallel to the pixel pair: d alpha / d position ~ 1 / (by - ay)^2), 1e-3 ..
# 8e-3 in L2, texture 5e-5 in L2 (probe in profiles/r02_parity.txt).  The bounds below sit 2x above the measured values; a wrong term or a
# dropped factor shows up at O(1e-1 .. 1) -- the round-1 bound was 0.2 in L2.
TOL = 3e-2                      # end to end from the parameters, fp32 engine vs fp64 oracle, masked: max norm relative to the largest component
TOL_L2 = 1.5e-2                 # ... and in relative L2
TOL_IDENTICAL = 2e-3            # identical fp32 inputs at 512^2 / T=2048
MAX_MASKED = 1e-3               # at most 0.1 % of the pixels may be masked (measured 0.05 .. 0.09 %)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def record(entry):
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        with open(out / "parity_r02.jsonl", "a") as f:
            f.write(json.dumps(entry) + "\n")
    except OSError:
        pass
    print("PARITY", json.dumps(entry))


def dilate(m):
    """3x3 dilation of a [B,H,W] bool array"""
    t = torch.as_tensor(m)[:, None].float()
    return (torch.nn.functional.max_pool2d(t, 3, 1, 1)[:, 0] > 0).numpy()


GROUPS = {"shape": ("shape",), "expr": ("expr",), "pose": ("rotation", "translation"), "joints": ("neck_pose", "jaw_pose", "eyes_pose"),
          "lights": ("lights",), "static_offset": ("static_offset",), "cam": ("focal_length",)}


def masked_e2e(sc, stage_name, label, tol=TOL, probe=False):
    """fused vhap_energy_forward/backward from the parameters vs oracle.compute_energy (float64), masked as described above"""
    from oracle import energy as E
    from vhap_b200.config import STAGES, opt_dict_for
    from vhap_b200.engine import Engine
    stage = STAGES[stage_name] if isinstance(stage_name, str) else stage_name
    stage_name = stage.name
    cfg = copy.deepcopy(sc["cfg"])
    cfg.render.disturb_rate_fg = cfg.render.disturb_rate_bg = None
    B, H, W, T = sc["B"], sc["H"], sc["W"], sc["T"]
    n_t = sc["params"]["expr"].shape[0]
    # ---- oracle forward
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc["params"].items()}
    if cfg.calibrated:
        P["focal_length"] = P["focal_length"].detach()
    sample = dict(rgb=sc["rgb16"].to(torch.float64), lmk2d=torch.tensor(sc["lmk2d"]), timestep_index=sc["ts"])
    if sc["RT"] is not None:
        sample["extrinsic"], sample["intrinsic"] = torch.tensor(sc["RT"]), torch.tensor(sc["K"])
    tp = torch.tensor(sc["tex_painted"], dtype=torch.float64)
    Et, log, aux = E.compute_energy(P, sample, stage, cfg, sc["m"], sc["model"], disturbance=None, tex_painted=tp, return_aux=True)
    # ---- engine forward + planes
    e = Engine(sc["m"], cfg, n_t, tex_painted=sc["tex_painted"])
    try:
        e.load_params(sc["params"])
        e.set_stage(stage)
        e.inject_random(None, None, None)
        batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"], RT=sc["RT"], K=sc["K"])
        planes = e.render_planes(batch, training=True)
        ids_ref = aux["rast"][..., 3].detach().numpy().astype(np.int32)[:, ::-1]              # image orientation
        ids_got = planes["cid"][..., 1].cpu().numpy().astype(np.int32)
        rg, rr = planes["rgba"].cpu().numpy(), aux["render"]["rgba"].detach().numpy()
        derr = np.abs(rg - rr).max(-1)
        bad = (ids_got != ids_ref) | (derr > 1e-3)
        # a differing pixel also enters its 4 neighbours' colours, but only through an antialias blend: mask the neighbours that the
        # antialias touched on either side (render_nvdiffrast.py:465-466), not the whole 3x3 block
        aa_o = ((aux["render"]["rgba"] - aux["render"]["rgba_pre"]) != 0).any(-1).numpy()
        aa_e = planes["albedo"][..., 3].cpu().numpy() > 0
        keep = ~(bad | (dilate(bad) & (aa_o | aa_e)))
        frac_bad, frac_masked = float(bad.mean()), float(1 - keep.mean())
        assert (ids_ref > 0).mean() > 0.1
        # values on the unmasked pixels (+ the planes of render_out: albedo / normal / diffuse, render_nvdiffrast.py:476-483)
        val_err = float(derr[keep].max())
        fgk = keep & (ids_ref > 0) & ~dilate(ids_ref == 0)
        plane_err = {}
        for k in ("albedo", "normal", "diffuse"):          # [99.9 % quantile, max]: the max sits on sliver triangles (fp32 barycentrics)
            dk = np.abs(planes[k][..., :3].cpu().numpy() - aux["render"][k].detach().numpy()).max(-1)[fgk]
            plane_err[k] = [float(np.quantile(dk, 0.999)), float(dk.max())]
        val_q = [float(np.quantile(derr[keep], q)) for q in (0.5, 0.999)]
        # ---- gradients with the mask on both sides
        pred = aux["render"]["rgba"].permute(0, 3, 1, 2)[:, :3]
        err = (sample["rgb"] - pred).abs() * torch.as_tensor(keep.copy()).to(torch.float64)[:, None]
        photo_m = cfg.w.photo * err.sum() / aux["n_fg"]
        do_probe = bool(probe or os.environ.get("VHAP_PARITY_PROBE"))
        (Et - log["photo"] + photo_m).backward(retain_graph=do_probe)
        e.set_loss_mask(keep)
        e.zero_grad()
        e.energy(batch, backward=True, training=True)
        tex_g = e.texture_grad_dense().cpu().numpy() if "texture" in stage.optimizable_params else None     # (also completes the loss vector: TV / residual)
        torch.cuda.synchronize()
        got = e.loss_dict()
        flag = C.c_int32(0)
        e.L.vhap_overflow_flag(e.ctx, C.byref(flag))
        assert flag.value == 0
        opt = opt_dict_for(stage)
        errs, errs2 = {}, {}
        for flag_name, names in GROUPS.items():
            if not opt[flag_name]:
                continue
            for n in names:
                if P[n].grad is None:
                    continue
                ref = P[n].grad.numpy().reshape(-1)
                g = e.g[n].cpu().numpy()
                errs[n], errs2[n] = rel(g, ref), rel_l2(g, ref)
        if tex_g is not None:
            errs["tex_extra"], errs2["tex_extra"] = rel(tex_g, P["tex_extra"].grad.numpy()), rel_l2(tex_g, P["tex_extra"].grad.numpy())
        loss_err = {k: abs(got[k] - float(v)) / max(abs(float(v)), 1e-3) for k, v in log.items() if k not in ("total", "photo")}
        loss_err["photo_masked"] = abs(got["photo"] - float(photo_m)) / float(photo_m)
        entry = dict(test=label, stage=stage_name, B=B, H=H, W=W, T=T, calibrated=bool(cfg.calibrated), fg_fraction=float((ids_ref > 0).mean()),
                     pixels_differing=int(bad.sum()), frac_differing=frac_bad, frac_masked=frac_masked, rgba_max_err_unmasked=val_err,
                     rgba_err_median_p999=val_q, plane_max_err_unmasked_fg=plane_err,
                     grad_rel_max={k: float("%.3g" % v) for k, v in errs.items()}, grad_rel_l2={k: float("%.3g" % v) for k, v in errs2.items()},
                     loss_rel={k: float("%.3g" % v) for k, v in loss_err.items()})
        if do_probe:
            entry["probe"] = _probe(e, sc, cfg, batch, P, sample, Et, log, aux, keep, ids_ref, opt)
        record(entry)
        assert frac_masked < MAX_MASKED, entry
        assert val_err <= 1e-3 and val_q[1] < 2e-4, entry
        assert all(v[0] < 1e-3 and v[1] < 3e-2 for v in plane_err.values()), entry
        # (reg_diffuse contains relu(max diffuse - 1): one pixel, possibly on a sliver triangle)
        assert all(v < (1e-3 if k == "reg_diffuse" else 2e-4) for k, v in loss_err.items()), entry
        bad_g = {k: (v, errs2[k]) for k, v in errs.items() if not (v < tol and errs2[k] < TOL_L2)}
        assert not bad_g, entry
    finally:
        e.set_loss_mask(None)
        e.close()


def _probe(e, sc, cfg, batch, P, sample, Et, log, aux, keep, ids_ref, opt):
    """diagnostic: which pixel class carries the fp32-vs-fp64 gradient difference (interior vs id-edge pixels, small triangles)"""
    B, H, W = ids_ref.shape
    edge = np.zeros_like(keep)
    edge[:, :, 1:] |= ids_ref[:, :, 1:] != ids_ref[:, :, :-1]
    edge[:, :, :-1] |= ids_ref[:, :, 1:] != ids_ref[:, :, :-1]
    edge[:, 1:] |= ids_ref[:, 1:] != ids_ref[:, :-1]
    edge[:, :-1] |= ids_ref[:, 1:] != ids_ref[:, :-1]
    clip = aux["clip"].detach().numpy()
    f = sc["m"].faces
    ndc = clip[..., :2] / clip[..., 3:4]
    px = ndc * np.array([W / 2.0, H / 2.0])
    a, b, c = px[:, f[:, 0]], px[:, f[:, 1]], px[:, f[:, 2]]
    area = 0.5 * np.abs((b[..., 0] - a[..., 0]) * (c[..., 1] - a[..., 1]) - (c[..., 0] - a[..., 0]) * (b[..., 1] - a[..., 1]))   # [B,F] px^2
    area_pix = np.where(ids_ref > 0, np.take_along_axis(area, np.maximum(ids_ref.reshape(B, -1) - 1, 0), 1).reshape(B, H, W), 1e9)
    variants = {"interior_only": keep & ~edge, "edge_only": keep & edge, "no_small_tris(<0.5px2)": keep & (area_pix > 0.5),
                "no_small_tris(<2px2)": keep & (area_pix > 2.0)}
    out = {}
    pred = aux["render"]["rgba"].permute(0, 3, 1, 2)[:, :3]
    for name, mk in variants.items():
        for p in P.values():
            p.grad = None
        err = (sample["rgb"] - pred).abs() * torch.as_tensor(mk.copy()).to(torch.float64)[:, None]
        (Et - log["photo"] + cfg.w.photo * err.sum() / aux["n_fg"]).backward(retain_graph=True)
        e.set_loss_mask(mk)
        e.zero_grad()
        e.energy(batch, backward=True, training=True)
        e.texture_grad_dense()
        torch.cuda.synchronize()
        r = {}
        for flag_name, names in GROUPS.items():
            if opt[flag_name]:
                for n in names:
                    if P[n].grad is not None:
                        r[n] = float("%.3g" % rel(e.g[n].cpu().numpy(), P[n].grad.numpy().reshape(-1)))
        out[name] = dict(pixels=int(mk.sum()), grad_rel_max=r)
    return out


# ---------------------------------------------------------------------------------------------- the benched configurations
def test_cfg1_monocular_512_T2048():
    """BASELINE configs[1] (monocular 512x512, 2048^2 texture), 4 of its 16 frames, stage rgb_global_tracking"""
    sc = make_scene(B=4, H=512, W=512, T=2048, n_t=5, timesteps=[1, 2, 3, 4])
    masked_e2e(sc, "rgb_global_tracking", "cfg1_512_T2048", probe=True)


def test_cfg2_nersemble_views_802x550_T2048():
    """BASELINE configs[2] (NeRSemble 802x550 calibrated multi-view, single timestep): 4 views with per-view extrinsic / intrinsic
    (tracker.py:141-147, nersemble_dataset.py:75-127), shared timestep, NeRSemble loss weights and stage (config/nersemble.py:36-57)"""
    from vhap_b200.config import nersemble_config, NERSEMBLE_STAGES
    sc = make_scene(B=4, H=802, W=550, T=2048, n_t=3, timesteps=[1, 1, 1, 1], views=True)
    sc["cfg"] = nersemble_config(tex_resolution=2048)
    masked_e2e(sc, NERSEMBLE_STAGES["rgb_global_tracking"], "cfg2_nersemble_802x550_T2048")


def test_cfg3_monocular_1024_T2048():
    """BASELINE configs[3] resolution (1024x1024), 2 frames per GPU of its 16"""
    sc = make_scene(B=2, H=1024, W=1024, T=2048, n_t=3, timesteps=[1, 2], seed=1)
    masked_e2e(sc, "rgb_global_tracking", "cfg3_1024_T2048")


def test_calibrated_cameras_landmark_stage():
    """per-frame RT / K through FLAME -> landmarks -> landmark energy and its backward (no rasteriser): smooth, held in max norm"""
    from oracle import energy as E
    from vhap_b200.config import STAGES, opt_dict_for, nersemble_config
    from vhap_b200.engine import Engine
    sc = make_scene(B=6, H=802, W=550, T=64, n_t=4, timesteps=[1, 1, 1, 2, 2, 2], views=True)
    cfg = nersemble_config(tex_resolution=64)
    stage = STAGES["lmk_init_all"]
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=(k != "focal_length")) for k, v in sc["params"].items()}
    sample = dict(rgb=sc["rgb16"].to(torch.float64), lmk2d=torch.tensor(sc["lmk2d"]), timestep_index=sc["ts"],
                  extrinsic=torch.tensor(sc["RT"]), intrinsic=torch.tensor(sc["K"]))
    Et, log = E.compute_energy(P, sample, stage, cfg, sc["m"], sc["model"])
    Et.backward()
    e = Engine(sc["m"], cfg, 4, tex_painted=sc["tex_painted"])
    try:
        e.load_params(sc["params"])
        e.set_stage(stage)
        batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"], RT=sc["RT"], K=sc["K"])
        e.zero_grad()
        e.energy(batch, backward=True, training=True)
        got = e.loss_dict()
        for k, v in log.items():
            if k != "total":
                assert abs(got[k] - float(v)) <= 1e-4 * max(abs(float(v)), 1e-3), (k, got[k], float(v))
        errs = {}
        for n in ("shape", "expr", "rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose"):
            errs[n] = rel(e.g[n].cpu().numpy(), P[n].grad.numpy().reshape(-1))
        record(dict(test="calibrated_lmk_stage", grad_rel_max={k: float("%.3g" % v) for k, v in errs.items()}))
        assert all(v < 1e-3 for v in errs.values()), errs      # cameras at unit distance: fp32 view transforms of O(1) coordinates
        assert float(e.g["focal_length"].abs().sum()) == 0.0          # calibrated: no focal length to optimise (tracker.py:1330-1339)
    finally:
        e.close()


# ---------------------------------------------------------------------------------------------- identical inputs at the bench size
def test_render_identical_inputs_bench_size():
    """vhap_render_photometric at 512^2 / T=2048 WITH the cluster disturbance, on the same fp32 clip positions and vertex normals the
    float64 oracle consumes: ids bit-exact, RGBA and all gradients 1e-4 (the T<=256 version lives in tests/test_gpu_modular.py)"""
    from oracle import lbs as L, energy as E, camera as Cm, raster as RA, render as RE
    from vhap_b200.config import STAGES
    from vhap_b200.engine import Engine
    sc = make_scene(B=2, H=512, W=512, T=2048, n_t=3, timesteps=[0, 2])
    m, model = sc["m"], sc["model"]
    B, H, W, T = sc["B"], sc["H"], sc["W"], sc["T"]
    dt = torch.float64
    stage = STAGES["rgb_global_tracking"]
    P = {k: torch.tensor(v, dtype=dt) for k, v in sc["params"].items()}
    ts = sc["ts"]
    with torch.no_grad():
        verts, _, _ = L.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts],
                                      P["jaw_pose"][ts], P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
        K, RT = E.fill_cam_params(P, B, H, W)
        clip32 = Cm.world_to_clip(verts, RT, K, (H, W)).to(torch.float32)
        vn32 = RE.compute_v_normals(verts, model["faces"]).to(torch.float32)
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
    bnd_exc = m.get_vid_by_region(list(stage.align_boundary_except))
    tex_exc = m.get_fid_by_region(list(stage.align_texture_except))
    dist = dict(w_fg=sc["w_fg"], w_bg=sc["w_bg"], u_rand=sc["u_rand"])
    out = RE.render_rgba(rast, rast_db, None, clip, model["faces"], vuv, model["faces_uv"], tex, lights, bg, m.face_adjacency_opposite(),
                         fid2cid, tex_exc, bnd_exc, dist, v_normal=vn)
    pred = out["rgba"].permute(0, 3, 1, 2)
    n_fg = (pred[:, [3]].detach() > 0).sum()
    w = sc["cfg"].w
    photo = w.photo * (sc["rgb16"].to(dt) - pred[:, :3]).abs().sum() / (3 * n_fg)
    dd = out["diffuse_detach_normal"].permute(0, 3, 1, 2)
    regd = w.reg_diffuse * (torch.relu(dd.max() - 1) + dd.var(dim=1).mean())
    (photo + regd).backward()
    e = Engine(m, sc["cfg"], 3, tex_painted=sc["tex_painted"])
    try:
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
        cfg_save = e.cfg
        e.cfg = copy.deepcopy(e.cfg); e.cfg.w.reg_tex_tv = None; e.cfg.w.reg_tex_res_clusters = None
        tex_g = e.texture_grad_dense().cpu().numpy()
        e.cfg = cfg_save
        torch.cuda.synchronize()
        ids_got = planes["cid"][..., 1].astype(np.int32)[:, ::-1]
        assert np.array_equal(ids_got, ids), "rasterised ids must be bit-exact"
        errs = dict(rgba=rel(planes["rgba"], out["rgba"].detach().numpy()),
                    clip=rel(g_clip.cpu().numpy()[..., [0, 1, 3]], clip.grad.numpy()[..., [0, 1, 3]]), vnorm=rel(g_vn.cpu().numpy(), vn.grad.numpy()),
                    lights=rel(g_l.cpu().numpy().reshape(9, 3), lights.grad.numpy()), tex=rel(tex_g, tex.grad.numpy()),
                    photo=abs(got["photo"] - float(photo)) / float(photo), reg_diffuse=abs(got["reg_diffuse"] - float(regd)) / float(regd))
        record(dict(test="render_identical_inputs_512_T2048_disturbed", rel_max={k: float("%.3g" % v) for k, v in errs.items()},
                    n_fg=int(n_fg), ids_bit_exact=True))
        assert abs(got["n_fg"] - float(n_fg)) < 0.5
        assert all(v < TOL_IDENTICAL for v in errs.values()), errs
    finally:
        e.close()


# ---------------------------------------------------------------------------------------------- texture update at T = 2048
def test_tex_fold_adam_T2048_matches_torch_adam():
    """k_tex_fold with 8 strips per texture row (T = 2048; the other tests use T <= 256, one strip): dense TV + residual gradient
    vs the oracle (1e-4), then two fused fold + Adam + pyramid-rebuild steps vs torch.optim.Adam fed with the engine's own dense
    gradient (same inputs: 2e-6), and the fold-emitted levels 0 / 1 + mips vs an explicit rebuild of the pyramid."""
    from oracle import energy as E
    from vhap_b200.config import STAGES
    from vhap_b200.engine import Engine
    sc = make_scene(B=1, H=64, W=64, T=2048, n_t=2, timesteps=[1])
    cfg = copy.deepcopy(sc["cfg"])
    cfg.w.photo = None
    cfg.w.landmark = None
    stage = STAGES["rgb_global_tracking"]
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc["params"].items()}
    sample = dict(rgb=sc["rgb16"].to(torch.float64), lmk2d=torch.tensor(sc["lmk2d"]), timestep_index=sc["ts"])
    tp = torch.tensor(sc["tex_painted"], dtype=torch.float64)
    Et, log = E.compute_energy(P, sample, stage, cfg, sc["m"], sc["model"], tex_painted=tp)
    Et.backward()
    e = Engine(sc["m"], cfg, 2, tex_painted=sc["tex_painted"])
    try:
        e.load_params(sc["params"])
        e.set_stage(stage)
        batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
        e.zero_grad()
        e.energy(batch, backward=True, training=True)
        g1 = e.texture_grad_dense().clone()
        got = e.loss_dict()
        err_g = rel(g1.cpu().numpy(), P["tex_extra"].grad.numpy())
        assert abs(got["reg_tex_tv"] - float(log["reg_tex_tv"])) < 1e-5 * float(log["reg_tex_tv"])
        assert abs(got["reg_tex_res_clusters"] - float(log["reg_tex_res_clusters"])) < 1e-5 * float(log["reg_tex_res_clusters"])
        assert err_g < 1e-4, err_g
        ref = e.tex_extra.clone().view(3, 2048, 2048).requires_grad_(True)
        optim = torch.optim.Adam([ref], lr=e._lr("tex"))
        errs = []
        for step in range(2):
            g = e.texture_grad_dense(with_losses=False).clone()       # pyramid is empty (no photometric term): TV + residual of the current texture
            ref.grad = g.view(3, 2048, 2048).clone()
            optim.step()
            e.step_count += 1
            e.tex_update(None)
            torch.cuda.synchronize()
            errs.append(float((e.tex_extra.view(3, 2048, 2048) - ref.detach()).abs().max() / e._lr("tex")))
        # the pyramid written by the fold kernel (levels 0, 1) + k_mip_down vs a rebuild from tex_extra: render through both
        e.cfg.w.photo = 30.0
        pl_a = e.render_planes(batch, training=False)["rgba"].clone()
        e.rebuild_texture()
        pl_b = e.render_planes(batch, training=False)["rgba"].clone()
        e.cfg.w.photo = None
        pyr = float((pl_a - pl_b).abs().max())
        record(dict(test="tex_fold_adam_T2048", dense_grad_rel_max=float("%.3g" % err_g), adam_step_err_over_lr=[float("%.3g" % x) for x in errs],
                    pyramid_render_diff=pyr))
        assert max(errs) < 2e-4, errs          # |delta| < 2e-4 lr: Adam's m / (sqrt(v) + eps) is O(1), fp32 rounding of it is 1e-7
        assert pyr < 1e-6, pyr
    finally:
        e.close()


# ---------------------------------------------------------------------------------------------- every regulariser gradient
@pytest.mark.parametrize("stage_name", ["rgb_global_tracking", "rgb_init_offset"])
def test_regulariser_gradients_photometric_off(stage_name):
    """photometric and landmark terms off: what is left is compute_regularization_energy (tracker.py:480-690) alone -- texture TV and
    residual, offset L1 / Laplacian / rigidity with region relaxation, temporal smoothness, joint / expression / shape priors --
    every gradient held to 1e-4 (T = 512: two strips per texture row)"""
    from oracle import energy as E
    from vhap_b200.config import STAGES, opt_dict_for
    from vhap_b200.engine import Engine
    sc = make_scene(B=3, H=64, W=64, T=512, n_t=4, timesteps=[1, 2, 3])
    cfg = copy.deepcopy(sc["cfg"])
    cfg.w.photo = None
    cfg.w.landmark = None
    cfg.w.reg_light = 1e1               # off by default in the reference (base.py:158); exercised here
    stage = STAGES[stage_name]
    p = dict(sc["params"])
    rng = np.random.default_rng(4)
    p["static_offset"] = (rng.normal(0, 2e-3, p["static_offset"].shape)).astype(np.float32)      # well away from the |.| kink at 0
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    sample = dict(rgb=sc["rgb16"].to(torch.float64), lmk2d=torch.tensor(sc["lmk2d"]), timestep_index=sc["ts"])
    tp = torch.tensor(sc["tex_painted"], dtype=torch.float64)
    Et, log = E.compute_energy(P, sample, stage, cfg, sc["m"], sc["model"], tex_painted=tp)
    Et.backward()
    e = Engine(sc["m"], cfg, 4, tex_painted=sc["tex_painted"])
    try:
        e.load_params(p)
        e.set_stage(stage)
        batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
        e.zero_grad()
        e.energy(batch, backward=True, training=True)
        tex_g = e.texture_grad_dense().cpu().numpy()
        torch.cuda.synchronize()
        got = e.loss_dict()
        assert set(log) >= {"reg_tex_tv", "reg_tex_res_clusters", "reg_offset", "reg_offset_lap", "reg_offset_rigid", "reg_joint", "reg_expr", "reg_shape", "reg_light"}
        lerr = {k: abs(got[k] - float(v)) / max(abs(float(v)), 1e-6) for k, v in log.items() if k != "total"}
        errs = {}
        opt = opt_dict_for(stage)
        for flag_name, names in GROUPS.items():
            if not opt[flag_name] or flag_name == "cam":
                continue
            for n in names:
                if P[n].grad is not None:
                    errs[n] = rel(e.g[n].cpu().numpy(), P[n].grad.numpy().reshape(-1))
        errs["tex_extra"] = rel(tex_g, P["tex_extra"].grad.numpy())
        record(dict(test="regulariser_gradients_photometric_off", stage=stage_name, grad_rel_max={k: float("%.3g" % v) for k, v in errs.items()},
                    loss_rel={k: float("%.3g" % v) for k, v in lerr.items()}))
        assert all(v < 1e-4 for v in lerr.values()), lerr
        assert all(v < 1e-4 for v in errs.values()), errs
    finally:
        e.close()
