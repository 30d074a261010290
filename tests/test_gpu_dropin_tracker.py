"""The drop-in classes under the reference tracker's own call sequence.

/root/reference is absent on the GPU box, so the reference's FlameTracker cannot be imported here.  `TrackerShaped` below restates,
line for line, the part of it that touches the two replaced objects -- forward_flame (tracker.py:213-235), fill_cam_params_into_sample
(:141-157), rasterize_flame (:262-285), get_background_color (:287-303), render_rgba (:305-331), compute_lmk_energy (:347-389),
compute_photometric_energy (:391-478), the reg_diffuse lines of compute_regularization_energy (:541-550) and the sum of compute_energy
(:692-750) -- with `self.flame = B200FlameHead(...)`, `self.render = B200Renderer(...)` injected exactly where the reference constructs
FlameHead / NVDiffRenderer (tracker.py:58-77).  Everything flows through torch.autograd like in the reference; the result is held to the
float64 oracle (itself pinned against the reference's unmodified compute_energy, tests/golden/make_e2e_golden.py) with the pixel mask of
tests/test_gpu_bench_configs.py.  tools/run_reference_with_b200.py runs the REAL reference tracker with the same injection on a machine
that has both the reference checkout and a GPU."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.scene import make_scene
from tests.test_gpu_bench_configs import dilate, rel, rel_l2, record

pytestmark = pytest.mark.gpu


class TrackerShaped:
    def __init__(self, sc, cfg, stage, device):
        from vhap_b200.reference_api import B200FlameHead, B200Renderer
        m = sc["m"]
        self.cfg, self.stage, self.device = cfg, stage, device
        self.flame = B200FlameHead(cfg.n_shape, cfg.n_expr, model=m)                                   # tracker.py:58-63
        self.render = B200Renderer(use_opengl=False, lighting_type=cfg.render.lighting_type, lighting_space=cfg.render.lighting_space,
                                   disturb_rate_fg=cfg.render.disturb_rate_fg, disturb_rate_bg=cfg.render.disturb_rate_bg,
                                   model=m, tex_size=cfg.tex_resolution)                                 # tracker.py:65-77
        P = {k: torch.tensor(v, device=device, requires_grad=True) for k, v in sc["params"].items()}      # tracker.py:1279-1341
        self.P = P
        self.shape, self.expr, self.rotation, self.translation = P["shape"], P["expr"], P["rotation"], P["translation"]
        self.neck_pose, self.jaw_pose, self.eyes_pose = P["neck_pose"], P["jaw_pose"], P["eyes_pose"]
        self.static_offset, self.tex_extra, self.lights, self.focal_length = P["static_offset"], P["tex_extra"], P["lights"], P["focal_length"]
        self.tex_painted = torch.tensor(sc["tex_painted"], device=device)[None]
        self.RT = torch.eye(3, 4, device=device)
        self.RT[2, 3] = -1

    def fill_cam_params_into_sample(self, sample):
        b, _, h, w = sample["rgb"].shape
        f = self.focal_length * max(h, w)
        cx, cy = torch.tensor([[0.5 * w], [0.5 * h]]).to(f)
        sample["intrinsic"] = torch.stack([f, f, cx, cy], dim=1)
        sample["extrinsic"] = self.RT[None, ...].expand(b, -1, -1)

    def forward_flame(self, timesteps):
        ret = self.flame(self.shape[None, ...].expand(len(timesteps), -1), self.expr[timesteps], self.rotation[timesteps], self.neck_pose[timesteps],
                         self.jaw_pose[timesteps], self.eyes_pose[timesteps], self.translation[timesteps], return_verts_cano=True,
                         static_offset=self.static_offset, dynamic_offset=None)
        verts, verts_cano, lmks = ret[0], ret[1], ret[2]
        albedos = (self.tex_painted + self.tex_extra[None]).expand(len(timesteps), -1, -1, -1)           # get_albedo, tracker.py:247-258
        return verts, verts_cano, lmks, albedos

    def compute_lmk_energy(self, sample, pred_lmks, disable_jawline_landmarks=False):
        img_size = sample["rgb"].shape[-2:]
        lmk2d = sample["lmk2d"].clone().to(pred_lmks)
        lmk2d, confidence = lmk2d[:, :, :2], lmk2d[:, :, 2]
        lmk2d[:, :, 0] = 2 * (lmk2d[:, :, 0] - img_size[1] / 2.0) / img_size[1]                          # normalize_image_points, mesh.py:41-51
        lmk2d[:, :, 1] = 2 * (lmk2d[:, :, 1] - img_size[0] / 2.0) / img_size[0]
        pred_lmk2d = self.render.world_to_ndc(pred_lmks, sample["extrinsic"], sample["intrinsic"], img_size, flip_y=True)[:, :, :2]
        if not self.cfg.w.always_enable_jawline_landmarks and disable_jawline_landmarks:
            diff = lmk2d[:, 17:68] - pred_lmk2d[:, 17:68]
            confidence = confidence[:, 17:68]
        else:
            diff = lmk2d[:, :68] - pred_lmk2d[:, :68]
            confidence = confidence[:, :68]
            confidence[:, 27:36] *= 10
        return (torch.norm(diff, dim=2, p=1) * confidence).mean()

    def compute_photometric_energy(self, sample, verts, faces, albedos, rast_dict, loss_mask=None):
        stage = self.stage
        gt_rgb = sample["rgb"].to(verts)
        lights = self.lights[None]
        bg = self.cfg.render.background_train
        bg_color = gt_rgb.permute(0, 2, 3, 1) if bg == "target" else ([1, 1, 1] if bg == "white" else [0, 0, 0])
        fid = self.flame.mask.get_fid_by_region(list(stage.align_texture_except))
        vid = self.flame.mask.get_vid_by_region(list(stage.align_boundary_except))
        faces_uv = self.flame.textures_idx
        verts_uv = self.flame.verts_uvs.clone()
        verts_uv[:, 1] = 1 - verts_uv[:, 1]
        render_out = self.render.render_rgba(rast_dict, verts, faces, verts_uv, faces_uv, albedos, lights, bg_color, fid, vid, enable_disturbance=True)
        render_out = {k: v.permute(0, 3, 1, 2) for k, v in render_out.items()}
        pred_rgb = render_out["rgba"][:, :3]
        pred_mask = (render_out["rgba"][:, [3]].detach() > 0).expand(-1, 3, -1, -1)
        err = (gt_rgb - pred_rgb).abs()
        if loss_mask is not None:
            err = err * loss_mask[:, None]
        return err.sum() / pred_mask.detach().sum(), render_out

    def compute_energy(self, sample, loss_mask=None):
        w = self.cfg.w
        timesteps = sample["timestep_index"]
        self.fill_cam_params_into_sample(sample)
        verts, verts_cano, lmks, albedos = self.forward_flame(timesteps)
        faces = self.flame.faces
        log = {}
        log["lmk"] = w.landmark * self.compute_lmk_energy(sample, lmks, self.stage.disable_jawline_landmarks)
        H, W = sample["rgb"].shape[-2:]
        rast_dict = self.render.rasterize(verts, faces, sample["extrinsic"], sample["intrinsic"].clone(), (H, W), False, True)
        photo, result = self.compute_photometric_energy(sample, verts, faces, albedos, rast_dict, loss_mask)
        log["photo"] = w.photo * photo
        diffuse = result["diffuse_detach_normal"]                                                         # tracker.py:547-550
        log["reg_diffuse"] = w.reg_diffuse * (F.relu(diffuse.max() - 1) + diffuse.var(dim=1).mean())
        return sum(log.values()), log, result


def test_dropin_classes_under_the_tracker_call_sequence():
    from oracle import energy as E
    from vhap_b200.config import STAGES
    sc = make_scene(B=3, H=256, W=256, T=512, n_t=4, timesteps=[1, 2, 3])
    stage = STAGES["rgb_global_tracking"]
    cfg = copy.deepcopy(sc["cfg"])
    cfg.render.disturb_rate_fg = cfg.render.disturb_rate_bg = None
    w = cfg.w
    w.reg_tex_tv = w.reg_tex_res_clusters = w.reg_offset = w.reg_offset_lap = w.reg_offset_rigid = None
    w.reg_shape = w.reg_expr = w.reg_neck = w.reg_jaw = w.reg_eyes = 0.0
    w.smooth_trans = w.smooth_rot = w.smooth_neck = w.smooth_jaw = w.smooth_eyes = w.smooth_expr = 0.0
    dev = torch.device("cuda:0")
    trk = TrackerShaped(sc, cfg, stage, dev)
    ts = torch.as_tensor(sc["ts"], device=dev)
    sample = dict(rgb=sc["rgb16"].to(torch.float32).to(dev), lmk2d=torch.tensor(sc["lmk2d"], device=dev), timestep_index=ts)
    # ---- oracle
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc["params"].items()}
    osample = dict(rgb=sc["rgb16"].to(torch.float64), lmk2d=torch.tensor(sc["lmk2d"]), timestep_index=sc["ts"])
    tp = torch.tensor(sc["tex_painted"], dtype=torch.float64)
    Et, olog, aux = E.compute_energy(P, osample, stage, cfg, sc["m"], sc["model"], disturbance=None, tex_painted=tp, return_aux=True)
    # ---- forward through the drop-in classes, pixel mask from the two renders
    with torch.no_grad():
        _, _, res = trk.compute_energy(dict(sample))
    rg = res["rgba"].permute(0, 2, 3, 1).cpu().numpy()
    rr = aux["render"]["rgba"].detach().numpy()
    keep = ~dilate(np.abs(rg - rr).max(-1) > 1e-3)
    assert 1 - keep.mean() < 2e-3, 1 - keep.mean()
    Etot, log, res = trk.compute_energy(dict(sample), loss_mask=torch.as_tensor(keep.copy(), device=dev).float())
    Etot.backward()
    pred = aux["render"]["rgba"].permute(0, 3, 1, 2)[:, :3]
    err = (osample["rgb"] - pred).abs() * torch.as_tensor(keep.copy()).to(torch.float64)[:, None]
    photo_m = cfg.w.photo * err.sum() / aux["n_fg"]
    (Et - olog["photo"] + photo_m).backward()
    lerr = dict(lmk=abs(float(log["lmk"]) - float(olog["lmk"])) / float(olog["lmk"]), photo=abs(float(log["photo"]) - float(photo_m)) / float(photo_m),
                reg_diffuse=abs(float(log["reg_diffuse"]) - float(olog["reg_diffuse"])) / float(olog["reg_diffuse"]))
    errs, errs2 = {}, {}
    for k in ("shape", "expr", "rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose", "static_offset", "lights", "focal_length", "tex_extra"):
        g = trk.P[k].grad
        assert g is not None and float(g.abs().sum()) > 0, f"no gradient reached {k} through the drop-in classes"
        errs[k], errs2[k] = rel(g.cpu().numpy(), P[k].grad.numpy()), rel_l2(g.cpu().numpy(), P[k].grad.numpy())
    record(dict(test="dropin_tracker_call_sequence", frac_masked=float(1 - keep.mean()), loss_rel={k: float("%.3g" % v) for k, v in lerr.items()},
                grad_rel_max={k: float("%.3g" % v) for k, v in errs.items()}, grad_rel_l2={k: float("%.3g" % v) for k, v in errs2.items()}))
    assert all(v < 2e-4 for v in lerr.values()), lerr
    # the lights gradient includes the reg_diffuse path through diffuse_detach_normal (render_nvdiffrast.py:402-403, tracker.py:547-550)
    assert all(v < 1e-2 for v in errs.values()), errs          # fp32 drop-in path vs fp64 oracle from the parameters (see TOL in test_gpu_bench_configs.py)


def test_renderer_rejects_foreign_topology():
    from vhap_b200.reference_api import B200Renderer
    sc = make_scene(B=1, H=32, W=32, T=64, n_t=2, timesteps=[0])
    rnd = B200Renderer(lighting_type="SH", model=sc["m"], tex_size=64)
    faces = torch.as_tensor(sc["m"].faces.astype(np.int64), device=rnd.eng.dev)
    bad = faces.clone()
    bad[0] = bad[0].flip(0)
    with pytest.raises(ValueError):
        rnd.rasterize(torch.zeros(1, 5143, 3, device=rnd.eng.dev), bad, torch.eye(3, 4, device=rnd.eng.dev)[None], torch.tensor([[100.0, 100.0, 16.0, 16.0]], device=rnd.eng.dev), (32, 32))
