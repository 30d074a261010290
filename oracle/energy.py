"""Oracle restatement of FlameTracker.compute_energy and its terms.  TEST INFRASTRUCTURE ONLY.

Follows vhap/model/tracker.py: fill_cam_params_into_sample :141-157, forward_flame :213-235, get_albedo :247-258,
compute_lmk_energy :347-389, compute_photometric_energy :391-478, compute_regularization_energy :480-605,
smoothness / joint priors :616-680, compute_laplacian_smoothing_loss :682-690, compute_energy :692-750.
Defaults of the reference's monocular config are assumed where the reference branches on config that the B200
engine does not implement (tex_painted + residual tex_extra, SH lighting in world space, static offset only).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import camera as C
from . import lbs as L
from . import raster as RA
from . import render as RE


def fill_cam_params(params, B, H, W, RT=None):
    """tracker.py:141-157 (uncalibrated): K = [f,f,cx,cy], f = focal*max(h,w); RT = [I | (0,0,-1)] (:1335-1337)."""
    f = params["focal_length"] * max(H, W)
    dt = f.dtype
    cx = torch.full((1,), 0.5 * W, dtype=dt, device=f.device)
    cy = torch.full((1,), 0.5 * H, dtype=dt, device=f.device)
    K = torch.stack([f, f, cx, cy], dim=1).expand(B, -1)
    if RT is None:
        RT = torch.eye(3, 4, dtype=dt, device=f.device)
        RT[2, 3] = -1
        RT = RT[None].expand(B, -1, -1)
    return K, RT


def lmk_energy(lmks, lmk2d, K, RT, img_size, always_enable_jawline=True, disable_jawline=False):
    """tracker.py:347-389.  lmks [B,70,3]; lmk2d [B,68,3] (x_px, y_px, confidence)."""
    gt = lmk2d.clone().to(device=lmks.device, dtype=lmks.dtype)
    xy, conf = gt[:, :, :2], gt[:, :, 2]
    u, v = C.normalize_image_points(xy[:, :, 0], xy[:, :, 1], img_size)
    gt2 = torch.stack([u, v], -1)
    pred = C.world_to_ndc(lmks, RT, K, img_size, flip_y=True)[:, :, :2]
    if (not always_enable_jawline) and disable_jawline:
        diff = gt2[:, 17:68] - pred[:, 17:68]
        conf = conf[:, 17:68]
    else:
        diff = gt2[:, :68] - pred[:, :68]
        conf = conf[:, :68].clone()
        conf[:, 27:36] = conf[:, 27:36] * 10
    return (diff.abs().sum(-1) * conf).mean()


def joint_L2_energy(neck, jaw, eyes, w):
    """tracker.py:650-680."""
    E = 0
    for name, pose in (("neck", neck), ("jaw", jaw), ("eyes", eyes[:, :3]), ("eyes", eyes[:, 3:])):
        rot = L.batch_rodrigues(torch.cat([torch.zeros_like(pose), pose], 0))
        diff = ((rot[[0]] - rot[1:]) ** 2).mean()
        if name == "jaw":
            diff = diff + F.relu(-pose[:, 0]).mean() * 10 + (pose[:, 1:] ** 2).mean() * 3
        elif name == "eyes":
            diff = diff + ((eyes[:, :3] - eyes[:, 3:]) ** 2).mean()
        E = E + diff * getattr(w, f"reg_{name}")
    return E


def laplacian_dense(model_data, dtype):
    indptr, idx, val = model_data.laplacian_csr()
    V = model_data.v_template.shape[0]
    Lm = torch.zeros(V, V, dtype=dtype)
    rows = np.repeat(np.arange(V), np.diff(indptr))
    Lm[torch.as_tensor(rows), torch.as_tensor(idx.astype(np.int64))] = torch.as_tensor(val, dtype=dtype)
    return Lm


def scale_vertex_weights_by_region(model_data, V, scale, regions, dtype, device="cpu"):
    """tracker.py:607-614 with blur_iter = 0."""
    w = torch.ones(V, 1, dtype=dtype, device=device)
    w[torch.as_tensor(model_data.get_vid_by_region(list(regions)))] *= scale
    return w


def regularization_energy(params, ts, stage, cfg, model_data, verts_cano=None, diffuse_detach_normal=None, lap=None, tex_painted=None,
                          uvmask_res=None):
    """compute_regularization_energy (tracker.py:480-605) + the helpers it calls (:607-690).  `diffuse_detach_normal` [B,3,H,W]
    (render_out permuted like tracker.py:322) or None when the stage is not photometric.  Returns the log dict of weighted terms."""
    from vhap_b200.config import opt_dict_for
    w = cfg.w
    dt = params["shape"].dtype
    ts = torch.as_tensor(np.asarray(ts)).long()
    idx_prev = (ts - 1).clamp(0, params["expr"].shape[0] - 1)
    log = {}
    opt = opt_dict_for(stage)
    tracking = "tracking" in stage.name
    if opt["pose"] and tracking:
        log["smooth_pose"] = ((params["translation"][ts] - params["translation"][idx_prev].detach()) ** 2).mean() * w.smooth_trans \
            + ((params["rotation"][ts] - params["rotation"][idx_prev].detach()) ** 2).mean() * w.smooth_rot
    if opt["joints"]:
        log["reg_joint"] = joint_L2_energy(params["neck_pose"][ts], params["jaw_pose"][ts], params["eyes_pose"][ts], w)
        if tracking:
            E = 0
            for k, ww in (("neck_pose", w.smooth_neck), ("jaw_pose", w.smooth_jaw), ("eyes_pose", w.smooth_eyes)):
                E = E + ((params[k][ts] - params[k][idx_prev].detach()) ** 2).mean() * ww
            log["smooth_joint"] = E
    if opt["expr"]:
        log["reg_expr"] = w.reg_expr * (params["expr"][ts] ** 2).mean()
        if tracking:
            log["smooth_expr"] = ((params["expr"][ts] - params["expr"][idx_prev].detach()) ** 2).mean() * w.smooth_expr
    if opt["shape"]:
        log["reg_shape"] = w.reg_shape * (params["shape"] ** 2).mean()
    if opt["texture"]:
        if w.reg_tex_tv is not None:
            tex = (tex_painted if tex_painted is not None else 0) + params["tex_extra"]
            tv_y = (tex[..., :-1, :] - tex[..., 1:, :]) ** 2
            tv_x = (tex[..., :, :-1] - tex[..., :, 1:]) ** 2
            tv = tv_y.reshape(3, -1) + tv_x.reshape(3, -1)          # NB: only valid for square T (as in the reference)
            w_tv = w.reg_tex_tv * cfg.scale_factor ** 2
            if cfg.n_downsample_rgb is not None:
                w_tv /= cfg.n_downsample_rgb ** 2
            log["reg_tex_tv"] = w_tv * tv.mean()
        if w.reg_tex_res_clusters is not None:
            m = torch.as_tensor(uvmask_res) if uvmask_res is not None else torch.as_tensor(model_data.uvmask_res)
            T = params["tex_extra"].shape[-1]
            if m.shape[-1] != T:
                m = m[:: m.shape[0] // T, :: m.shape[1] // T]
            log["reg_tex_res_clusters"] = w.reg_tex_res_clusters * (params["tex_extra"] ** 2 * m[None].to(params["tex_extra"])).mean()
    if opt["lights"]:
        if w.reg_light is not None:
            lu = torch.zeros(9, 3, dtype=dt, device=params["lights"].device)
            lu[0] = np.sqrt(4 * np.pi)
            log["reg_light"] = w.reg_light * ((params["lights"] - lu) ** 2).mean()
        if w.reg_diffuse is not None and diffuse_detach_normal is not None:
            diffuse = diffuse_detach_normal
            log["reg_diffuse"] = w.reg_diffuse * (F.relu(diffuse.max() - 1) + diffuse.var(dim=1).mean())
    if opt["static_offset"] and params.get("static_offset") is not None:
        offset = params["static_offset"]
        V = offset.shape[1]
        if w.reg_offset_lap is not None:
            Lm = lap if lap is not None else laplacian_dense(model_data, dt).to(offset.device)
            base = (verts_cano - offset).detach()
            diff = ((Lm @ (base + offset)) - (Lm @ base).detach()) ** 2
            diff = diff.sum(-1, keepdim=True)
            if len(w.reg_offset_lap_relax_for) > 0:
                diff = diff * scale_vertex_weights_by_region(model_data, V, w.reg_offset_lap_relax_coef, w.reg_offset_lap_relax_for, dt, offset.device)
            log["reg_offset_lap"] = w.reg_offset_lap * diff.mean()
        if w.reg_offset is not None:
            ro = offset.abs()
            if len(w.reg_offset_relax_for) > 0:
                ro = ro * scale_vertex_weights_by_region(model_data, V, w.reg_offset_relax_coef, w.reg_offset_relax_for, dt, offset.device)
            log["reg_offset"] = w.reg_offset * ro.mean()
        if w.reg_offset_rigid is not None:
            r = 0
            for region in w.reg_offset_rigid_for:
                vids = torch.as_tensor(model_data.get_vid_by_region([region]))
                r = r + offset[:, vids, :].var(dim=-2).mean()
            log["reg_offset_rigid"] = w.reg_offset_rigid * r
    return log


def compute_energy(params, sample, stage, cfg, model_data, model, lap=None, disturbance=None, tex_painted=None,
                   return_aux=False, rasterize_fn=None):
    """tracker.py:692-750.  `stage` is a vhap_b200.config.StageConfig or None (evaluation mode).
    sample: rgb [B,3,H,W], lmk2d [B,68,3], timestep_index (array of ints).  Returns E_total, log_dict(, aux)."""
    from vhap_b200.config import opt_dict_for
    w = cfg.w
    ts = torch.as_tensor(np.asarray(sample["timestep_index"])).long()
    idx_prev = (ts - 1).clamp(0, params["expr"].shape[0] - 1)
    B = ts.shape[0]
    gt_rgb = sample["rgb"]
    H, W = gt_rgb.shape[-2:]
    dt = params["shape"].dtype
    log = {}
    verts, verts_cano, lmks = L.flame_forward(
        model, params["shape"][None].expand(B, -1), params["expr"][ts], params["rotation"][ts], params["neck_pose"][ts],
        params["jaw_pose"][ts], params["eyes_pose"][ts], params["translation"][ts], static_offset=params.get("static_offset"))
    ext = sample.get("extrinsic")
    K, RT = fill_cam_params(params, B, H, W, None if ext is None else torch.as_tensor(ext).to(dt))
    if sample.get("intrinsic") is not None:
        K = torch.as_tensor(sample["intrinsic"]).to(dt)
    aux = {"verts": verts, "lmks": lmks, "verts_cano": verts_cano}
    if w.landmark is not None:
        dis = stage.disable_jawline_landmarks if (stage is not None and not w.always_enable_jawline_landmarks) else False
        log["lmk"] = w.landmark * lmk_energy(lmks, sample["lmk2d"], K, RT, (H, W), w.always_enable_jawline_landmarks, dis)
    photometric = stage is None or stage.photometric
    if photometric and w.photo is not None:
        faces = model["faces"]
        cam = C.world_to_camera(verts, RT)
        clip = C.camera_to_clip(cam, K, (H, W))
        rast, rast_db = (rasterize_fn or RA.rasterize)(clip, faces, (H, W))      # (rasterize_fn: bench.py's GPU stand-in feeds ids from the B200 rasteriser)
        verts_uv = model["verts_uv"].clone()
        verts_uv[:, 1] = 1 - verts_uv[:, 1]                                      # tracker.py:315-316
        tex = (tex_painted if tex_painted is not None else 0) + params["tex_extra"]   # tracker.py:247-258
        bg_mode = cfg.render.background_eval if stage is None else cfg.render.background_train
        if bg_mode == "target":
            bg = gt_rgb.permute(0, 2, 3, 1).to(dt)
        else:
            bg = [1.0, 1.0, 1.0] if bg_mode == "white" else [0.0, 0.0, 0.0]
        tex_exc = model_data.get_fid_by_region(list(stage.align_texture_except)) if stage is not None else None
        bnd_exc = model_data.get_vid_by_region(list(stage.align_boundary_except)) if stage is not None else None
        fid2cid = torch.as_tensor(np.concatenate([[0], model_data.fid2cid(cfg.tex_clusters)])).to(verts.device)   # render_nvdiffrast.py:77-79
        out = RE.render_rgba(rast, rast_db, verts, clip, faces, verts_uv, model["faces_uv"], tex, params["lights"], bg,
                             model_data.face_adjacency_opposite(), fid2cid, tex_exc, bnd_exc,
                             disturbance if stage is not None else None)
        pred = out["rgba"].permute(0, 3, 1, 2)
        pred_rgb = pred[:, :3]
        n_fg = (pred[:, [3]].detach() > 0).expand(-1, 3, -1, -1).sum()
        err = (gt_rgb.to(dt) - pred_rgb).abs()
        if sample.get("loss_mask") is not None:                                    # test hook (vhap_set_loss_mask): [B,H,W], False = pixel left out
            err = err * torch.as_tensor(sample["loss_mask"]).to(dt)[:, None]
        log["photo"] = w.photo * (err.sum() / n_fg)                                # tracker.py:438-439
        aux.update(render=out, rast=rast, rast_db=rast_db, clip=clip, n_fg=n_fg)
    if stage is not None:
        diff_dn = aux["render"]["diffuse_detach_normal"].permute(0, 3, 1, 2) if (photometric and "render" in aux) else None
        log.update(regularization_energy(params, ts, stage, cfg, model_data, verts_cano, diff_dn, lap, tex_painted, sample.get("uvmask_res")))
    E_total = torch.stack([v for v in log.values()]).sum()
    log["total"] = E_total
    if return_aux:
        return E_total, log, aux
    return E_total, log


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (default betas/eps, no weight decay, tracker.py:210) single-tensor update, in place on copies."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / np.sqrt(bc2)) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v
