"""Host side of the B200 photometric engine: parameter store, stage configuration, input staging and the per-iteration
`step()` that replaces FlameTracker.optimize_iter (vhap/model/tracker.py:1418-1435: compute_energy, backward, Adam).

PyTorch is used for device memory, streams and torch.distributed only; all arithmetic of the hot path happens in
libvhap_b200.so (hand-written sm_100a CUDA, include/vhap_b200.h) through raw pointers.  There is no fallback path."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .config import adam_param_lrs, EngineConfig, StageConfig, STAGES, opt_dict_for
from .flame_model import FlameModelData

PER_FRAME = (("expr", None), ("rotation", 3), ("neck_pose", 3), ("jaw_pose", 3), ("eyes_pose", 6), ("translation", 3))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Batch:
    """One staged batch of frames (the reference's `sample` dict, video_dataset.py:209-241)."""

    def __init__(self, B, H, W, timesteps, target, lmk2d, RT=None, K=None, geo=None, geo_ts=None):
        self.B, self.H, self.W = B, H, W
        self.timesteps, self.target, self.lmk2d, self.RT, self.K = timesteps, target, lmk2d, RT, K
        self.geo, self.geo_ts = geo, geo_ts              # view sharing (see vhap_frame_batch): FLAME once per distinct timestep
        fmt = 1 if (target is not None and target.dtype == torch.uint8) else 0      # uint8 RGB as decoded / fp16 RGBA
        self.c = _lib.FrameBatch(B, H, W, fmt, _ptr(timesteps), _ptr(target), _ptr(lmk2d), _ptr(RT), _ptr(K), _ptr(geo), _ptr(geo_ts),
                                 0 if geo_ts is None else int(geo_ts.numel()))


class Engine:
    def __init__(self, model: FlameModelData, cfg: EngineConfig, n_timesteps: int, device: str = "cuda:0",
                 tex_painted: Optional[np.ndarray] = None, world_size: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("vhap_b200.Engine needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.L = _lib.lib()
        self.model, self.cfg, self.n_t = model, cfg, n_timesteps
        self.dev = torch.device(device)
        self.world_size = world_size
        torch.cuda.set_device(self.dev)
        self.T = cfg.tex_resolution
        V, F, K = model.v_template.shape[0], model.faces.shape[0], model.shapedirs.shape[2]
        self.V, self.F, self.K = V, F, K
        self.n_shape, self.n_expr = model.n_shape, model.n_expr
        # ---- static tables -> ctx
        vuv = model.verts_uv.astype(np.float32).copy()
        vuv[:, 1] = 1.0 - vuv[:, 1]                                         # tracker.py:315-316
        fid2cid = np.concatenate([[0], model.fid2cid(cfg.tex_clusters)]).astype(np.uint8)   # render_nvdiffrast.py:77-79
        vf_indptr, vf_faces = model.vertex_face_csr()
        lap_ip, lap_idx, lap_val = model.laplacian_csr()
        keep = dict(
            v_template=np.ascontiguousarray(model.v_template, np.float32), shapedirs=np.ascontiguousarray(model.shapedirs, np.float32),
            posedirs=np.ascontiguousarray(model.posedirs, np.float32), Jreg=np.ascontiguousarray(model.J_regressor, np.float32),
            lbs=np.ascontiguousarray(model.lbs_weights, np.float32), faces=np.ascontiguousarray(model.faces, np.int32),
            faces_uv=np.ascontiguousarray(model.faces_uv, np.int32), vuv=vuv, lmk_f=np.ascontiguousarray(model.lmk_faces_idx, np.int32),
            lmk_b=np.ascontiguousarray(model.lmk_bary, np.float32), adj=np.ascontiguousarray(model.face_adjacency_opposite(), np.int32),
            fid2cid=fid2cid, vf_ip=vf_indptr, vf_f=vf_faces, lap_ip=lap_ip, lap_idx=lap_idx, lap_val=lap_val)
        hp = lambda a: a.ctypes.data_as(C.c_void_p)
        md = _lib.MeshDesc(V, F, vuv.shape[0], K, model.n_shape, model.lmk_faces_idx.shape[0], int(fid2cid.max()) + 1,
                           hp(keep["v_template"]), hp(keep["shapedirs"]), hp(keep["posedirs"]), hp(keep["Jreg"]), hp(keep["lbs"]),
                           hp(keep["faces"]), hp(keep["faces_uv"]), hp(keep["vuv"]), hp(keep["lmk_f"]), hp(keep["lmk_b"]), hp(keep["adj"]),
                           hp(keep["fid2cid"]), hp(keep["vf_ip"]), hp(keep["vf_f"]), hp(keep["lap_ip"]), hp(keep["lap_idx"]), hp(keep["lap_val"]))
        self.ctx = C.c_void_p()
        self._ck(self.L.vhap_ctx_create(C.byref(self.ctx), C.byref(md), self.T, self.dev.index or 0), None)
        self.fid2cid = fid2cid
        # ---- parameter slab (shared parameters first = the data-parallel allreduce slab)
        self.layout: Dict[str, tuple] = {}
        off = 0
        for name, n in (("shape", model.n_shape), ("static_offset", 3 * V), ("lights", 27), ("focal_length", 1)):
            self.layout[name] = (off, n)
            off += n
        self.n_shared = off
        for name, d in PER_FRAME:
            d = model.n_expr if d is None else d
            self.layout[name] = (off, n_timesteps * d)
            off += n_timesteps * d
        self.n_small = off
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.slab, self.grad, self.m, self.v = z(off), z(off), z(off), z(off)
        nt = 3 * self.T * self.T
        self.tex_extra, self.tex_m, self.tex_v, self.tex_grad_dense = z(nt), z(nt), z(nt), None
        self._dp_buf = None
        if world_size > 1:
            # data parallel: the dense texture gradient and the gradient slab live in ONE buffer -> one all-reduce per step
            self._dp_buf = z(nt + off)
            self.tex_grad_dense, self.grad = self._dp_buf[:nt], self._dp_buf[nt:]
        self.p = {k: self.slab[o:o + n] for k, (o, n) in self.layout.items()}
        self.g = {k: self.grad[o:o + n] for k, (o, n) in self.layout.items()}
        self.p["lights"][0:3] = float(np.sqrt(4 * np.pi))                  # tracker.py:1301-1304
        self.p["focal_length"][0] = 1.5                                     # tracker.py:1333
        self.losses = z(_lib.N_LOSS)
        self.slab_local, self.slab_global = z(8), z(8)
        self.tex_painted = None
        if tex_painted is not None:
            tp = torch.as_tensor(np.ascontiguousarray(tex_painted, np.float32), device=self.dev)
            if tp.dim() != 3 or tp.shape[0] != 3:
                raise ValueError(f"tex_painted must be [3,T,T], got {tuple(tp.shape)}")
            if tuple(tp.shape[1:]) != (self.T, self.T):                      # flame.py:655-657: resized to tex_resolution
                tp = torch.nn.functional.interpolate(tp[None], (self.T, self.T), mode="bilinear", align_corners=False)[0]
            self.tex_painted = tp.contiguous().reshape(-1)
            self._ck(self.L.vhap_set_tex_painted(self.ctx, self.tex_painted.data_ptr(), self._stream()))
        if cfg.w.blur_iter != 0:
            raise NotImplementedError("blur_iter != 0 (blurred region weights, tracker.py:607-614) is not supported by the b200 backend")
        if tuple(cfg.w.reg_tex_res_for) != ("sclerae", "teeth"):
            raise NotImplementedError("reg_tex_res_for other than ('sclerae', 'teeth'): the residual mask is the model's uvmask_res (flame.py:1042-1055)")
        self.stage: Optional[StageConfig] = None
        self.step_count = 0
        self.global_step = 0
        self._lr_scale = 1.0
        self._graph_live = False
        self._tex_split = False
        self._inj = None
        self._loss_mask = None
        self._peer_slab = False
        self._tex_persist = False
        if world_size > 1:
            # the dense texture gradient is a persistent buffer of this object: its fold may run on the library's aux stream
            torch.cuda.current_stream(self.dev).synchronize()
            self.L.vhap_set_tex_grad_persistent(self.ctx, 1)
            self._tex_persist = True
        self.rebuild_texture()

    # learning-rate scale (lr_scale of configure_optimizer, tracker.py:159-211, times ExponentialLR's 0.9 ** epoch, :1407-1412).  While
    # step graphs are live the captured Adam kernels read it from device memory, so assigning it between replays takes effect.
    @property
    def lr_scale(self) -> float:
        return self._lr_scale

    @lr_scale.setter
    def lr_scale(self, v: float):
        self._lr_scale = float(v)
        if self._graph_live:
            self._ck(self.L.vhap_set_lr_scale(self.ctx, self._lr_scale, self._stream()))

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _ck(self, r, _=None):
        if r != 0:
            raise RuntimeError("libvhap_b200: " + self.L.vhap_last_error(self.ctx).decode())

    def close(self):
        if self.ctx:
            torch.cuda.synchronize(self.dev)
            self.L.vhap_ctx_destroy(self.ctx)
            self.ctx = None

    def load_params(self, params: Dict[str, np.ndarray]):
        for k, v in params.items():
            t = torch.as_tensor(np.ascontiguousarray(v, np.float32)).reshape(-1).to(self.dev)
            if k == "tex_extra":
                self.tex_extra.copy_(t)
            else:
                self.p[k].copy_(t)
        self.rebuild_texture()

    def get_params(self) -> Dict[str, np.ndarray]:
        self._flush_pending_tex()                  # (pipelined graph replay keeps one texture update in flight)
        out = {k: v.detach().cpu().numpy().copy() for k, v in self.p.items()}
        out["tex_extra"] = self.tex_extra.cpu().numpy().reshape(3, self.T, self.T)
        return out

    # ------------------------------------------------------------------ on-disk format (tracker.py:79-129, 1152-1218)
    def save_result(self, out_dir, timestep_ids, n_processed_frames, image_size, fname=None, epoch=None):
        """FlameTracker.save_result: writes tracked_flame_params[_<epoch>].npz with the reference's keys / shapes (vhap_b200/io_params.py)."""
        from . import io_params
        rep = io_params.engine_params_to_report(self.get_params(), timestep_ids, n_processed_frames, image_size, calibrated=self.cfg.calibrated)
        return io_params.save_tracked_flame_params(out_dir, rep, fname, epoch)

    def load_from_tracked_flame_params(self, fp, warn=print):
        """FlameTracker.load_from_tracked_flame_params: the reference's loader semantics (first min(N_t, len) rows, optional keys)."""
        from . import io_params
        rep = io_params.load_tracked_flame_params(fp)
        self.load_params(io_params.report_to_engine_params(rep, self.get_params(), self.n_t, calibrated=self.cfg.calibrated, warn=warn))

    def rebuild_texture(self):
        self._ck(self.L.vhap_tex_rebuild(self.ctx, self.tex_extra.data_ptr(), self._stream()), None)

    def reserve(self, B, H, W):
        self._ck(self.L.vhap_ctx_reserve(self.ctx, B, H, W), None)

    def _c_params(self) -> _lib.Params:
        p = self.p
        return _lib.Params(_ptr(p["shape"]), _ptr(p["expr"]), _ptr(p["rotation"]), _ptr(p["neck_pose"]), _ptr(p["jaw_pose"]), _ptr(p["eyes_pose"]),
                           _ptr(p["translation"]), _ptr(p["static_offset"]), _ptr(p["lights"]),
                           None if self.cfg.calibrated else _ptr(p["focal_length"]), _ptr(self.tex_extra), self.n_t)

    def _c_grads(self, opt: dict) -> _lib.Grads:
        g = self.g
        on = lambda flag, t: _ptr(t) if flag else None
        tex_ptr = self.L.vhap_tex_grad_ptr(self.ctx) if opt["texture"] else None
        return _lib.Grads(on(opt["shape"], g["shape"]), on(opt["expr"], g["expr"]), on(opt["pose"], g["rotation"]), on(opt["joints"], g["neck_pose"]),
                          on(opt["joints"], g["jaw_pose"]), on(opt["joints"], g["eyes_pose"]), on(opt["pose"], g["translation"]),
                          on(opt["static_offset"], g["static_offset"]), on(opt["lights"], g["lights"]),
                          on(opt["cam"] and not self.cfg.calibrated, g["focal_length"]), tex_ptr)

    # ------------------------------------------------------------------ stages
    def set_stage(self, stage, lr_scale: float = 1.0):
        """begin of FlameTracker.optimize_stage (tracker.py:1391-1416): new Adam state, stage masks."""
        if isinstance(stage, str):
            stage = STAGES[stage]
        if getattr(self, "_graph_live", False):
            raise RuntimeError("set_stage while step graphs are live: call graph_end() first")
        self.stage = stage
        self.lr_scale = lr_scale
        self.step_count = 0
        self.m.zero_(); self.v.zero_(); self.tex_m.zero_(); self.tex_v.zero_()
        m, w = self.model, self.cfg.w
        face_flags = np.zeros(self.F, np.uint8)
        vert_flags = np.zeros(self.V, np.uint8)
        if stage is not None:
            face_flags[m.get_fid_by_region(list(stage.align_texture_except))] = 1
            vert_flags[m.get_vid_by_region(list(stage.align_boundary_except))] = 1
        w_off = np.ones(self.V, np.float32)
        w_off[m.get_vid_by_region(list(w.reg_offset_relax_for))] *= w.reg_offset_relax_coef
        w_lap = np.ones(self.V, np.float32)
        w_lap[m.get_vid_by_region(list(w.reg_offset_lap_relax_for))] *= w.reg_offset_lap_relax_coef
        ip, vids = [0], []
        for r in w.reg_offset_rigid_for:
            ids = m.get_vid_by_region([r])
            vids.append(ids.astype(np.int32))
            ip.append(ip[-1] + len(ids))
        ip = np.asarray(ip, np.int32)
        vids = np.concatenate(vids).astype(np.int32) if vids else np.zeros(0, np.int32)
        mask = m.uvmask_res
        if mask is None:
            mask = np.zeros((self.T, self.T), bool)
        if mask.shape[0] != self.T:
            if mask.shape[0] < self.T or mask.shape[0] % self.T or mask.shape[0] != mask.shape[1]:
                raise ValueError(f"uvmask_res {mask.shape} cannot be subsampled to the texture resolution {self.T}")
            mask = mask[:: mask.shape[0] // self.T, :: mask.shape[1] // self.T]
        mask = np.ascontiguousarray(mask.astype(np.uint8))
        hp = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self.L.vhap_set_stage_masks(self.ctx, hp(face_flags), hp(vert_flags), hp(w_off), hp(w_lap), hp(ip), hp(vids), len(ip) - 1, hp(mask)), None)

    def _c_stage(self, training: bool) -> _lib.StageCfg:
        w, st = self.cfg.w, self.stage
        nz = lambda x: -1.0 if x is None else float(x)
        s = _lib.StageCfg()
        s.w_landmark, s.w_photo = nz(w.landmark), nz(w.photo)
        s.w_reg_shape, s.w_reg_expr, s.w_reg_neck, s.w_reg_jaw, s.w_reg_eyes = w.reg_shape, w.reg_expr, w.reg_neck, w.reg_jaw, w.reg_eyes
        w_tv = w.reg_tex_tv
        if w_tv is not None:
            w_tv = w_tv * self.cfg.scale_factor ** 2
            if self.cfg.n_downsample_rgb is not None:
                w_tv /= self.cfg.n_downsample_rgb ** 2
        s.w_reg_tex_tv, s.w_reg_tex_res, s.w_reg_diffuse, s.w_reg_light = nz(w_tv), nz(w.reg_tex_res_clusters), nz(w.reg_diffuse), nz(w.reg_light)
        s.w_reg_offset, s.w_reg_offset_lap, s.w_reg_offset_rigid = nz(w.reg_offset), nz(w.reg_offset_lap), nz(w.reg_offset_rigid)
        s.w_smooth_trans, s.w_smooth_rot, s.w_smooth_neck, s.w_smooth_jaw = w.smooth_trans, w.smooth_rot, w.smooth_neck, w.smooth_jaw
        s.w_smooth_eyes, s.w_smooth_expr = w.smooth_eyes, w.smooth_expr
        train = training and st is not None
        s.photometric = 1 if (st is None or st.photometric) else 0
        s.jawline_off = 1 if (train and (not w.always_enable_jawline_landmarks) and st.disable_jawline_landmarks) else 0
        s.tracking = 1 if (train and "tracking" in st.name) else 0
        s.training = 1 if train else 0
        opt = opt_dict_for(st) if train else {k: False for k in ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset")}
        s.opt_pose, s.opt_joints, s.opt_expr, s.opt_shape = int(opt["pose"]), int(opt["joints"]), int(opt["expr"]), int(opt["shape"])
        s.opt_texture, s.opt_lights, s.opt_static_offset, s.opt_cam = int(opt["texture"]), int(opt["lights"]), int(opt["static_offset"]), int(opt["cam"])
        bg = self.cfg.render.background_train if train else self.cfg.render.background_eval
        if bg == "target":
            s.bg_mode = 0
        elif bg in ("white", "black"):
            s.bg_mode = 1
            s.bg_color[0] = s.bg_color[1] = s.bg_color[2] = 1.0 if bg == "white" else 0.0
        else:
            raise NotImplementedError(f"Unknown background mode: {bg}")                # tracker.py:302
        if self.cfg.render.lighting_type != "SH":
            raise NotImplementedError(f"Unknown lighting type: {self.cfg.render.lighting_type}")   # render_nvdiffrast.py:346
        if self.cfg.render.lighting_space != "world":
            raise NotImplementedError(f"Unknown lighting space: {self.cfg.render.lighting_space}")  # render_nvdiffrast.py:379
        s.disturb_rate_fg = nz(self.cfg.render.disturb_rate_fg)
        s.disturb_rate_bg = nz(self.cfg.render.disturb_rate_bg)
        s.rng_seed, s.rng_step = 0x5EED5EED, self.global_step
        s.shared_scale = 1.0 / self.world_size
        return s

    # ------------------------------------------------------------------ input staging
    def stage_sample(self, rgb, lmk2d, timesteps, RT=None, K=None, non_blocking=True, share_views=None) -> Batch:
        """rgb: [B,3,H,W] float (host or device, like sample['rgb']) or an already packed [B,H,W,4] fp16 tensor
        (host pinned memory for the end-to-end path).  Returns device-resident Batch."""
        if rgb.dtype == torch.uint8:
            # the image as the dataset decodes it, [B,H,W,3] uint8 (video_dataset.py:209-241 before F.to_tensor): 3 bytes per pixel over
            # PCIe, the /255 of to_tensor happens inside the kernels
            if rgb.dim() != 4 or rgb.shape[-1] != 3:
                raise ValueError(f"uint8 targets must be [B,H,W,3], got {tuple(rgb.shape)}")
            tgt = rgb.to(self.dev, non_blocking=non_blocking).contiguous()
            B, H, W = tgt.shape[:3]
        elif rgb.dtype == torch.float16 and rgb.dim() == 4 and rgb.shape[-1] == 4:
            tgt = rgb.to(self.dev, non_blocking=non_blocking)
            B, H, W = tgt.shape[:3]
        else:
            B, _, H, W = rgb.shape
            r = rgb.to(self.dev, non_blocking=non_blocking)
            tgt = torch.cat([r.permute(0, 2, 3, 1), torch.zeros(B, H, W, 1, device=self.dev, dtype=r.dtype)], -1).to(torch.float16).contiguous()
        lm = torch.as_tensor(lmk2d, dtype=torch.float32)
        if lm.dim() != 3 or lm.shape[0] != B or lm.shape[1] < 68 or lm.shape[2] != 3:
            raise ValueError(f"lmk2d must be [B,>=68,3], got {tuple(lm.shape)}")
        lm = lm[:, :68].to(self.dev, non_blocking=non_blocking).contiguous()                    # tracker.py:358-362 uses [:, :68]
        ts = torch.as_tensor(np.asarray(timesteps), dtype=torch.int32).to(self.dev, non_blocking=non_blocking)
        if ts.shape != (B,):
            raise ValueError(f"timestep_index must have {B} entries")
        if self.cfg.calibrated and (RT is None or K is None):
            raise AssertionError("calibrated data: 'intrinsic' and 'extrinsic' must be in the sample")   # tracker.py:145-147
        RTd, Kd = None, None
        if RT is not None:                                  # sample['extrinsic']: [B,3,4] or [B,4,4] world-to-camera (render_nvdiffrast.py:162-179)
            RTd = torch.as_tensor(RT, dtype=torch.float32)
            if RTd.dim() == 2:
                RTd = RTd[None].expand(B, -1, -1)
            if RTd.shape[-2:] not in ((3, 4), (4, 4)) or RTd.shape[0] != B:
                raise ValueError(f"Expected RT to be (N, 3, 4) or (N, 4, 4) but got: {tuple(RTd.shape)}")
            RTd = RTd[:, :3, :].to(self.dev).contiguous()
        if K is not None:                                   # sample['intrinsic']: [B,3,3] matrices or [B,4] = (fx,fy,cx,cy) (render_nvdiffrast.py:133-141)
            Kd = torch.as_tensor(K, dtype=torch.float32)
            if Kd.dim() >= 2 and Kd.shape[-2:] == (3, 3):
                Kd = Kd.reshape(-1, 3, 3)
                Kd = torch.stack([Kd[:, 0, 0], Kd[:, 1, 1], Kd[:, 0, 2], Kd[:, 1, 2]], -1)
            elif Kd.shape[-1] == 4:
                Kd = Kd.reshape(-1, 4)
            else:
                raise ValueError(f"Expected K to be (N, 3, 3) or (N, 4) but got: {tuple(Kd.shape)}")   # render_nvdiffrast.py:141
            if Kd.shape[0] == 1 and B > 1:
                Kd = Kd.expand(B, -1)
            if Kd.shape[0] != B:
                raise ValueError(f"K has {Kd.shape[0]} rows for a batch of {B}")
            Kd = Kd.to(self.dev).contiguous()
        self.reserve(B, H, W)
        # view sharing: several cameras of one timestep (NeRSemble: all 16 views of a batch) share one FLAME evaluation.  Default: on
        # whenever timesteps repeat inside a batch with per-frame cameras; the arithmetic per view is unchanged (tests/test_gpu_views.py)
        geo = geo_ts = None
        tsn = np.asarray(timesteps).reshape(-1)
        uniq, inv = np.unique(tsn, return_inverse=True)
        if share_views is None:
            share_views = RTd is not None and len(uniq) < B
        if share_views and len(uniq) < B:
            geo = torch.as_tensor(inv.astype(np.int32)).to(self.dev)
            geo_ts = torch.as_tensor(uniq.astype(np.int32)).to(self.dev)
        return Batch(B, H, W, ts, tgt, lm, RTd, Kd, geo, geo_ts)

    def dp_connect(self, rank: int, world: int, gather_bytes) -> None:
        """data parallel: map every rank's slab mailbox over NVLink (CUDA IPC).  `gather_bytes(b)` returns the list of every rank's bytes
        (e.g. torch.distributed.all_gather_object).  Afterwards energy() needs no reduce_fn: see include/vhap_b200.h vhap_dp_init."""
        h = (C.c_ubyte * 64)()
        self._ck(self.L.vhap_dp_init(self.ctx, rank, world, h))
        handles = gather_bytes(bytes(h))
        if len(handles) != world or any(len(x) != 64 for x in handles):
            raise RuntimeError("dp_connect: gather_bytes must return one 64-byte handle per rank")
        buf = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        self._ck(self.L.vhap_dp_connect(self.ctx, buf))
        self._peer_slab = world > 1

    def dp_tex_connect(self, grm_ptrs, grm_mc, exrm_ptrs, exrm_mc) -> None:
        """hand the symmetric exchange buffers of the peer-memory texture update to the library (include/vhap_b200.h vhap_dp_tex_connect)"""
        n = len(grm_ptrs)
        a = (C.c_void_p * n)(*[int(p) for p in grm_ptrs])
        b = (C.c_void_p * n)(*[int(p) for p in exrm_ptrs])
        self._ck(self.L.vhap_dp_tex_connect(self.ctx, a, C.c_void_p(int(grm_mc) or None), b, C.c_void_p(int(exrm_mc) or None)))

    def dp_status(self) -> int:
        out = C.c_int32(0)
        self._ck(self.L.vhap_dp_status(self.ctx, C.byref(out)))
        return out.value

    def set_loss_mask(self, mask):
        """test hook: [B,H,W] bool/uint8 (image orientation like sample['rgb']); False = pixel left out of the L1 photometric sum."""
        if mask is None:
            self._loss_mask = None
            self.L.vhap_set_loss_mask(self.ctx, None)
            return
        self._loss_mask = torch.as_tensor(mask).to(torch.uint8).to(self.dev).contiguous()
        self.L.vhap_set_loss_mask(self.ctx, self._loss_mask.data_ptr())

    def inject_random(self, w_fg, w_bg, u):
        """test hook: fix the disturbance randomness (render_nvdiffrast.py:429-435,455)."""
        if w_fg is None:
            self._inj = None
            self.L.vhap_set_injected_random(self.ctx, None, None)
            return
        wb = (w_fg.to(torch.uint8) | (w_bg.to(torch.uint8) << 1)).to(self.dev).contiguous()
        uu = u.to(torch.float32).to(self.dev).contiguous()
        self._inj = (wb, uu)
        self.L.vhap_set_injected_random(self.ctx, wb.data_ptr(), uu.data_ptr())

    # ------------------------------------------------------------------ energy / step
    def zero_grad(self):
        self.grad.zero_()

    def energy(self, batch: Batch, backward: bool = True, training: bool = True, global_B: Optional[int] = None, reduce_fn=None) -> torch.Tensor:
        """compute_energy (+ backward).  Returns the loss vector (device tensor, see _lib.LOSS_NAMES).  With `reduce_fn`
        (data parallel) the forward slab is reduced across ranks between the forward and backward halves."""
        cs = self._c_stage(training)
        cp = self._c_params()
        train = training and self.stage is not None
        opt = opt_dict_for(self.stage) if train else None
        cg = self._c_grads(opt) if (backward and train) else None
        s = self._stream()
        gB = batch.B if global_B is None else global_B
        self._ck(self.L.vhap_energy_forward(self.ctx, C.byref(cp), C.byref(batch.c), C.byref(cs), self.slab_local.data_ptr(), s), None)
        gslab = self.slab_local                        # single process: the local slab IS the global one (no copy kernel on the critical chain)
        if reduce_fn is not None and not self._peer_slab:      # (peer mailboxes: the two kernels exchange the slab themselves, dp_connect)
            reduce_fn(self.slab_local, self.slab_global)
            gslab = self.slab_global
        self._ck(self.L.vhap_energy_backward(self.ctx, C.byref(cp), C.byref(batch.c), C.byref(cs), gslab.data_ptr(),
                                             self.slab_local.data_ptr(), gB, C.byref(cg) if cg is not None else None,
                                             self.losses.data_ptr(), s), None)
        self._last = (cs, opt)
        return self.losses

    def texture_grad_dense(self, training=True, with_losses=True) -> torch.Tensor:
        """Folds the texel-gradient pyramid (+ TV / residual regularisers) into a dense [3,T,T] gradient (no Adam)."""
        if self.tex_grad_dense is None:
            self.tex_grad_dense = torch.zeros(3 * self.T * self.T, dtype=torch.float32, device=self.dev)
        if not self._tex_persist:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("texture_grad_dense: call it once eagerly (or construct the Engine with world_size > 1) before capturing")
            torch.cuda.current_stream(self.dev).synchronize()       # the zero fill must not race the library's aux stream
            self.L.vhap_set_tex_grad_persistent(self.ctx, 1)        # persistent from here on: the fold may leave the caller's stream
            self._tex_persist = True
        cs = self._c_stage(training)
        self._ck(self.L.vhap_tex_reg_fold_adam(self.ctx, self.tex_extra.data_ptr(), self.tex_grad_dense.data_ptr(), None, None, 0.0, 1,
                                               C.byref(cs), 1.0, self.losses.data_ptr() if with_losses else None, self._stream()), None)
        return self.tex_grad_dense.view(3, self.T, self.T)

    def _host_lr_scale(self):
        # with live step graphs the scale lives in device memory (vhap_set_lr_scale) and the kernels apply it themselves
        return 1.0 if self._graph_live else self._lr_scale

    def _lr(self, name):
        from .config import _LR_OF                                           # tracker.py:159-211 (pinned: tests/test_config_golden.py)
        return getattr(self.cfg.lr, _LR_OF.get(name, "base")) * self._host_lr_scale()

    def _shard_bufs(self, comm):
        """exchange buffers of the sharded texture update (row-major [T,3,T] so that a rank's row band is one contiguous chunk)"""
        key = (comm.world, tuple(comm.owned()))
        if getattr(self, "_shard_key", None) != key:
            if self.T % (comm.world * 8):
                raise ValueError(f"sharded texture update: texture size {self.T} must be a multiple of 8 * world ({comm.world})")
            nt, nb = 3 * self.T * self.T, 3 * self.T * self.T // comm.world
            z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.dev)
            self._shard = dict(g_rm=z(nt), ex_rm=z(nt), g={i: z(nb) for i in comm.owned()}, ex={i: z(nb) for i in comm.owned()})
            self._shard_key = key
        return self._shard

    def tex_update(self, allreduce_fn=None, deferred=False, reg_loss=True, tex_comm=None, second_half=False):
        """Texture part of the Adam step on the CURRENT stream.  deferred=True: the update of the PREVIOUS step, executed at the start
        of the next one (pipelined graph replay): no in-call fork, device Adam step - 1, no loss bookkeeping of the finished step, and
        the regulariser loss values of the updated texture are produced for the step that is about to run.
        tex_comm (data parallel, parallel.TexShardComm): the SHARDED update -- photometric fold -> reduce-scatter by row band -> the band
        owner adds TV / residual gradients and runs Adam on its T/world rows -> all-gather of the updated rows -> pyramid rebuild;
        allreduce_fn without tex_comm: the round-1 path (fold + regularisers -> dense all-reduce -> full-texture Adam on every rank)."""
        cs = self._c_stage(True)
        s = self._stream()
        if deferred:
            self.L.vhap_tex_defer(self.ctx, -1)
        if tex_comm is not None and getattr(tex_comm, "peer", False):
            # peer-memory update (csrc/dp_tex.cu): fold -> barrier -> in-switch band reduction -> band Adam -> multicast store -> barrier -> rebuild
            if not deferred:
                self._ck(self.L.vhap_tex_reg_loss(self.ctx, self.tex_extra.data_ptr(), C.byref(cs), s))
                self._ck(self.L.vhap_assemble_losses(self.ctx, C.byref(cs), self.losses.data_ptr(), s))
            # second_half: the fold / band reduction of this update already ran at the end of the step that produced the gradients
            # (vhap_dp_tex_part1 in _step_body, beside the geometry backward); what is left is Adam on the band, the multicast store, the rebuild
            fn = self.L.vhap_dp_tex_part2 if second_half else self.L.vhap_dp_tex_update
            self._ck(fn(self.ctx, self.tex_extra.data_ptr(), self.tex_m.data_ptr(), self.tex_v.data_ptr(), self._lr("tex"), self.step_count, C.byref(cs), s))
        elif tex_comm is not None:
            sh = self._shard_bufs(tex_comm)
            if not deferred:
                # loss values of the texture regularisers for the texture this step rendered with (the fused path gets them from its fold)
                self._ck(self.L.vhap_tex_reg_loss(self.ctx, self.tex_extra.data_ptr(), C.byref(cs), s))
                self._ck(self.L.vhap_assemble_losses(self.ctx, C.byref(cs), self.losses.data_ptr(), s))
            self._ck(self.L.vhap_tex_fold_grad_rm(self.ctx, self.tex_extra.data_ptr(), sh["g_rm"].data_ptr(), s))
            tex_comm.reduce_scatter(sh["g_rm"], sh["g"])
            rows = self.T // tex_comm.world
            for i in tex_comm.owned():
                self._ck(self.L.vhap_tex_band_adam(self.ctx, self.tex_extra.data_ptr(), sh["g"][i].data_ptr(), i * rows, (i + 1) * rows, self.tex_m.data_ptr(),
                                                   self.tex_v.data_ptr(), self._lr("tex"), self.step_count, C.byref(cs), sh["ex"][i].data_ptr(), s))
            tex_comm.all_gather(sh["ex_rm"], sh["ex"])
            self._ck(self.L.vhap_tex_rebuild_rm(self.ctx, self.tex_extra.data_ptr(), sh["ex_rm"].data_ptr(), s))
        elif allreduce_fn is None:
            # fused: fold + regularisers + Adam + pyramid; eagerly it runs on an aux stream right behind the fused backward
            # (vhap_tex_reg_fold_adam waits only for the texel-gradient event), the call itself only joins
            self._ck(self.L.vhap_tex_reg_fold_adam(self.ctx, self.tex_extra.data_ptr(), None, self.tex_m.data_ptr(), self.tex_v.data_ptr(),
                                                   self._lr("tex"), self.step_count, C.byref(cs), 1.0,
                                                   None if deferred else self.losses.data_ptr(), s), None)
        else:
            g = self.texture_grad_dense(with_losses=not deferred)   # fold + regularisers (beside the geometry backward when not deferred)
            if self._dp_buf is not None and not deferred:
                allreduce_fn(self._dp_buf)                 # texture gradient + gradient slab in one collective
            else:
                allreduce_fn(g.reshape(-1))
            self._ck(self.L.vhap_tex_apply_grad(self.ctx, self.tex_extra.data_ptr(), g.data_ptr(), self.tex_m.data_ptr(), self.tex_v.data_ptr(),
                                                self._lr("tex"), self.step_count, C.byref(cs), s), None)
        if deferred:
            if reg_loss:
                self._ck(self.L.vhap_tex_reg_loss(self.ctx, self.tex_extra.data_ptr(), C.byref(cs), s), None)
            self.L.vhap_tex_defer(self.ctx, 0)

    def adam_step(self, allreduce_fn=None, texture=True, tex_comm=None):
        """torch.optim.Adam.step() over the parameter groups of the current stage (dense rows, tracker.py:1284-1293,210)."""
        opt = opt_dict_for(self.stage)
        self.step_count += 1
        s = self._stream()
        # parameter groups and learning rates of the stage (config.adam_param_lrs, pinned against the reference's get_train_parameters +
        # configure_optimizer by tests/test_config_golden.py); slab order
        lrs = adam_param_lrs(self.stage, self.cfg.lr, self._host_lr_scale(), self.cfg.calibrated)
        groups = [n for n in ("shape", "static_offset", "lights", "focal_length", "expr", "rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose")
                  if n in lrs]
        tex = texture and opt["texture"]
        dp = allreduce_fn is not None or tex_comm is not None
        if tex and dp:
            self.tex_update(allreduce_fn, tex_comm=tex_comm)
        if allreduce_fn is not None and not (tex and tex_comm is None and self._dp_buf is not None):
            allreduce_fn(self.grad)
        if groups:
            off = np.asarray([self.layout[g][0] for g in groups], np.int64)
            ln = np.asarray([self.layout[g][1] for g in groups], np.int64)
            lr = np.asarray([lrs[g] for g in groups], np.float32)
            hp = lambda a: a.ctypes.data_as(C.c_void_p)
            self._ck(self.L.vhap_adam_multi(self.ctx, self.slab.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                            len(groups), hp(off), hp(ln), hp(lr), self.step_count, s), None)
        if tex and not dp:
            self.tex_update(None)                          # issued last: only its join and the loss bookkeeping are on this stream

    def step(self, batch: Batch) -> torch.Tensor:
        """One optimisation iteration (tracker.py:1418-1435): zero_grad, energy + backward, Adam."""
        self.zero_grad()
        losses = self.energy(batch, backward=True, training=True)
        self.adam_step()
        self.global_step += 1
        return losses

    def optimize_stage(self, stage_name: str, batches, lr_scale: float = 1.0, per_sample: bool = False, on_step=None):
        """FlameTracker.optimize_stage (tracker.py:1391-1416): fresh Adam state, then the schedule of config.stage_schedule (pinned against
        the reference's own loop by tests/test_config_golden.py) -- eager steps; the learning-rate scale changes between epochs without
        touching the Adam moments, like torch's ExponentialLR.  `on_step(i, losses)` is called after every iteration."""
        from .config import stage_schedule
        self.set_stage(stage_name, lr_scale)
        for i, (b, scale) in enumerate(stage_schedule(stage_name, len(batches), lr_scale, per_sample)):
            self.lr_scale = scale
            losses = self.step(batches[b])
            if on_step is not None:
                on_step(i, losses)

    # ------------------------------------------------------------------ CUDA-graph replay of whole steps
    def graph_begin(self, batches, body=None, reduce_fn=None, allreduce_fn=None, world=1, pipelined=True, global_Bs=None, tex_comm=None):
        """Capture one optimisation step per (batch, texture ping-pong parity) as CUDA graphs.  All step-dependent values
        (Adam step, RNG step) live in device memory (vhap_step_counters), so the graphs are replayable indefinitely.
        Data parallel: pass reduce_fn / allreduce_fn / world (the NCCL collectives are captured); `body(batch)` overrides the captured
        step entirely (not pipelined).
        pipelined (stages that optimise the texture photometrically): the texture update of step k -- one HBM-streaming pass, plus the
        50 MB all-reduce when data parallel -- is the tail of the step, while the first third of the next step (FLAME, rasteriser,
        pixel pools) never reads the texture.  The replayed graph therefore starts with the update of the PREVIOUS step on a side
        stream and joins it right before the shading pass; the first graph_step runs an eager prologue, graph_end flushes the last
        update.  Same arithmetic in the same order on every buffer, only the schedule differs."""
        s = self._stream()
        if tex_comm is not None and not getattr(tex_comm, "peer", False):
            self._shard_bufs(tex_comm)                        # allocations must not happen inside capture
        elif allreduce_fn is not None and self.tex_grad_dense is None:
            self.texture_grad_dense(with_losses=False)       # one-time eager set-up (allocation, stream hand-over) must not happen inside capture
        self._ck(self.L.vhap_step_counters(self.ctx, 1, self.step_count + 1, self.global_step, s))
        self._graph_live = True
        self.lr_scale = self._lr_scale                       # -> device
        torch.cuda.synchronize(self.dev)
        opt = opt_dict_for(self.stage)
        self._pipe = bool(pipelined and body is None and self.stage is not None and self.stage.photometric and opt["texture"])
        self._tex_split = bool(self._pipe and tex_comm is not None and getattr(tex_comm, "peer", False))
        self._hooks = (reduce_fn, allreduce_fn, world, tex_comm)
        self._graph_batches = list(batches)
        # true global batch size per staged batch (uneven shards: not B * world); parallel.DataParallelStep passes the all-reduced values
        self._graph_gB = {id(b): (g if global_Bs is not None else b.B * world) for b, g in zip(self._graph_batches, global_Bs or [0] * len(self._graph_batches))}
        self._primed = False
        self._graphs, self._graph_events = {}, []
        if self._pipe and not hasattr(self, "_tex_stream"):
            self._tex_stream = torch.cuda.Stream(self.dev)
        parity0 = self.L.vhap_get_cur_mip(self.ctx)
        step_save, gstep_save = self.step_count, self.global_step
        # the capture stream has a higher priority than the side stream of the deferred texture update (and the library's aux streams):
        # the step's latency-bound kernel chain is scheduled ahead of the remaining CTAs of the machine-filling texture pass
        side = torch.cuda.Stream(self.dev, priority=-1)
        for bi, batch in enumerate(batches):
            for par in (0, 1):
                self.L.vhap_set_cur_mip(self.ctx, par)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    if body is not None:
                        body(batch)
                        self._ck(self.L.vhap_step_advance(self.ctx, self._stream()))
                    else:
                        self._step_body(batch, deferred_tex=self._pipe, texture_now=not self._pipe)
                self._graphs[(bi, par)] = g
        self.step_count, self.global_step = step_save, gstep_save       # capture executed nothing
        self._parity = parity0
        self.L.vhap_set_cur_mip(self.ctx, parity0)

    def _step_body(self, batch, deferred_tex: bool, texture_now: bool = False):
        """one step with device-resident counters: [texture update of the previous step on a side stream] + zero_grad + energy +
        backward + Adam of the small groups (+ the texture update right away when texture_now) + counter advance"""
        reduce_fn, allreduce_fn, world, tex_comm = self._hooks
        if deferred_tex:
            cur = torch.cuda.current_stream(self.dev)
            e0, e1 = torch.cuda.Event(), torch.cuda.Event()
            e0.record(cur)
            self._tex_stream.wait_event(e0)
            e2 = torch.cuda.Event()
            cs = self._c_stage(True)
            with torch.cuda.stream(self._tex_stream):
                self.tex_update(allreduce_fn, deferred=True, reg_loss=False, tex_comm=tex_comm, second_half=self._tex_split)
                e1.record(self._tex_stream)               # the new pyramid is complete: what the shading pass waits for
                # regulariser loss VALUES of the updated texture: only the loss vector needs them, joined at the end of the step
                self._ck(self.L.vhap_tex_reg_loss(self.ctx, self.tex_extra.data_ptr(), C.byref(cs), self._stream()), None)
                e2.record(self._tex_stream)
            self._graph_events += [e0, e1, e2]
            self.L.vhap_set_render_wait_event(self.ctx, C.c_void_p(e1.cuda_event))    # joined right before the shading pass
        self.zero_grad()
        self.energy(batch, backward=True, training=True, global_B=self._graph_gB.get(id(batch), batch.B * world), reduce_fn=reduce_fn)
        split = self._tex_split and not texture_now
        if split:
            # peer-memory texture update, first half (fold -> barrier -> in-switch band reduction): starts as soon as the texel gradients are
            # complete, on the library's bulk stream beside the geometry backward / parameter all-reduce / Adam of THIS step; the second half
            # (band Adam, multicast store, rebuild) opens the next step
            self._ck(self.L.vhap_dp_tex_part1(self.ctx, self.tex_extra.data_ptr(), self._stream()))
        self.adam_step(allreduce_fn=allreduce_fn, texture=texture_now, tex_comm=tex_comm)
        if split:
            self._ck(self.L.vhap_dp_tex_join(self.ctx, self._stream()))
        if deferred_tex:
            torch.cuda.current_stream(self.dev).wait_event(e2)
            self._ck(self.L.vhap_assemble_losses(self.ctx, C.byref(cs), self.losses.data_ptr(), self._stream()), None)
        self._ck(self.L.vhap_step_advance(self.ctx, self._stream()))

    def graph_step(self, bi: int):
        if self._pipe and not self._primed:
            # prologue of the pipeline: the first step eagerly, leaving its texture update pending for the first replay
            cs = self._c_stage(True)
            self._ck(self.L.vhap_tex_reg_loss(self.ctx, self.tex_extra.data_ptr(), C.byref(cs), self._stream()), None)   # its loss vector
            self._step_body(self._graph_batches[bi], deferred_tex=False)
            self.step_count -= 1                           # (_step_body's adam_step counted on the host; counted below like a replay)
            self._primed = True
        else:
            self._graphs[(bi, self._parity)].replay()
            self._parity ^= 1
        self.step_count += 1
        self.global_step += 1

    def _flush_pending_tex(self):
        """pipelined replay: apply the texture update that is still pending from the last replayed step (eagerly, on the current
        stream).  The next graph_step then starts the pipeline again with its eager prologue."""
        if getattr(self, "_pipe", False) and self._primed:
            self.L.vhap_set_cur_mip(self.ctx, self._parity)     # the replays did not touch the host-side ping-pong index
            self.tex_update(self._hooks[1], deferred=True, tex_comm=self._hooks[3], second_half=self._tex_split)
            self._parity ^= 1
            self._primed = False

    def graph_end(self):
        torch.cuda.synchronize(self.dev)
        self._flush_pending_tex()
        self.L.vhap_set_cur_mip(self.ctx, self._parity)
        torch.cuda.synchronize(self.dev)
        self._ck(self.L.vhap_step_counters(self.ctx, 0, 0, 0, self._stream()))
        self._graph_live = False
        self._graphs, self._graph_events = {}, []
        self._primed = False
        self._pipe = False
        self._tex_split = False

    # ------------------------------------------------------------------ logging planes (render_out dict)
    def render_planes(self, batch: Batch, training=False) -> Dict[str, torch.Tensor]:
        self._flush_pending_tex()
        self.L.vhap_set_want_planes(self.ctx, 1)
        self.energy(batch, backward=False, training=training)
        self.L.vhap_set_want_planes(self.ctx, 0)
        out = {}
        for which, name in ((0, "rgba"), (1, "rgba_pre"), (2, "albedo"), (3, "normal"), (4, "diffuse"), (5, "cid")):
            t = torch.empty(batch.B, batch.H, batch.W, 4, dtype=torch.float32, device=self.dev)
            self._ck(self.L.vhap_get_plane(self.ctx, which, t.data_ptr(), self._stream()), None)
            out[name] = t
        return out

    def loss_dict(self) -> Dict[str, float]:
        v = self.losses.cpu().numpy()
        return {n: float(v[i]) for i, n in enumerate(_lib.LOSS_NAMES)}
