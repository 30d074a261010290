"""Drop-in classes with the reference's names, constructor arguments, method signatures and return dictionaries, backed by
libvhap_b200.so.  Selected in the reference by `cfg.render.backend == 'b200'`:

    FlameHead      -> B200FlameHead      (vhap/model/flame.py:64-646: forward)
    NVDiffRenderer -> B200Renderer       (vhap/util/render_nvdiffrast.py:56-484: rasterize, render_rgba, world_to_ndc, clear_cache)

Both are differentiable through torch.autograd.Functions whose backward calls the analytic CUDA backward kernels.  The fused
fast path that replaces the whole `compute_energy` + `backward()` + `Adam.step()` iteration is `vhap_b200.engine.Engine`.
Unsupported options raise the same exception types as the reference for unknown settings (NotImplementedError / ValueError)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .config import EngineConfig
from .engine import Engine, Batch, _ptr
from .flame_model import FlameModelData


class _Core:
    """one shared engine context per (model, device): owns the static tables and scratch"""
    _inst = {}

    @classmethod
    def get(cls, model: FlameModelData, tex_size: int, device="cuda:0") -> Engine:
        key = (id(model), tex_size, str(device))
        if key not in cls._inst:
            cfg = EngineConfig(tex_resolution=tex_size)
            cls._inst[key] = Engine(model, cfg, 1, device=device)
        return cls._inst[key]


class _FlameFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, shape, expr, rotation, neck, jaw, eyes, translation, static_offset):
        B = expr.shape[0]
        f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
        tens = dict(shape=f32(shape[0]), expr=f32(expr), rotation=f32(rotation), neck=f32(neck), jaw=f32(jaw), eyes=f32(eyes),
                    translation=f32(translation), off=f32(static_offset))
        ts = torch.arange(B, dtype=torch.int32, device=eng.dev)
        eng.reserve(max(B, 1), 16, 16)
        fb = _lib.FrameBatch(B, 16, 16, 0, ts.data_ptr(), None, None, None, None, None, None, 0)
        p = _lib.Params(_ptr(tens["shape"]), _ptr(tens["expr"]), _ptr(tens["rotation"]), _ptr(tens["neck"]), _ptr(tens["jaw"]), _ptr(tens["eyes"]),
                        _ptr(tens["translation"]), _ptr(tens["off"]), _ptr(eng.p["lights"]), _ptr(eng.p["focal_length"]), None, B)
        verts = torch.empty(B, eng.V, 3, device=eng.dev); cano = torch.empty(B, eng.V, 3, device=eng.dev); lmks = torch.empty(B, 70, 3, device=eng.dev)
        eng._ck(eng.L.vhap_flame_forward(eng.ctx, C.byref(p), C.byref(fb), verts.data_ptr(), cano.data_ptr(), lmks.data_ptr(), eng._stream()))
        ctx.eng, ctx.tens, ctx.ts, ctx.B = eng, tens, ts, B
        ctx.has_off = static_offset is not None
        return verts, cano, lmks

    @staticmethod
    def backward(ctx, g_verts, g_cano, g_lmks):
        eng, t, B = ctx.eng, ctx.tens, ctx.B
        z = lambda like: torch.zeros_like(like)
        g = dict(shape=z(t["shape"]), expr=z(t["expr"]), rotation=z(t["rotation"]), neck=z(t["neck"]), jaw=z(t["jaw"]), eyes=z(t["eyes"]),
                 translation=z(t["translation"]), off=z(t["off"]) if ctx.has_off else None)
        fb = _lib.FrameBatch(B, 16, 16, 0, ctx.ts.data_ptr(), None, None, None, None, None, None, 0)
        p = _lib.Params(_ptr(t["shape"]), _ptr(t["expr"]), _ptr(t["rotation"]), _ptr(t["neck"]), _ptr(t["jaw"]), _ptr(t["eyes"]),
                        _ptr(t["translation"]), _ptr(t["off"]), _ptr(eng.p["lights"]), _ptr(eng.p["focal_length"]), None, B)
        cg = _lib.Grads(_ptr(g["shape"]), _ptr(g["expr"]), _ptr(g["rotation"]), _ptr(g["neck"]), _ptr(g["jaw"]), _ptr(g["eyes"]), _ptr(g["translation"]),
                        _ptr(g["off"]), None, None, None)
        # the forward state (pose chain, v_posed) lives in the ctx scratch: recompute it, then run the backward
        eng._ck(eng.L.vhap_flame_forward(eng.ctx, C.byref(p), C.byref(fb), None, None, None, eng._stream()))
        gv = g_verts.to(torch.float32).contiguous() if g_verts is not None else None
        gl = g_lmks.to(torch.float32).contiguous() if g_lmks is not None else None
        eng._ck(eng.L.vhap_flame_backward(eng.ctx, C.byref(p), C.byref(fb), _ptr(gv), _ptr(gl), C.byref(cg), eng._stream()))
        if g_cano is not None and g["off"] is not None:       # verts_cano = v_shaped: d/d offset = identity summed over the batch
            g["off"] += g_cano.sum(0).reshape(-1, 3).reshape(g["off"].shape)
        g_shape = (g["shape"] / B)[None].expand(B, -1)          # the tracker passes shape[None].expand(B, -1) (tracker.py:221)
        return (None, g_shape, g["expr"], g["rotation"], g["neck"], g["jaw"], g["eyes"], g["translation"],
                None if g["off"] is None else g["off"].reshape(1, -1, 3))


class B200FlameHead(torch.nn.Module):
    """FlameHead (vhap/model/flame.py:64) on the B200 engine.  `model` defaults to the synthetic FLAME (real topology, seeded
    bases); pass FlameModelData.from_pkl(...) for the licensed model."""

    def __init__(self, shape_params, expr_params, model: Optional[FlameModelData] = None, add_teeth=True, remove_lip_inside=False,
                 face_clusters=(), device="cuda:0", **unused):
        super().__init__()
        if remove_lip_inside:
            raise NotImplementedError("remove_lip_inside is not supported by the b200 backend")
        self.model = model if model is not None else FlameModelData.synthetic(shape_params, expr_params, add_teeth=add_teeth)
        self.n_shape_params, self.n_expr_params = shape_params, expr_params
        self.eng = _Core.get(self.model, 256, device)
        dev = self.eng.dev
        self.register_buffer("faces", torch.as_tensor(self.model.faces.astype(np.int64), device=dev), persistent=False)
        self.register_buffer("verts_uvs", torch.as_tensor(self.model.verts_uv, device=dev), persistent=False)
        self.register_buffer("textures_idx", torch.as_tensor(self.model.faces_uv.astype(np.int64), device=dev), persistent=False)
        self.register_buffer("v_template", torch.as_tensor(self.model.v_template, device=dev), persistent=False)
        self.mask = self.model          # get_vid_by_region / get_fid_by_region / fid2cid live on the model data

    def forward(self, shape, expr, rotation, neck, jaw, eyes, translation, zero_centered_at_root_node=False, return_landmarks=True,
                return_verts_cano=False, static_offset=None, dynamic_offset=None):
        if dynamic_offset is not None:
            raise NotImplementedError("dynamic_offset is not supported by the b200 backend (base.py:69 default is off)")
        if zero_centered_at_root_node:
            raise NotImplementedError("zero_centered_at_root_node is not supported by the b200 backend")
        off = None if static_offset is None else static_offset.reshape(-1, 3)
        verts, cano, lmks = _FlameFn.apply(self.eng, shape, expr, rotation, neck, jaw, eyes, translation, off)
        ret = [verts]
        if return_verts_cano:
            ret.append(cano)
        if return_landmarks:
            ret.append(lmks)
        return ret if len(ret) > 1 else ret[0]


# constant factors of the first three SH bands (render_nvdiffrast.py:83-96)
_SH_CONST = (0.28209479177387814, 1.0233267079464885, 1.0233267079464885, 1.0233267079464885, 0.8580855308097834, 0.8580855308097834,
             0.8580855308097834, 0.4290427654048917, 0.24770795610037571)


def _sh_shade(normal: torch.Tensor, lights: torch.Tensor) -> torch.Tensor:
    """get_SH_shading (render_nvdiffrast.py:19-53) on a [B,H,W,3] normal plane, lights [9,3]: a tiny torch op that keeps the
    autograd edge to `lights` -- used for `diffuse_detach_normal` (render_nvdiffrast.py:403), the input of reg_diffuse (tracker.py:547-550)."""
    x, y, z = normal[..., 0], normal[..., 1], normal[..., 2]
    basis = torch.stack([torch.ones_like(x), x, y, z, x * y, x * z, y * z, x * x - y * y, 3 * z * z - 1], -1)
    basis = basis * torch.tensor(_SH_CONST, dtype=normal.dtype, device=normal.device)
    return basis @ lights.to(normal.dtype).reshape(9, 3)


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rnd, verts_clip, v_normal, tex, lights, batch, cs):
        eng = rnd.eng
        B, H, W = batch.B, batch.H, batch.W
        eng.tex_extra.copy_(tex.detach().to(torch.float32).reshape(-1))
        eng.rebuild_texture()
        eng.p["lights"].copy_(lights.detach().reshape(-1))
        cp = eng._c_params()
        clip = verts_clip.detach().to(torch.float32).contiguous(); vn = v_normal.detach().to(torch.float32).contiguous()
        eng.L.vhap_set_want_planes(eng.ctx, 1)
        eng._ck(eng.L.vhap_render_photometric(eng.ctx, C.byref(cp), C.byref(batch.c), C.byref(cs), clip.data_ptr(), vn.data_ptr(), eng.losses.data_ptr(),
                                              None, None, None, None, eng._stream()))
        eng.L.vhap_set_want_planes(eng.ctx, 0)
        planes = []
        for which in (0, 2, 3, 4, 5):
            t = torch.empty(B, H, W, 4, device=eng.dev)
            eng._ck(eng.L.vhap_get_plane(eng.ctx, which, t.data_ptr(), eng._stream()))
            planes.append(t)
        ctx.rnd, ctx.batch, ctx.cs, ctx.clip, ctx.vn = rnd, batch, cs, clip, vn
        ctx.mark_non_differentiable(*planes[1:])
        return tuple(planes)

    @staticmethod
    def backward(ctx, g_rgba, *unused):
        eng, batch, cs = ctx.rnd.eng, ctx.batch, ctx.cs
        B, V = batch.B, eng.V
        cp = eng._c_params()
        # recompute the forward state (pass A colours, pools, pair cache) for these inputs, then run the adjoint
        eng._ck(eng.L.vhap_render_photometric(eng.ctx, C.byref(cp), C.byref(batch.c), C.byref(cs), ctx.clip.data_ptr(), ctx.vn.data_ptr(),
                                              eng.losses.data_ptr(), None, None, None, None, eng._stream()))
        g_clip = torch.zeros(B, V, 4, device=eng.dev); g_vn = torch.zeros(B, V, 3, device=eng.dev); g_l = torch.zeros(27, device=eng.dev)
        gtex = eng.L.vhap_tex_grad_ptr(eng.ctx)
        g = g_rgba.to(torch.float32).contiguous()
        eng._ck(eng.L.vhap_render_rgba_backward(eng.ctx, C.byref(cp), C.byref(batch.c), C.byref(cs), g.data_ptr(), g_clip.data_ptr(), g_vn.data_ptr(),
                                                g_l.data_ptr(), gtex, eng._stream()))
        import copy
        save = eng.cfg
        eng.cfg = copy.deepcopy(save); eng.cfg.w.reg_tex_tv = None; eng.cfg.w.reg_tex_res_clusters = None
        g_tex = eng.texture_grad_dense().clone()
        eng.cfg = save
        return None, g_clip, g_vn, g_tex, g_l.reshape(9, 3), None, None


class B200Renderer(torch.nn.Module):
    """NVDiffRenderer (vhap/util/render_nvdiffrast.py:56) on the B200 engine."""

    def __init__(self, use_opengl=False, lighting_type="SH", lighting_space="world", disturb_rate_fg=0.5, disturb_rate_bg=0.5, fid2cid=None,
                 shade_smooth=True, model: Optional[FlameModelData] = None, tex_size=2048, device="cuda:0"):
        super().__init__()
        self.backend = "b200"
        if lighting_type != "SH":
            raise NotImplementedError(f"Unknown lighting type: {lighting_type}")          # render_nvdiffrast.py:346
        if lighting_space != "world":
            raise NotImplementedError(f"Unknown lighting space: {lighting_space}")        # render_nvdiffrast.py:379
        self.lighting_type, self.lighting_space = lighting_type, lighting_space
        self.disturb_rate_fg, self.disturb_rate_bg = disturb_rate_fg, disturb_rate_bg
        self.model = model if model is not None else FlameModelData.synthetic()
        self.eng = _Core.get(self.model, tex_size, device)
        self.eng.cfg.render.disturb_rate_fg, self.eng.cfg.render.disturb_rate_bg = disturb_rate_fg, disturb_rate_bg
        self.fragment_cache = None

    def clear_cache(self):
        self.fragment_cache = None

    def _check_topology(self, faces=None, verts_uv=None, faces_uv=None):
        """the engine renders the mesh it was built with (static tables on the device): a caller passing another topology gets an error
        instead of a silently different picture.  verts_uv arrives with v already flipped by the tracker (tracker.py:315-316)."""
        m = self.model
        if faces is not None and (tuple(faces.shape[-2:]) != m.faces.shape or not np.array_equal(faces.reshape(-1, 3).cpu().numpy(), m.faces)):
            raise ValueError("faces differ from the FLAME topology this B200Renderer was built with")
        if faces_uv is not None and (tuple(faces_uv.shape[-2:]) != m.faces_uv.shape or not np.array_equal(faces_uv.reshape(-1, 3).cpu().numpy(), m.faces_uv)):
            raise ValueError("faces_uv differ from the FLAME topology this B200Renderer was built with")
        if verts_uv is not None:
            vuv = verts_uv.reshape(-1, 2).detach().cpu().numpy()
            ref = m.verts_uv.astype(np.float32).copy()
            ref[:, 1] = 1.0 - ref[:, 1]
            if vuv.shape != ref.shape or not np.allclose(vuv, ref, atol=1e-6):
                raise ValueError("verts_uv differ from the (v-flipped) FLAME texture coordinates this B200Renderer was built with")

    # ---- camera (render_nvdiffrast.py:117-214); tiny host-side tensor algebra, kept in torch like the reference
    def projection_from_intrinsics(self, K, image_size, near=0.1, far=10.0):
        B = K.shape[0]
        h, w = image_size
        if K.shape[-2:] == (3, 3):
            fx, fy, cx, cy = K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2]
        elif K.shape[-1] == 4:
            fx, fy, cx, cy = K[..., 0], K[..., 1], K[..., 2], K[..., 3]
        else:
            raise ValueError(f"Expected K to be (N, 3, 3) or (N, 4) but got: {K.shape}")
        proj = torch.zeros([B, 4, 4], device=K.device, dtype=K.dtype)
        proj[:, 0, 0] = fx * 2 / w; proj[:, 1, 1] = fy * 2 / h
        proj[:, 0, 2] = (w - 2 * cx) / w; proj[:, 1, 2] = (h - 2 * cy) / h
        proj[:, 2, 2] = -(far + near) / (far - near); proj[:, 2, 3] = -2 * far * near / (far - near); proj[:, 3, 2] = -1
        return proj

    def world_to_camera(self, vtx, RT):
        # [B,3,4] extrinsics give [B,V,3], [B,4,4] give [B,V,4] (render_nvdiffrast.py:162-179)
        posw = torch.cat([vtx, torch.ones_like(vtx[..., :1])], -1) if vtx.shape[-1] == 3 else vtx
        return torch.bmm(posw, RT.transpose(-1, -2))

    def camera_to_clip(self, vtx, K, image_size):
        proj = self.projection_from_intrinsics(K, image_size)
        posw = torch.cat([vtx, torch.ones_like(vtx[..., :1])], -1) if vtx.shape[-1] == 3 else vtx
        if proj.shape[0] < posw.shape[0]:
            proj = proj.expand(posw.shape[0], -1, -1)
        return torch.bmm(posw, proj.transpose(-1, -2))

    def world_to_ndc(self, vtx, RT, K, image_size, flip_y=False):
        clip = self.camera_to_clip(self.world_to_camera(vtx, RT), K, image_size)
        ndc = clip[:, :, :3] / clip[:, :, 3:]
        if flip_y:
            ndc = ndc * torch.tensor([1.0, -1.0, 1.0], device=ndc.device, dtype=ndc.dtype)
        return ndc

    # ---- rasterize (render_nvdiffrast.py:216-260)
    def rasterize(self, verts, faces, RT, K, image_size, use_cache=False, require_grad=False):
        self._check_topology(faces)
        verts_camera = self.world_to_camera(verts, RT)
        verts_clip = self.camera_to_clip(verts_camera, K, image_size)
        if not use_cache or self.fragment_cache is None:
            eng = self.eng
            B = verts.shape[0]
            H, W = image_size
            eng.reserve(B, H, W)
            clip = verts_clip.detach().to(torch.float32).contiguous()
            rast = torch.empty(B, H, W, 4, device=eng.dev); db = torch.empty(B, H, W, 4, device=eng.dev)
            eng._ck(eng.L.vhap_rasterize(eng.ctx, clip.data_ptr(), B, H, W, None, rast.data_ptr(), db.data_ptr(), 0, eng._stream()))
            self.fragment_cache = (rast, db)
        rast_out, rast_out_db = self.fragment_cache
        return {"rast_out": rast_out, "rast_out_db": rast_out_db, "verts": verts, "verts_camera": verts_camera[..., :3], "verts_clip": verts_clip}

    def compute_v_normals(self, verts, faces=None):
        eng = self.eng
        B = verts.shape[0]
        out = torch.empty(B, eng.V, 3, device=eng.dev)
        v = verts.detach().to(torch.float32).contiguous()
        eng.reserve(max(B, 1), 16, 16)
        eng._ck(eng.L.vhap_vertex_normals(eng.ctx, v.data_ptr(), B, out.data_ptr(), eng._stream()))
        return out

    # ---- render_rgba (render_nvdiffrast.py:354-484)
    def render_rgba(self, rast_dict, verts, faces, verts_uv, faces_uv, tex, lights, background_color=[1.0, 1.0, 1.0],
                    align_texture_except_fid=None, align_boundary_except_vid=None, enable_disturbance=False):
        eng = self.eng
        self._check_topology(faces, verts_uv, faces_uv)
        verts_clip = rast_dict["verts_clip"]
        B = verts_clip.shape[0]
        H, W = rast_dict["rast_out"].shape[1:3]
        if isinstance(background_color, torch.Tensor):
            tgt = torch.cat([background_color, torch.zeros_like(background_color[..., :1])], -1).to(torch.float16).contiguous()
            eng.cfg.render.background_train = eng.cfg.render.background_eval = "target"
        elif isinstance(background_color, (list, tuple)):
            c = list(background_color)
            tgt = torch.zeros(B, H, W, 4, dtype=torch.float16, device=eng.dev)
            mode = "white" if c == [1, 1, 1] or c == [1.0, 1.0, 1.0] else ("black" if c == [0, 0, 0] or c == [0.0, 0.0, 0.0] else None)
            if mode is None:
                raise NotImplementedError(f"constant background {c}: only white / black like the tracker (tracker.py:296-302)")
            eng.cfg.render.background_train = eng.cfg.render.background_eval = mode
        else:
            raise ValueError(f"Unknown background type: {type(background_color)}")       # render_nvdiffrast.py:418
        # stage masks for this call
        from .config import StageConfig
        st = StageConfig(name="render_rgba", photometric=True)
        eng.stage = st if enable_disturbance or align_texture_except_fid is not None or align_boundary_except_vid is not None else None
        face_flags = np.zeros(eng.F, np.uint8); vert_flags = np.zeros(eng.V, np.uint8)
        if align_texture_except_fid is not None:
            face_flags[np.asarray(align_texture_except_fid.cpu() if torch.is_tensor(align_texture_except_fid) else align_texture_except_fid)] = 1
        if align_boundary_except_vid is not None:
            vert_flags[np.asarray(align_boundary_except_vid.cpu() if torch.is_tensor(align_boundary_except_vid) else align_boundary_except_vid)] = 1
        hp = lambda a: a.ctypes.data_as(C.c_void_p)
        eng._ck(eng.L.vhap_set_stage_masks(eng.ctx, hp(face_flags), hp(vert_flags), None, None, None, None, 0, None))
        if not enable_disturbance:
            eng.cfg.render.disturb_rate_fg = eng.cfg.render.disturb_rate_bg = None
        else:
            eng.cfg.render.disturb_rate_fg, eng.cfg.render.disturb_rate_bg = self.disturb_rate_fg, self.disturb_rate_bg
        cs = eng._c_stage(training=eng.stage is not None)
        ts = torch.zeros(B, dtype=torch.int32, device=eng.dev)
        lm = torch.zeros(B, 68, 3, device=eng.dev)
        eng.reserve(B, H, W)
        batch = Batch(B, H, W, ts, tgt, lm)
        v_normal = _NormalsFn.apply(eng, rast_dict["verts"])
        tex_chw = tex[0] if tex.dim() == 4 else tex
        rgba, albedo, normal, diffuse, cid = _RenderFn.apply(self, verts_clip, v_normal, tex_chw, lights.reshape(9, 3), batch, cs)
        # diffuse_detach_normal keeps its gradient w.r.t. the lights (only the normal is detached, render_nvdiffrast.py:402-403):
        # SH shading of the (non-differentiable) normal plane in torch; background pixels have normal 0 like the reference's
        return {"albedo": albedo[..., :3], "normal": normal[..., :3], "diffuse": diffuse[..., :3],
                "diffuse_detach_normal": _sh_shade(normal[..., :3].detach(), lights.reshape(9, 3)),
                "rgba": rgba, "aa": albedo[..., 3:4].detach().expand(-1, -1, -1, 3).contiguous(), "cid": cid[..., :1].long()}

    # ---- render_rgba_vis (render_nvdiffrast.py:486-567): the visualisation render of the viewer / editor / NeRF export
    def render_rgba_vis(self, verts, faces, RT, K, image_size, background_color=[1.0, 1.0, 1.0], v_color=None, verts_uv=None, faces_uv=None,
                        tex=None, lights=None):
        """same keys as the reference ('albedo', 'normal', 'diffuse', 'rgba', 'verts_clip'), image orientation, no gradients (a
        visualisation call).  Built from this class's rasterize + render_rgba: no texture -> albedo 1 (:531), no lights -> lighting_type
        'constant' (diffuse 1, :332-333, realised as the SH vector (1 / c0, 0, ...)), normal / diffuse carry the background outside the
        mesh (:553-554).  Constant backgrounds other than white / black go through the fp16 image path (rounded to fp16)."""
        if v_color is not None:
            raise NotImplementedError("render_rgba_vis: per-vertex colours (render_nvdiffrast.py:526-530) are not an input of the B200 engine")
        eng = self.eng
        with torch.no_grad():
            rd = self.rasterize(verts, faces, RT, K, image_size)
            B = verts.shape[0]
            H, W = image_size
            if tex is None or verts_uv is None or faces_uv is None:
                tex_, verts_uv, faces_uv = torch.ones(3, eng.T, eng.T, device=eng.dev), None, None
            else:
                tex_ = tex
            if lights is None:
                lights_ = torch.zeros(9, 3, device=eng.dev); lights_[0] = 1.0 / _SH_CONST[0]
            else:
                lights_ = lights.reshape(9, 3)
            bg = background_color
            if isinstance(bg, (list, tuple)) and [float(v) for v in bg] not in ([1.0, 1.0, 1.0], [0.0, 0.0, 0.0]):
                bg = torch.tensor([float(v) for v in bg], device=eng.dev).expand(B, H, W, 3)
            out = self.render_rgba(rd, verts, faces, verts_uv, faces_uv, tex_, lights_, bg)
            fg = (rd["rast_out"][..., 3:4] > 0).flip(1)
            if isinstance(bg, torch.Tensor):
                bg3 = bg.to(torch.float16).to(torch.float32)
            else:
                bg3 = torch.tensor([float(v) for v in bg], device=eng.dev).expand(B, H, W, 3)
            return {"albedo": out["albedo"], "normal": torch.where(fg, out["normal"], bg3), "diffuse": torch.where(fg, out["diffuse"], bg3),
                    "rgba": out["rgba"].detach(), "verts_clip": rd["verts_clip"]}


class _NormalsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, verts):
        B = verts.shape[0]
        v = verts.detach().to(torch.float32).contiguous()
        out = torch.empty(B, eng.V, 3, device=eng.dev)
        eng.reserve(max(B, 1), 16, 16)
        eng._ck(eng.L.vhap_vertex_normals(eng.ctx, v.data_ptr(), B, out.data_ptr(), eng._stream()))
        ctx.eng, ctx.v = eng, v
        return out

    @staticmethod
    def backward(ctx, g):
        eng, v = ctx.eng, ctx.v
        gv = torch.zeros_like(v)
        gg = g.to(torch.float32).contiguous()
        eng._ck(eng.L.vhap_vertex_normals_backward(eng.ctx, v.data_ptr(), gg.data_ptr(), v.shape[0], gv.data_ptr(), eng._stream()))
        return None, gv
