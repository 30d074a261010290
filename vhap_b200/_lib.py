"""ctypes binding of libvhap_b200.so (include/vhap_b200.h).  No torch types cross the boundary: only raw device
pointers, sizes and the CUDA stream handle.  Fails loudly if the library is missing -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_SO = Path(os.environ.get("VHAP_B200_SO", _HERE / "libvhap_b200.so"))     # override: experiment variants built by build_ext.py (dev only)

c_float_p = C.c_void_p      # device pointers are passed as integers
c_int_p = C.c_void_p

N_LOSS = 24
LOSS_NAMES = ["total", "lmk", "photo", "reg_shape", "reg_expr", "reg_joint", "smooth_pose", "smooth_joint", "smooth_expr",
              "reg_tex_tv", "reg_tex_res_clusters", "reg_diffuse", "reg_light", "reg_offset", "reg_offset_lap", "reg_offset_rigid",
              "n_fg", "abs_err"]


class MeshDesc(C.Structure):
    _fields_ = [("V", C.c_int32), ("F", C.c_int32), ("VT", C.c_int32), ("K", C.c_int32), ("n_shape", C.c_int32), ("n_lmk", C.c_int32),
                ("n_clusters", C.c_int32)] + [(n, C.c_void_p) for n in (
                    "v_template_host", "shapedirs_host", "posedirs_host", "J_regressor_host", "lbs_weights_host", "faces_host", "faces_uv_host",
                    "verts_uv_host", "lmk_faces_host", "lmk_bary_host", "adj_opp_host", "fid2cid_host", "vf_indptr_host", "vf_faces_host",
                    "lap_indptr_host", "lap_indices_host", "lap_values_host")]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("shape", "expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation", "static_offset",
                                           "lights", "focal_length", "tex_extra")] + [("n_timesteps", C.c_int32)]


class Grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("shape", "expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation", "static_offset",
                                           "lights", "focal_length", "tex_grad_pyramid")]


class FrameBatch(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("target_format", C.c_int32), ("timesteps", C.c_void_p), ("target", C.c_void_p), ("lmk2d", C.c_void_p),
                ("RT", C.c_void_p), ("K", C.c_void_p), ("geo", C.c_void_p), ("geo_timesteps", C.c_void_p), ("n_geo", C.c_int32)]


class StageCfg(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "w_landmark", "w_photo", "w_reg_shape", "w_reg_expr", "w_reg_neck", "w_reg_jaw", "w_reg_eyes", "w_reg_tex_tv", "w_reg_tex_res",
        "w_reg_diffuse", "w_reg_light", "w_reg_offset", "w_reg_offset_lap", "w_reg_offset_rigid", "w_smooth_trans", "w_smooth_rot",
        "w_smooth_neck", "w_smooth_jaw", "w_smooth_eyes", "w_smooth_expr")] + [(n, C.c_int32) for n in (
            "photometric", "jawline_off", "tracking", "training", "opt_pose", "opt_joints", "opt_expr", "opt_shape", "opt_texture", "opt_lights",
            "opt_static_offset", "opt_cam", "bg_mode")] + [("bg_color", C.c_float * 3), ("disturb_rate_fg", C.c_float), ("disturb_rate_bg", C.c_float),
                                                           ("rng_seed", C.c_uint64), ("rng_step", C.c_uint64), ("shared_scale", C.c_float)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        raise RuntimeError(f"{_SO} is missing: build it with `python -m vhap_b200.build_ext` (or __graft_entry__.build()). "
                           "The B200 engine has no CPU or PyTorch fallback.")
    L = C.CDLL(str(_SO))
    vp, i32, f32, i64 = C.c_void_p, C.c_int32, C.c_float, C.c_int64
    P = C.POINTER
    sig = {
        "vhap_abi_version": (i32, []),
        "vhap_last_error": (C.c_char_p, [vp]),
        "vhap_ctx_create": (i32, [P(vp), P(MeshDesc), i32, i32]),
        "vhap_ctx_reserve": (i32, [vp, i32, i32, i32]),
        "vhap_ctx_destroy": (None, [vp]),
        "vhap_set_stage_masks": (i32, [vp] * 7 + [i32, vp]),
        "vhap_flame_forward": (i32, [vp, P(Params), P(FrameBatch), vp, vp, vp, vp]),
        "vhap_flame_backward": (i32, [vp, P(Params), P(FrameBatch), vp, vp, P(Grads), vp]),
        "vhap_project": (i32, [vp, P(Params), P(FrameBatch), vp, vp, vp]),
        "vhap_rasterize": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, i32, vp]),
        "vhap_energy_forward_backward": (i32, [vp, P(Params), P(FrameBatch), P(StageCfg), P(Grads), vp, vp]),
        "vhap_energy_forward": (i32, [vp, P(Params), P(FrameBatch), P(StageCfg), vp, vp]),
        "vhap_energy_backward": (i32, [vp, P(Params), P(FrameBatch), P(StageCfg), vp, vp, i32, P(Grads), vp, vp]),
        "vhap_get_plane": (i32, [vp, i32, vp, vp]),
        "vhap_project_backward": (i32, [vp, P(Params), P(FrameBatch), vp, vp, vp, vp, vp]),
        "vhap_vertex_normals": (i32, [vp, vp, i32, vp, vp]),
        "vhap_vertex_normals_backward": (i32, [vp, vp, vp, i32, vp, vp]),
        "vhap_render_photometric": (i32, [vp, P(Params), P(FrameBatch), P(StageCfg), vp, vp, vp, vp, vp, vp, vp, vp]),
        "vhap_render_rgba_backward": (i32, [vp, P(Params), P(FrameBatch), P(StageCfg), vp, vp, vp, vp, vp, vp]),
        "vhap_set_want_planes": (i32, [vp, i32]),
        "vhap_profile_enable": (i32, [vp, i32]),
        "vhap_profile_kernel_count": (i32, []),
        "vhap_profile_kernel_name": (C.c_char_p, [i32]),
        "vhap_profile_read": (i32, [vp, vp, vp]),
        "vhap_profile_timeline": (i32, [vp, vp, vp, vp, i32]),
        "vhap_set_overlap": (i32, [vp, i32]),
        "vhap_get_geometry": (i32, [vp, i32, vp, vp]),
        "vhap_overflow_flag": (i32, [vp, P(i32)]),
        "vhap_set_injected_random": (i32, [vp, vp, vp]),
        "vhap_set_loss_mask": (i32, [vp, vp]),
        "vhap_set_lr_scale": (i32, [vp, f32, vp]),
        "vhap_tex_grad_ptr": (vp, [vp]),
        "vhap_set_tex_painted": (i32, [vp, vp, vp]),
        "vhap_tex_rebuild": (i32, [vp, vp, vp]),
        "vhap_tex_reg_fold_adam": (i32, [vp, vp, vp, vp, vp, f32, i32, P(StageCfg), f32, vp, vp]),
        "vhap_tex_apply_grad": (i32, [vp, vp, vp, vp, vp, f32, i32, P(StageCfg), vp]),
        "vhap_set_tex_grad_persistent": (i32, [vp, i32]),
        "vhap_dp_init": (i32, [vp, i32, i32, vp]),
        "vhap_dp_connect": (i32, [vp, vp]),
        "vhap_dp_status": (i32, [vp, P(i32)]),
        "vhap_dp_wait_stats": (i32, [vp, P(C.c_uint64), i32]),
        "vhap_dp_tex_connect": (i32, [vp, vp, vp, vp, vp]),
        "vhap_dp_tex_update": (i32, [vp, vp, vp, vp, f32, i32, P(StageCfg), vp]),
        "vhap_dp_tex_part1": (i32, [vp, vp, vp]),
        "vhap_dp_tex_join": (i32, [vp, vp]),
        "vhap_dp_tex_part2": (i32, [vp, vp, vp, vp, f32, i32, P(StageCfg), vp]),
        "vhap_tex_fold_grad_rm": (i32, [vp, vp, vp, vp]),
        "vhap_tex_band_adam": (i32, [vp, vp, vp, i32, i32, vp, vp, f32, i32, P(StageCfg), vp, vp]),
        "vhap_tex_rebuild_rm": (i32, [vp, vp, vp, vp]),
        "vhap_tex_defer": (i32, [vp, i32]),
        "vhap_tex_reg_loss": (i32, [vp, vp, P(StageCfg), vp]),
        "vhap_set_render_wait_event": (i32, [vp, vp]),
        "vhap_assemble_losses": (i32, [vp, P(StageCfg), vp, vp]),
        "vhap_step_counters": (i32, [vp, i32, i32, i32, vp]),
        "vhap_step_advance": (i32, [vp, vp]),
        "vhap_get_cur_mip": (i32, [vp]),
        "vhap_set_cur_mip": (i32, [vp, i32]),
        "vhap_adam_multi": (i32, [vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp]),
        "vhap_adam": (i32, [vp, vp, vp, vp, vp, i64, f32, i32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED = ["vhap_abi_version", "vhap_last_error", "vhap_ctx_create", "vhap_ctx_reserve", "vhap_ctx_destroy", "vhap_set_stage_masks",
            "vhap_flame_forward", "vhap_flame_backward", "vhap_project", "vhap_rasterize", "vhap_energy_forward_backward",
            "vhap_energy_forward", "vhap_energy_backward", "vhap_get_plane", "vhap_get_geometry", "vhap_profile_enable", "vhap_profile_kernel_count", "vhap_profile_kernel_name",
            "vhap_profile_read", "vhap_profile_timeline", "vhap_set_overlap", "vhap_set_want_planes", "vhap_overflow_flag",
            "vhap_set_injected_random", "vhap_set_loss_mask", "vhap_set_lr_scale", "vhap_project_backward", "vhap_vertex_normals",
            "vhap_vertex_normals_backward", "vhap_render_photometric", "vhap_render_rgba_backward", "vhap_tex_grad_ptr", "vhap_set_tex_painted", "vhap_tex_rebuild", "vhap_tex_reg_fold_adam", "vhap_tex_apply_grad", "vhap_set_tex_grad_persistent", "vhap_dp_init", "vhap_dp_connect", "vhap_dp_status", "vhap_dp_wait_stats", "vhap_dp_tex_connect", "vhap_dp_tex_update", "vhap_dp_tex_part1", "vhap_dp_tex_join", "vhap_dp_tex_part2", "vhap_tex_fold_grad_rm", "vhap_tex_band_adam", "vhap_tex_rebuild_rm", "vhap_tex_defer", "vhap_tex_reg_loss", "vhap_set_render_wait_event", "vhap_assemble_losses", "vhap_adam", "vhap_adam_multi", "vhap_step_counters", "vhap_step_advance", "vhap_get_cur_mip", "vhap_set_cur_mip"]
