"""On-disk format of the tracking result: `tracked_flame_params[_<epoch>].npz`.

Wire-compatible with the reference's writer `FlameTracker.save_result` (vhap/model/tracker.py:1152-1218) and reader
`load_from_tracked_flame_params` (tracker.py:79-129): same keys, shapes (tracker.py:1279-1341) and dtypes, so that the
files this engine writes are consumed unchanged by `export_as_nerf_dataset.py:151-349` / GaussianAvatars-style loaders and
the reference's own checkpoints can seed this engine.

    key                 shape            note
    rotation            [N_t, 3]         per-timestep axis-angle
    translation         [N_t, 3]
    neck_pose           [N_t, 3]
    jaw_pose            [N_t, 3]
    eyes_pose           [N_t, 6]
    shape               [n_shape]
    expr                [N_t, n_expr]
    timestep_id         [N_t]            the dataset's timestep ids (any dtype numpy stores)
    n_processed_frames  scalar           tracker.timestep
    focal_length        [1]              only for uncalibrated cameras
    tex_extra           [3, T, T]        only with cfg.model.tex_extra
    lights              [9, 3]           only with SH lighting
    static_offset       [1, V, 3]        only with cfg.model.use_static_offset
    image_size          [2]

Pure numpy: usable (and tested) without a GPU."""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Sequence

import numpy as np

PER_FRAME_KEYS = {"rotation": 3, "translation": 3, "neck_pose": 3, "jaw_pose": 3, "eyes_pose": 6}


def engine_params_to_report(params: Dict[str, np.ndarray], timestep_ids: Sequence, n_processed_frames: int, image_size: Sequence[int],
                            calibrated: bool = False, tex_extra: bool = True, use_static_offset: bool = True) -> Dict[str, np.ndarray]:
    """Engine.get_params() (flat fp32 arrays, see Engine.layout) -> the reference's export dict (tracker.py:1158-1213, same key order)."""
    n_t = len(timestep_ids)
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, np.float32))
    out: Dict[str, np.ndarray] = {}
    for k in ("rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose"):
        out[k] = f32(params[k]).reshape(n_t, PER_FRAME_KEYS[k])
    out["shape"] = f32(params["shape"]).reshape(-1)
    out["expr"] = f32(params["expr"]).reshape(n_t, -1)
    out["timestep_id"] = np.asarray(timestep_ids)
    out["n_processed_frames"] = np.asarray(n_processed_frames)
    if not calibrated:
        out["focal_length"] = f32(params["focal_length"]).reshape(1)
    if tex_extra:
        te = f32(params["tex_extra"])
        T = int(round((te.size // 3) ** 0.5))
        out["tex_extra"] = te.reshape(3, T, T)
    out["lights"] = f32(params["lights"]).reshape(9, 3)
    if use_static_offset:
        out["static_offset"] = f32(params["static_offset"]).reshape(1, -1, 3)
    out["image_size"] = np.asarray(image_size)
    return out


def save_tracked_flame_params(out_dir, report: Dict[str, np.ndarray], fname: Optional[str] = None, epoch: Optional[int] = None) -> Path:
    """np.savez under the reference's naming rule (tracker.py:1215-1218)."""
    fname = fname if fname is not None else "tracked_flame_params"
    if epoch is not None:
        fname = f"{fname}_{epoch}"
    path = Path(out_dir) / f"{fname}.npz"
    path.parent.mkdir(parents=True, exist_ok=True)
    np.savez(path, **report)
    return path


def load_tracked_flame_params(fp) -> Dict[str, np.ndarray]:
    with np.load(fp) as z:
        return {k: z[k] for k in z.files}


def report_to_engine_params(report: Dict[str, np.ndarray], current: Dict[str, np.ndarray], n_timesteps: int, calibrated: bool = False,
                            warn=None) -> Dict[str, np.ndarray]:
    """The reference's loader semantics (tracker.py:79-129) on top of the engine's current parameters:
    per-timestep arrays load their first min(N_t, len(file)) rows, missing optional keys keep the current value and warn."""
    warn = warn or (lambda msg: None)
    out = {k: np.array(v, np.float32, copy=True) for k, v in current.items()}
    for k, d in list(PER_FRAME_KEYS.items()) + [("expr", None)]:
        cur = out[k].reshape(n_timesteps, -1)
        src = np.asarray(report[k], np.float32).reshape(len(report[k]), -1)
        n = min(n_timesteps, src.shape[0])
        if src.shape[1] != cur.shape[1]:
            raise ValueError(f"{k}: file has {src.shape[1]} columns, engine expects {cur.shape[1]}")
        cur[:n] = src[:n]
        out[k] = cur.reshape(out[k].shape)
    out["shape"] = np.asarray(report["shape"], np.float32).reshape(out["shape"].shape)
    out["lights"] = np.asarray(report["lights"], np.float32).reshape(out["lights"].shape)
    if not calibrated:
        out["focal_length"] = np.asarray(report["focal_length"], np.float32).reshape(out["focal_length"].shape)
    for k, msg in (("tex_extra", "No tex_extra found in flame_params!"), ("static_offset", "No static_offset found in flame_params!")):
        if k in report:
            if np.asarray(report[k]).size != out[k].size:
                raise ValueError(f"{k}: file has {np.asarray(report[k]).size} values, engine expects {out[k].size}")
            out[k] = np.asarray(report[k], np.float32).reshape(out[k].shape)
        else:
            warn(msg)
    return out
