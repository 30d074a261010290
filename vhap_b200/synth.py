"""Seeded synthetic tracking inputs: parameters, cameras, landmarks, procedural texture and
target images.  Host-side numpy only; used by bench.py, smoke() and the tests."""
from __future__ import annotations

import numpy as np


def procedural_texture(T: int, seed: int = 0) -> np.ndarray:
    """Smooth skin-like base texture [3,T,T] float32 in [0,1] (stands in for tex_mean_painted.png, flame.py:653-659)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, T, dtype=np.float32), np.linspace(0, 1, T, dtype=np.float32), indexing="ij")
    base = np.array([0.72, 0.55, 0.47], np.float32)[:, None, None]
    tex = np.broadcast_to(base, (3, T, T)).copy()
    for _ in range(12):
        fx, fy = rng.uniform(1, 24, 2)
        ph = rng.uniform(0, 2 * np.pi, 2)
        amp = rng.uniform(0.01, 0.05, 3).astype(np.float32)[:, None, None]
        tex += amp * (np.sin(2 * np.pi * fx * xx + ph[0]) * np.cos(2 * np.pi * fy * yy + ph[1]))[None]
    return np.clip(tex, 0.02, 0.98).astype(np.float32)


def procedural_image(B: int, H: int, W: int, seed: int = 0) -> np.ndarray:
    """Smooth background-ish images [B,3,H,W] float32 in [0,1]."""
    rng = np.random.default_rng(seed + 77)
    yy, xx = np.meshgrid(np.linspace(0, 1, H, dtype=np.float32), np.linspace(0, 1, W, dtype=np.float32), indexing="ij")
    img = np.zeros((B, 3, H, W), np.float32)
    for b in range(B):
        for c in range(3):
            acc = np.full((H, W), rng.uniform(0.3, 0.6), np.float32)
            for _ in range(4):
                fx, fy = rng.uniform(0.5, 6, 2)
                ph = rng.uniform(0, 2 * np.pi, 2)
                acc += rng.uniform(0.03, 0.12) * np.sin(2 * np.pi * fx * xx + ph[0]) * np.cos(2 * np.pi * fy * yy + ph[1])
            img[b, c] = acc
    return np.clip(img, 0, 1)


def init_params(model, n_timesteps: int, T: int, seed: int = 0, perturb: float = 1.0) -> dict:
    """Parameter set with the reference's shapes (tracker.py:1279-1341), seeded non-zero values (section 8d)."""
    rng = np.random.default_rng(seed)
    V = model.v_template.shape[0]
    f32 = np.float32
    p = {
        "shape": (rng.normal(0, 0.5, model.n_shape) * perturb).astype(f32),
        "expr": (rng.normal(0, 0.3, (n_timesteps, model.n_expr)) * perturb).astype(f32),
        "rotation": (rng.uniform(-0.2, 0.2, (n_timesteps, 3)) * perturb).astype(f32),
        "neck_pose": (rng.uniform(-0.05, 0.05, (n_timesteps, 3)) * perturb).astype(f32),
        "jaw_pose": np.concatenate([rng.uniform(0, 0.2, (n_timesteps, 1)), rng.uniform(-0.02, 0.02, (n_timesteps, 2))], 1).astype(f32) * f32(perturb),
        "eyes_pose": (rng.uniform(-0.1, 0.1, (n_timesteps, 6)) * perturb).astype(f32),
        "translation": np.zeros((n_timesteps, 3), f32),
        "tex_extra": (rng.normal(0, 0.02, (3, T, T))).astype(f32) * f32(perturb),
        "lights": np.zeros((9, 3), f32),
        "static_offset": (rng.normal(0, 2e-4, (1, V, 3)) * perturb).astype(f32),
        "focal_length": np.array([1.5], f32),
    }
    p["lights"][0] = np.sqrt(4 * np.pi)
    p["lights"][1:] = (rng.normal(0, 0.15, (8, 3)) * perturb).astype(f32)
    # place the head: centre it on the optical axis, bbox height ~0.7 H at focal 1.5 and camera distance 1
    c = model.v_template.mean(0)
    height = model.v_template[:, 1].max() - model.v_template[:, 1].min()
    # image-plane height = f * height / depth (in units of max(h,w)); want 0.7 -> depth = 1.5*height/0.7
    depth = 1.5 * height / 0.7
    p["translation"][:] = np.array([-c[0], -c[1], -c[2] + (1.0 - depth)], f32)
    p["translation"] += (rng.normal(0, 0.004, (n_timesteps, 3)) * perturb).astype(f32)
    return p


def landmarks_px(lmks_ndc_flipped: np.ndarray, H: int, W: int, seed: int = 0, noise_px: float = 1.0) -> np.ndarray:
    """Turn projected landmarks (NDC, y flipped like world_to_ndc(flip_y=True)) into the dataset's `lmk2d`
    [B,68,3] = (x_px, y_px, confidence) (video_dataset.py:226-231; inverse of mesh.py:41-51)."""
    rng = np.random.default_rng(seed + 5)
    x = (lmks_ndc_flipped[:, :68, 0] * 0.5 + 0.5) * W
    y = (lmks_ndc_flipped[:, :68, 1] * 0.5 + 0.5) * H
    # keep every residual at least 0.2 noise_px away from 0: the L1 landmark energy has a kink there and fp32 / fp64
    # would legitimately disagree on the sign of the gradient
    nx, ny = rng.normal(0, noise_px, x.shape), rng.normal(0, noise_px, y.shape)
    nx = np.where(np.abs(nx) < 0.2 * noise_px, np.where(nx < 0, -0.2, 0.2) * noise_px, nx)
    ny = np.where(np.abs(ny) < 0.2 * noise_px, np.where(ny < 0, -0.2, 0.2) * noise_px, ny)
    x = x + nx
    y = y + ny
    conf = rng.uniform(0.6, 1.0, x.shape)
    return np.stack([x, y, conf], -1).astype(np.float32)


def project_ndc(pts: np.ndarray, RT, K, H: int, W: int, focal: float = 1.5) -> np.ndarray:
    """world points [B,N,3] -> NDC (x, y flipped) like NVDiffRenderer.world_to_ndc(flip_y=True) (render_nvdiffrast.py:117-214).
    RT [B,3,4] / K [B,3,3] or [B,4]; None = the uncalibrated default camera (tracker.py:141-157,1333-1337: RT = [I | (0,0,-1)],
    fx = fy = focal * max(H, W), principal point at the image centre)."""
    B = pts.shape[0]
    if RT is None:
        RT = np.tile(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, -1]], np.float64), (B, 1, 1))
    if K is None:
        f = focal * max(H, W)
        K = np.tile(np.array([f, f, 0.5 * W, 0.5 * H], np.float64), (B, 1))
    K = np.asarray(K, np.float64)
    if K.shape[-2:] == (3, 3):
        K = np.stack([K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]], -1)
    RT = np.asarray(RT, np.float64)
    cam = np.einsum("bij,bnj->bni", RT[:, :, :3], pts.astype(np.float64)) + RT[:, None, :, 3]
    fx, fy, cx, cy = (K[:, None, i] for i in range(4))
    clip_x = (2 * fx / W) * cam[..., 0] + ((W - 2 * cx) / W) * cam[..., 2]
    clip_y = (2 * fy / H) * cam[..., 1] + ((H - 2 * cy) / H) * cam[..., 2]
    w = -cam[..., 2]
    return np.stack([clip_x / w, -(clip_y / w)], -1)


def ring_cameras(model, n_views: int, H: int, W: int, head_centre, radius: float = 1.0, yaw_span_deg: float = 120.0, pitch_deg: float = 8.0,
                 fill: float = 1.0, seed: int = 0):
    """Calibrated multi-view rig standing in for NeRSemble's 16 cameras (nersemble_dataset.py:75-127: one shared intrinsic matrix, one
    world-to-camera extrinsic per camera, OpenGL convention): cameras on an arc of `yaw_span_deg` around the head at `radius`,
    alternating +-pitch, all looking at `head_centre`; principal point slightly off-centre.  Returns (RT [n,3,4], K [n,3,3]) float32,
    i.e. sample['extrinsic'] / sample['intrinsic'] of n_views frames of ONE timestep (video_dataset.py:216-217)."""
    rng = np.random.default_rng(seed + 31)
    c = np.asarray(head_centre, np.float64)
    height = float(model.v_template[:, 1].max() - model.v_template[:, 1].min())
    f_px = fill * min(H, W) * radius / height
    RT = np.zeros((n_views, 3, 4), np.float64)
    K = np.zeros((n_views, 3, 3), np.float64)
    yaws = np.deg2rad(np.linspace(-0.5 * yaw_span_deg, 0.5 * yaw_span_deg, n_views)) if n_views > 1 else np.zeros(1)
    for i, yaw in enumerate(yaws):
        pitch = np.deg2rad(pitch_deg) * (1 if i % 2 == 0 else -1)
        z = np.array([np.sin(yaw) * np.cos(pitch), np.sin(pitch), np.cos(yaw) * np.cos(pitch)])     # camera looks down -z (OpenGL)
        o = c + radius * z
        x = np.cross(np.array([0.0, 1.0, 0.0]), z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z], 0)
        RT[i, :, :3] = R
        RT[i, :, 3] = -R @ o
        K[i] = np.array([[f_px, 0, 0.5 * W + rng.uniform(-0.02, 0.02) * W], [0, f_px, 0.5 * H + rng.uniform(-0.02, 0.02) * H], [0, 0, 1]])
    return RT.astype(np.float32), K.astype(np.float32)
