"""oracle/raster.py: the batched `rasterize_ids` against the per-triangle statement of the specification (`rasterize_ids_loop`):
triangle ids AND winning depths bit-identical -- FLAME scene at several image sizes (incl. odd ones), both cull modes, and a random
triangle soup with huge / degenerate / behind-the-camera / non-finite triangles (every bounding-box class incl. the per-triangle fallback)."""
import numpy as np
import pytest
import torch

from oracle import camera as Cm, energy as E, lbs as L, raster as RA
from tests.scene import make_scene


@pytest.mark.parametrize("size", [(40, 56), (97, 61), (256, 256)])
@pytest.mark.parametrize("cull", [False, True])
def test_flame_scene(size, cull):
    H, W = size
    sc = make_scene(B=2, H=32, W=32, T=8, n_t=3, timesteps=[0, 2])
    m, model = sc["m"], sc["model"]
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in sc["params"].items()}
    ts = torch.as_tensor(sc["ts"]).long()
    with torch.no_grad():
        v, _, _ = L.flame_forward(model, P["shape"][None].expand(2, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts], P["jaw_pose"][ts],
                                  P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
        K, RT = E.fill_cam_params(P, 2, H, W)
        clip = Cm.world_to_clip(v, RT, K, (H, W)).to(torch.float32).numpy()
    a, za = RA.rasterize_ids(clip, m.faces, H, W, cull)
    b, zb = RA.rasterize_ids_loop(clip, m.faces, H, W, cull)
    assert (a > 0).mean() > 0.1
    assert np.array_equal(a, b)
    assert np.array_equal(za.view(np.uint32) & 0x7FFFFFFF, zb.view(np.uint32) & 0x7FFFFFFF) and np.array_equal(za, zb)


@pytest.mark.parametrize("cull", [False, True])
def test_triangle_soup(cull):
    rng = np.random.default_rng(3)
    V = 300
    clip = rng.normal(size=(2, V, 4)).astype(np.float32)
    clip[..., 3] = np.abs(clip[..., 3]) + 0.05
    clip[0, :5, 3] = -1.0                                   # behind the camera
    clip[1, 7, 0] = np.nan
    F = rng.integers(0, V, size=(500, 3))
    F[:10, 1] = F[:10, 0]                                   # degenerate
    small = rng.integers(0, V, size=200)                    # + a cloud of small triangles (the batched classes)
    clip2 = np.concatenate([clip, clip[:, small] + rng.normal(scale=0.02, size=(2, 200, 4)).astype(np.float32) * np.array([1, 1, 1, 0], np.float32),
                            clip[:, small] + rng.normal(scale=0.02, size=(2, 200, 4)).astype(np.float32) * np.array([1, 1, 1, 0], np.float32)], 1)
    F2 = np.concatenate([F, np.stack([small, V + np.arange(200), V + 200 + np.arange(200)], 1)], 0)
    for H, W in ((64, 80), (33, 47)):
        with np.errstate(all="ignore"):
            a, za = RA.rasterize_ids(clip2, F2, H, W, cull)
            b, zb = RA.rasterize_ids_loop(clip2, F2, H, W, cull)
        assert np.array_equal(a, b) and np.array_equal(za, zb)
        assert len(np.unique(a)) > 50
