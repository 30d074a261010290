"""oracle/lbs.py:flame_forward (+ the per-timestep gathers / albedo composition of the energy oracle) against the reference's own
FlameHead.forward (flame.py:571-646) and FlameTracker.forward_flame (tracker.py:213-235) run on bare instances carrying this repo's
synthetic FLAME buffers (tests/golden/make_flame_golden.py): values and gradients.  Pins SURVEY.md 8(a) rows a1 and a4."""
from pathlib import Path

import numpy as np
import torch

from oracle import lbs as OL
from tests.scene import get_model

G = dict(np.load(Path(__file__).parent / "golden" / "flame_golden.npz"))
PARAMS = ("shape", "expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation", "static_offset", "tex_extra")


def close(a, b, tol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.abs(a - b).max() <= tol * max(1e-12, np.abs(b).max()), (what, np.abs(a - b).max(), np.abs(b).max())


def _forward(P, ts, model, **kw):
    B = len(ts)
    return OL.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts], P["jaw_pose"][ts],
                            P["eyes_pose"][ts], P["translation"][ts], **kw)


def test_forward_flame_values_and_gradients():
    model = OL.model_tensors(get_model(), torch.float32)
    P = {k: torch.tensor(G["p_" + k]).requires_grad_(True) for k in PARAMS}
    ts = torch.as_tensor(G["ts"]).long()
    verts, verts_cano, lmks = _forward(P, ts, model, static_offset=P["static_offset"])
    albedos = (torch.tensor(G["tex_painted"]) + P["tex_extra"][None]).expand(len(ts), -1, -1, -1)          # get_albedo + expand (tracker.py:234,247-258)
    close(verts.detach().numpy(), G["verts"], 2e-6, "verts"); close(verts_cano.detach().numpy(), G["verts_cano"], 2e-6, "verts_cano")
    close(lmks.detach().numpy(), G["lmks"], 2e-6, "lmks"); close(albedos.detach().numpy(), G["albedos"], 1e-7, "albedos")
    T = lambda k: torch.tensor(G[k])
    ((verts * T("w_verts")).sum() + (verts_cano * T("w_cano")).sum() + (lmks * T("w_lmks")).sum() + (albedos * T("w_alb")).sum()).backward()
    for k in PARAMS:
        close(P[k].grad.numpy(), G["g_" + k], 5e-5, "g_" + k)


def test_zero_centered_at_root_node():
    model = OL.model_tensors(get_model(), torch.float32)
    P = {k: torch.tensor(G["p_" + k]) for k in PARAMS}
    ts = torch.as_tensor(G["ts"]).long()
    verts, _, lmks = _forward(P, ts, model, zero_centered_at_root_node=True)
    close(verts.numpy(), G["verts_zero_centered"], 2e-6, "verts"); close(lmks.numpy(), G["lmks_zero_centered"], 2e-6, "lmks")
