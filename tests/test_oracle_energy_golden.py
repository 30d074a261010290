"""oracle/energy.py against values AND autograd gradients produced by the reference's own tracker methods
(FlameTracker.compute_regularization_energy / compute_lmk_energy and their helpers, vhap/model/tracker.py:347-389,480-690,
called on a bare instance by tests/golden/make_energy_golden.py).  Pins SURVEY.md 8(a) rows a7 and a16 of the oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import energy as OE
from vhap_b200.config import STAGES, EngineConfig
from vhap_b200.flame_model import FlameModelData

G = dict(np.load(Path(__file__).parent / "golden" / "energy_golden.npz"))
PARAMS = ("shape", "expr", "rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose", "lights", "tex_extra", "static_offset")
# names in the oracle's log dict == the reference's log_dict keys
TERMS = ("smooth_pose", "reg_joint", "smooth_joint", "reg_expr", "smooth_expr", "reg_shape", "reg_tex_tv", "reg_tex_res_clusters", "reg_diffuse",
         "reg_offset_lap", "reg_offset", "reg_offset_rigid")


@pytest.fixture(scope="module")
def model():
    m = FlameModelData.synthetic()
    return m, OE.laplacian_dense(m, torch.float32)


@pytest.mark.parametrize("stage_name", ["rgb_global_tracking", "rgb_init_offset", "lmk_init_all"])
def test_regularisers_and_landmark_energy(model, stage_name):
    m, lap = model
    cfg = EngineConfig()
    stage = STAGES[stage_name]
    P = {k: torch.tensor(G["p_" + k]).requires_grad_(True) for k in PARAMS}
    P["focal_length"] = torch.tensor([1.5])
    ts = G["ts"]
    H, W = (int(v) for v in G["image_size"])
    B = len(ts)
    verts_cano = torch.tensor(G["verts_cano"]) + P["static_offset"]
    diffuse = torch.tensor(G["diffuse"]) if stage.photometric else None
    log = OE.regularization_energy(P, ts, stage, cfg, m, verts_cano, diffuse, lap, torch.tensor(G["tex_painted"]), G["uvmask"])
    ref_terms = {t for t in TERMS if f"{stage_name}/{t}" in G}
    assert set(log) == ref_terms, (sorted(log), sorted(ref_terms))                  # the same terms are active in this stage
    for t in ref_terms:
        a, b = float(log[t]), float(G[f"{stage_name}/{t}"])
        assert abs(a - b) <= 2e-5 * max(abs(b), 1e-6), (t, a, b)
    lmks = torch.tensor(G["lmks"]).requires_grad_(True)
    K, RT = OE.fill_cam_params(P, B, H, W)
    dis = stage.disable_jawline_landmarks
    e_lmk = OE.lmk_energy(lmks, torch.tensor(G["lmk2d"]), K, RT, (H, W), cfg.w.always_enable_jawline_landmarks, dis)
    assert abs(float(e_lmk) - float(G[f"{stage_name}/lmk_unweighted"])) <= 2e-5 * float(G[f"{stage_name}/lmk_unweighted"])
    (sum(log.values()) + e_lmk).backward()
    ref = G[f"{stage_name}/g_lmks"]
    assert np.abs(lmks.grad.numpy() - ref).max() <= 2e-5 * np.abs(ref).max()
    for k in PARAMS:
        ref = G[f"{stage_name}/g_{k}"]
        got = P[k].grad.numpy() if P[k].grad is not None else np.zeros_like(ref)
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(got - ref).max() <= 5e-5 * scale, (k, np.abs(got - ref).max(), scale)
