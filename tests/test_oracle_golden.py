"""CPU tests: the oracle's FLAME / LBS / camera restatement against golden vectors produced by the REFERENCE's own
vhap/model/lbs.py, vhap/util/mesh.py, vhap/util/vector_ops.py (tests/golden/make_golden.py, lbs_golden.npz)."""
from pathlib import Path

import numpy as np
import torch

from oracle import camera as C, energy as E, lbs as L, render as RE

G = np.load(Path(__file__).parent / "golden" / "lbs_golden.npz")
t = lambda k, dt=torch.float64: torch.as_tensor(G[k]).to(dt)


def _model(dt):
    return dict(v_template=t("v_template", dt), shapedirs=t("shapedirs", dt), posedirs=t("posedirs", dt), J_regressor=t("J_regressor", dt),
                parents=torch.tensor([-1, 0, 1, 1, 1]), lbs_weights=t("lbs_weights", dt), faces=torch.as_tensor(G["faces"]),
                lmk_faces_idx=torch.as_tensor(G["lmk_faces_idx"])[0], lmk_bary=t("lmk_bary", dt)[0])


def _run(dt, grad=False):
    b, p, tr, off = t("betas", dt), t("pose", dt), t("transl", dt), t("offset", dt)
    if grad:
        for x in (b, p, tr, off):
            x.requires_grad_(True)
    v, vs, lm = L.flame_forward(_model(dt), b[:, :5], b[:, 5:], p[:, 0:3], p[:, 3:6], p[:, 6:9], p[:, 9:15], tr, static_offset=off)
    return v, vs, lm, (b, p, tr, off)


def test_flame_forward_f64_matches_reference_exactly():
    v, vs, lm, _ = _run(torch.float64)
    assert (v - t("f64_verts")).abs().max() < 1e-14
    assert (vs - t("f64_v_shaped")).abs().max() < 1e-14
    assert (lm - t("f64_lmks")).abs().max() < 1e-14


def test_flame_forward_f32_matches_reference():
    v, vs, lm, _ = _run(torch.float32)
    assert (v - t("f32_verts", torch.float32)).abs().max() < 2e-6
    assert (lm - t("f32_lmks", torch.float32)).abs().max() < 2e-6


def test_rodrigues_and_joints_match_reference():
    p = t("pose")
    rot = L.batch_rodrigues(p.view(-1, 3))
    assert (rot - t("f64_rot")).abs().max() < 1e-14
    v, vs, lm, _ = _run(torch.float64)
    m = _model(torch.float64)
    _, J, A1 = L.lbs(p, vs, m["posedirs"], m["J_regressor"], m["parents"], m["lbs_weights"])
    assert (J - t("f64_joints")).abs().max() < 1e-14
    assert (A1 - t("f64_A1")).abs().max() < 1e-14


def test_gradients_match_reference_autograd():
    v, vs, lm, (b, p, tr, off) = _run(torch.float64, grad=True)
    ((v * t("g_wv")).sum() + (lm * t("g_wl")).sum()).backward()
    assert (b.grad - t("g_betas")).abs().max() < 1e-12
    assert (p.grad - t("g_pose")).abs().max() < 1e-12
    assert (tr.grad - t("g_transl")).abs().max() < 1e-12
    assert (off.grad - t("g_offset")).abs().max() < 1e-12


def test_small_helpers_match_reference():
    un, vn = C.normalize_image_points(t("nip_u"), t("nip_v"), (240, 320))
    assert (un - t("nip_un")).abs().max() < 1e-14 and (vn - t("nip_vn")).abs().max() < 1e-14
    assert (RE.safe_normalize(t("sn_x")) - t("sn_y")).abs().max() < 1e-14


def test_joint_prior_matches_reference():
    from vhap_b200.config import LossWeightConfig
    Eo = E.joint_L2_energy(t("jp_neck"), t("jp_jaw"), t("jp_eyes"), LossWeightConfig())
    assert abs(float(Eo) - float(G["jp_E"])) < 1e-14
