"""oracle/energy.py:compute_energy END TO END against the reference's own FlameTracker.compute_energy (tracker.py:692-750) run on CPU by
tests/golden/make_e2e_golden.py: the reference's tracker, FlameHead, lbs and NVDiffRenderer code, unmodified, with only the four
nvdiffrast entry points served by the oracle's restatements of those ops and the disturbance draws injected.  Every log term, the
total and the gradients w.r.t. all parameters (incl. the focal length), four stages."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import energy as OE
from oracle import lbs as OL
from tests.scene import get_model
from vhap_b200.config import STAGES, EngineConfig

G = dict(np.load(Path(__file__).parent / "golden" / "e2e_golden.npz"))
PARAMS = ("shape", "expr", "rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose", "lights", "tex_extra", "static_offset", "focal_length")


@pytest.fixture(scope="module")
def setup():
    m = get_model()
    return m, OL.model_tensors(m, torch.float32), OE.laplacian_dense(m, torch.float32)


@pytest.mark.parametrize("stage_name", ["rgb_global_tracking", "rgb_init_all", "lmk_init_all", None])
def test_compute_energy_matches_reference_code(setup, stage_name):
    m, model, lap = setup
    key = str(stage_name)
    T = G["p_tex_extra"].shape[-1]
    cfg = EngineConfig(tex_resolution=T)
    P = {k: torch.tensor(G["p_" + k]).requires_grad_(True) for k in PARAMS}
    sample = dict(rgb=torch.tensor(G["rgb"]), lmk2d=torch.tensor(G["lmk2d"]), timestep_index=G["ts"], uvmask_res=G["uvmask"])
    dist = dict(w_fg=torch.tensor(G["w_fg"]), w_bg=torch.tensor(G["w_bg"]), u_rand=torch.tensor(G["u_rand"]))
    stage = STAGES[stage_name] if stage_name else None
    E, log = OE.compute_energy(P, sample, stage, cfg, m, model, lap=lap, disturbance=dist, tex_painted=torch.tensor(G["tex_painted"]))
    ref_terms = {k.split("/", 1)[1] for k in G if k.startswith(key + "/") and not k.split("/", 1)[1].startswith("g_")}
    assert set(log) == ref_terms, (sorted(log), sorted(ref_terms))
    for t in ref_terms:
        a, b = float(log[t]), float(G[f"{key}/{t}"])
        assert abs(a - b) <= 3e-5 * max(abs(b), 1e-6), (t, a, b)
    E.backward()
    for k in PARAMS:
        ref = G[f"{key}/g_{k}"]
        got = P[k].grad.numpy() if P[k].grad is not None else np.zeros_like(ref)
        scale = np.abs(ref).max()
        if scale == 0:
            assert np.abs(got).max() == 0, k
        else:
            assert np.abs(got - ref).max() <= 3e-4 * scale, (k, np.abs(got - ref).max(), scale)
