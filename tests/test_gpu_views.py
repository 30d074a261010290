"""View sharing (SURVEY a1 "NeRSemble: FLAME evaluated 16x for identical params"): with several calibrated cameras per timestep the engine
evaluates FLAME once per distinct timestep and only projects per view (vhap_frame_batch::geo).  The result must equal the per-frame
evaluation the reference does (tracker.py:213-235) -- same losses, same gradients."""
import numpy as np
import pytest
import torch

from tests.scene import make_scene

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("timesteps", [[1, 1, 1, 1, 1, 1], [2, 1, 2, 1, 1, 2]])
def test_shared_geometry_equals_per_frame(timesteps):
    from vhap_b200.config import nersemble_config, NERSEMBLE_STAGES
    from vhap_b200.engine import Engine
    sc = make_scene(B=6, H=160, W=112, T=128, n_t=3, timesteps=timesteps, views=True)
    cfg = nersemble_config(tex_resolution=128)
    e = Engine(sc["m"], cfg, 3, tex_painted=sc["tex_painted"])
    try:
        res = {}
        for share in (False, True):
            e.load_params(sc["params"])
            e.set_stage(NERSEMBLE_STAGES["rgb_global_tracking"])
            e.inject_random(sc["w_fg"], sc["w_bg"], sc["u_rand"])
            batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"], RT=sc["RT"], K=sc["K"], share_views=share)
            assert (batch.geo is not None) == share
            if share:
                assert batch.c.n_geo == len(set(timesteps))
            e.zero_grad()
            e.energy(batch, backward=True, training=True)
            tex = e.texture_grad_dense().clone().cpu().numpy()
            torch.cuda.synchronize()
            res[share] = dict(loss=e.loss_dict(), g={k: v.clone().cpu().numpy() for k, v in e.g.items()}, tex=tex)
        for k, v in res[False]["loss"].items():
            assert abs(res[True]["loss"][k] - v) <= 1e-5 * max(abs(v), 1e-6), (k, res[True]["loss"][k], v)
        errs = {k: rel(res[True]["g"][k], res[False]["g"][k]) for k in res[False]["g"] if np.abs(res[False]["g"][k]).max() > 0}
        errs["tex"] = rel(res[True]["tex"], res[False]["tex"])
        print("view sharing vs per-frame", {k: float("%.3g" % v) for k, v in errs.items()})
        assert all(v < 1e-4 for v in errs.values()), errs          # same arithmetic per view; only the order of the float atomics differs
        # default: on for calibrated batches with repeated timesteps
        assert e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"], RT=sc["RT"], K=sc["K"]).geo is not None
    finally:
        e.close()
