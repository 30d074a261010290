"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA engine, called through the C-ABI, against the oracle
on the same seeded inputs.  Tolerances: triangle ids bit-exact; everything floating point 1e-4 relative to the
quantity's own scale on identical inputs (north_star), looser where fp32 FLAME output feeds the rasteriser (stated)."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.scene import make_scene, get_model

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def eng_small():
    from vhap_b200.engine import Engine
    sc = make_scene(B=3, H=96, W=96, T=256, n_t=4, timesteps=[1, 2, 1])
    e = Engine(sc["m"], sc["cfg"], 4, tex_painted=sc["tex_painted"])
    e.load_params(sc["params"])
    yield e, sc
    e.close()


def _oracle_params(sc, dt=torch.float64, grad=True):
    return {k: torch.tensor(v, dtype=dt, requires_grad=grad) for k, v in sc["params"].items()}


def test_library_loaded_is_cuda():
    from vhap_b200 import _lib
    L = _lib.lib()
    assert L.vhap_abi_version() == 2


def test_flame_forward_backward(eng_small):
    from oracle import lbs as L
    from vhap_b200 import _lib
    e, sc = eng_small
    m, model, ts = sc["m"], sc["model"], sc["ts"]
    B, V = len(ts), e.V
    P = _oracle_params(sc)
    verts, cano, lm = L.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts],
                                      P["jaw_pose"][ts], P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
    g = torch.Generator().manual_seed(3)
    wv = torch.randn(B, V, 3, generator=g, dtype=torch.float64)
    wl = torch.randn(B, 70, 3, generator=g, dtype=torch.float64)
    ((verts * wv).sum() + (lm * wl).sum()).backward()
    rgb = torch.zeros(B, 3, 32, 32)
    batch = e.stage_sample(rgb, np.zeros((B, 68, 3), np.float32), ts)
    dv = torch.empty(B, V, 3, device=e.dev); dc = torch.empty(B, V, 3, device=e.dev); dl = torch.empty(B, 70, 3, device=e.dev)
    cp = e._c_params()
    e._ck(e.L.vhap_flame_forward(e.ctx, C.byref(cp), C.byref(batch.c), dv.data_ptr(), dc.data_ptr(), dl.data_ptr(), e._stream()), None)
    assert rel(dv.cpu().numpy(), verts.detach().numpy()) < 1e-5
    assert rel(dc.cpu().numpy(), cano.detach().numpy()) < 1e-5
    assert rel(dl.cpu().numpy(), lm.detach().numpy()) < 1e-5
    e.zero_grad()
    opt = {k: True for k in ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset")}
    opt["texture"] = False
    cg = e._c_grads(opt)
    gv = wv.to(torch.float32).to(e.dev).contiguous(); gl = wl.to(torch.float32).to(e.dev).contiguous()
    e._ck(e.L.vhap_flame_backward(e.ctx, C.byref(cp), C.byref(batch.c), gv.data_ptr(), gl.data_ptr(), C.byref(cg), e._stream()), None)
    torch.cuda.synchronize()
    for k in ("shape", "expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation", "static_offset"):
        ref = P[k].grad.numpy().reshape(-1)
        got = e.g[k].cpu().numpy()
        # pose gradients pass through A.t = G.t - G.R J with |J| ~ 1.5 (template not centred): fp32 cancellation, 3e-4
        assert rel(got, ref) < (3e-4 if "pose" in k or k == "rotation" else 1e-4), (k, rel(got, ref))


@pytest.mark.parametrize("H,W", [(64, 64), (136, 200), (256, 256)])
def test_rasterize_ids_bit_exact(eng_small, H, W):
    from oracle import lbs as L, energy as E, camera as Cm, raster as RA
    e, sc = eng_small
    m, model = sc["m"], sc["model"]
    P = _oracle_params(sc, grad=False)
    ts = np.array([0, 3])
    B = 2
    verts, _, _ = L.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts], P["jaw_pose"][ts],
                                  P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
    K, RT = E.fill_cam_params(P, B, H, W)
    clip = Cm.world_to_clip(verts, RT, K, (H, W)).to(torch.float32)
    ids_ref, _ = RA.rasterize_ids(clip.numpy(), m.faces, H, W)
    e.reserve(B, H, W)
    dclip = clip.to(e.dev).contiguous()
    ids = torch.empty(B, H, W, dtype=torch.int32, device=e.dev)
    rast = torch.empty(B, H, W, 4, device=e.dev); db = torch.empty(B, H, W, 4, device=e.dev)
    e._ck(e.L.vhap_rasterize(e.ctx, dclip.data_ptr(), B, H, W, ids.data_ptr(), rast.data_ptr(), db.data_ptr(), 0, e._stream()), None)
    got = ids.cpu().numpy()
    assert (ids_ref > 0).mean() > 0.05
    assert np.array_equal(got, ids_ref), f"{(got != ids_ref).sum()} of {got.size} ids differ"
    r_ref, db_ref = RA.shade_pass(clip.to(torch.float64), model["faces"], ids_ref)
    # barycentrics of sliver triangles (near-zero screen area) are ill-conditioned in fp32, exactly as in nvdiffrast's
    # fp32 shader; require 1e-4 on all but a handful of such pixels
    fg = ids_ref > 0
    du = np.abs(rast.cpu().numpy()[..., :3] - r_ref.numpy()[..., :3]).max(-1)[fg]
    assert np.median(du) < 2e-6 and (du > 1e-4).mean() < 2e-3, (np.median(du), (du > 1e-4).mean())
    dref = db_ref.numpy()
    dd = (np.abs(db.cpu().numpy() - dref).max(-1) / np.maximum(np.abs(dref).max(-1), 1e-3))[fg]
    assert np.median(dd) < 1e-5 and (dd > 1e-3).mean() < 5e-3, (np.median(dd), (dd > 1e-3).mean())


def test_rasterize_random_soup_bit_exact(eng_small):
    """stress: random clip-space vertices incl. behind-camera, degenerate and huge triangles, back-face culling"""
    from oracle import raster as RA
    e, sc = eng_small
    m = sc["m"]
    rng = np.random.default_rng(5)
    B, H, W, V = 2, 72, 104, e.V
    clip = np.zeros((B, V, 4), np.float32)
    clip[..., :2] = rng.normal(0, 0.8, (B, V, 2))
    clip[..., 2] = rng.uniform(-1.2, 1.2, (B, V))
    clip[..., 3] = rng.uniform(0.5, 2.0, (B, V))
    clip[0, ::97, 3] = -0.3                      # behind the camera
    clip[1, ::131] = clip[1, 1::131][: clip[1, ::131].shape[0]]   # duplicated vertices -> degenerate triangles
    clip[0, 5, :2] = 400.0                       # far outside the guard band
    # make it sparse enough to be a meaningful test: shrink triangles by pulling vertices of a face together
    f = m.faces
    cen = clip[:, f[:, 0]]
    clip[:, f[:, 1], :3] = 0.9 * cen[..., :3] + 0.1 * clip[:, f[:, 1], :3]
    clip[:, f[:, 2], :3] = 0.9 * cen[..., :3] + 0.1 * clip[:, f[:, 2], :3]
    for cull in (0, 1):
        ids_ref, _ = RA.rasterize_ids(clip, f, H, W, cull_backface=bool(cull))
        e.reserve(B, H, W)
        d = torch.tensor(clip, device=e.dev)
        ids = torch.empty(B, H, W, dtype=torch.int32, device=e.dev)
        e._ck(e.L.vhap_rasterize(e.ctx, d.data_ptr(), B, H, W, ids.data_ptr(), None, None, cull, e._stream()), None)
        got = ids.cpu().numpy()
        assert np.array_equal(got, ids_ref), f"cull={cull}: {(got != ids_ref).sum()} ids differ"


def _oracle_energy(sc, stage, disturbance, P):
    from oracle import energy as E
    sample = dict(rgb=sc["rgb16"].to(torch.float64), lmk2d=torch.tensor(sc["lmk2d"]), timestep_index=sc["ts"])
    tp = torch.tensor(sc["tex_painted"], dtype=torch.float64)
    return E.compute_energy(P, sample, stage, sc["cfg"], sc["m"], sc["model"], disturbance=disturbance, tex_painted=tp, return_aux=True)


@pytest.mark.parametrize("stage_name", ["rgb_global_tracking", "rgb_init_all", "lmk_init_all", None])
def test_energy_and_gradients(eng_small, stage_name):
    _check_energy_and_gradients(eng_small, stage_name)


@pytest.fixture(scope="module")
def eng_pow2():
    from vhap_b200.engine import Engine
    sc = make_scene(B=2, H=64, W=128, T=128, n_t=3, timesteps=[0, 2])
    e = Engine(sc["m"], sc["cfg"], 3, tex_painted=sc["tex_painted"])
    e.load_params(sc["params"])
    yield e, sc
    e.close()


def test_energy_and_gradients_power_of_two_image(eng_pow2):
    """power-of-two image sizes take the shift/mask pixel-index path of the per-pixel passes (the bench sizes 512 / 1024)"""
    _check_energy_and_gradients(eng_pow2, "rgb_global_tracking")


def _check_energy_and_gradients(eng, stage_name):
    from vhap_b200.config import STAGES
    from vhap_b200 import _lib
    e, sc = eng
    stage = STAGES[stage_name] if stage_name else None
    P = _oracle_params(sc)
    dist = dict(w_fg=sc["w_fg"], w_bg=sc["w_bg"], u_rand=sc["u_rand"])
    Et, log, aux = _oracle_energy(sc, stage, dist, P)
    e.load_params(sc["params"])
    e.set_stage(stage)
    e.inject_random(sc["w_fg"], sc["w_bg"], sc["u_rand"])
    batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
    e.zero_grad()
    e.energy(batch, backward=stage is not None, training=True)
    tex_g = e.texture_grad_dense().cpu().numpy() if (stage is not None and "texture" in stage.optimizable_params) else None
    torch.cuda.synchronize()
    flag = C.c_int32(0)
    e.L.vhap_overflow_flag(e.ctx, C.byref(flag))
    assert flag.value == 0
    got = e.loss_dict()
    name_map = {"reg_tex_res_clusters": "reg_tex_res_clusters"}
    if "photo" in log:
        ids_ref = aux["rast"][..., 3].detach().numpy().astype(np.int32)
        planes = e.render_planes(batch, training=stage is not None)
        ids_got = planes["cid"][..., 1].cpu().numpy().astype(np.int32)[:, ::-1]       # plane is flipped to image orientation
        mism = int((ids_got != ids_ref).sum())
        assert mism <= 4, f"{mism} rasterised ids differ from the oracle (fp32 vs fp64 vertex positions)"
        assert abs(got["n_fg"] - float(aux["n_fg"]) / 3) <= 2
        rg = planes["rgba"].cpu().numpy()
        rr = aux["render"]["rgba"].detach().numpy()
        bad = np.abs(rg - rr).max(-1) > 2e-4
        # differences are confined to silhouette / sliver pixels whose fp32 barycentrics are ill-conditioned w.r.t. the
        # last-bit differences between this engine's fp32 FLAME output and the float64 oracle (and to pixels that
        # sampled such a pixel from a cluster pool); tests/test_gpu_modular.py checks 1e-4 on identical inputs
        assert bad.sum() <= 0.005 * bad.size, f"{bad.sum()} pixels differ"
    for k, v in log.items():
        if k == "total":
            continue
        assert abs(got[k] - float(v)) <= 2e-4 * max(abs(float(v)), 1e-3) + (1e-3 if k == "photo" else 0), (k, got[k], float(v))
    if stage is None:
        return
    Et.backward()
    # end to end the geometry gradients of the photometric term inherit the conditioning of sliver pixels (see above):
    # compare in relative L2 there; the landmark-only stage is smooth and is held to 3e-4 in max norm
    tol = 0.2 if stage.photometric else 3e-4
    metric = (lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64).ravel() - np.asarray(b, np.float64).ravel())
                                 / max(np.linalg.norm(np.asarray(b, np.float64).ravel()), 1e-30))) if stage.photometric else rel
    from vhap_b200.config import opt_dict_for
    opt = opt_dict_for(stage)
    groups = {"shape": "shape", "expr": "expr", "pose": ("rotation", "translation"), "joints": ("neck_pose", "jaw_pose", "eyes_pose"),
              "lights": "lights", "static_offset": "static_offset", "cam": "focal_length"}
    errs = {}
    for flag_name, names in groups.items():
        if not opt[flag_name]:
            continue
        for nme in ([names] if isinstance(names, str) else names):
            if P[nme].grad is None:
                continue
            ref = P[nme].grad.numpy().reshape(-1)
            g = e.g[nme].cpu().numpy()
            errs[nme] = metric(g, ref)
    if tex_g is not None:
        errs["tex_extra"] = metric(tex_g, P["tex_extra"].grad.numpy())
    print("gradient rel errors", stage_name, {k: float("%.3g" % v) for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, bad


def test_adam_matches_torch(eng_small):
    e, sc = eng_small
    n = 1000
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=g); gr = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    optim = torch.optim.Adam([ref], lr=5e-3)
    p = p0.clone().to(e.dev); m = torch.zeros(n, device=e.dev); v = torch.zeros(n, device=e.dev)
    for step in range(1, 4):
        ref.grad = gr * step
        optim.step()
        gd = (gr * step).to(e.dev)
        e._ck(e.L.vhap_adam(e.ctx, p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 5e-3, step, e._stream()), None)
    assert rel(p.cpu().numpy(), ref.detach().numpy()) < 1e-5


def _traj_tol(k, base):
    """tolerance for comparing two optimisation trajectories of the same step sequence: the per-vertex offsets start at zero, so their
    first Adam steps are +-lr * sign(g) and vertices whose tiny gradients sit at the atomics' noise level flip sign between runs
    (tools/debug_graph.py); a broken schedule changes every group by O(1e-1) and is still caught through the other parameters."""
    return 0.1 if k == "static_offset" else base


def test_full_step_decreases_energy(eng_small):
    e, sc = eng_small
    e.load_params(sc["params"])
    e.set_stage("rgb_global_tracking")
    e.inject_random(None, None, None)
    batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
    first = None
    for i in range(12):
        l = e.step(batch)
        tot = float(l[0].item())
        assert np.isfinite(tot)
        if first is None:
            first = tot
    assert tot < first


@pytest.mark.parametrize("pipelined", [True, False])
def test_graph_replay_matches_eager(eng_small, pipelined):
    """whole-step CUDA-graph replay (device-resident step counters) == eager stepping, up to the run-to-run noise of the
    floating-point atomics (measured by repeating the eager run).  Two steps only: both texture ping-pong parities and two
    Adam bias corrections are exercised, while the trajectory has not yet reached the first discrete bifurcation (from the
    third step on, last-bit differences of the atomics flip a coverage / arg-max decision and eager runs themselves split
    into two deterministic branches, tools/debug_graph.py)."""
    e, sc = eng_small
    batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
    res = []
    for mode in ("eager", "eager", "graph"):
        e.load_params(sc["params"])
        e.set_stage("rgb_global_tracking")
        e.inject_random(None, None, None)
        e.global_step = 5
        if mode == "graph":
            e.graph_begin([batch], pipelined=pipelined)
        step_losses = []
        for i in range(2):                          # pipelined: eager prologue + 1 replay + flush in graph_end
            e.graph_step(0) if mode == "graph" else e.step(batch)
            step_losses.append(e.loss_dict())
        if mode == "graph":
            e.graph_end()
        torch.cuda.synchronize()
        res.append({k: v.copy() for k, v in e.get_params().items()})
        res[-1]["__losses__"] = step_losses
    losses = [r.pop("__losses__") for r in res]
    for i in range(2):                              # every step's loss vector is complete (incl. the texture regularisers) in graph mode too
        for name in ("reg_tex_tv", "photo", "total"):
            a, b = losses[2][i][name], losses[0][i][name]
            assert abs(a - b) <= 2e-3 * max(abs(b), 1e-6), (i, name, a, b)
    for k in res[0]:
        noise = rel(res[1][k], res[0][k])
        # a broken replay (stuck Adam / RNG step counter, wrong texture ping-pong parity) changes the trajectory by O(1e-1)
        assert rel(res[2][k], res[0][k]) < _traj_tol(k, 10 * noise + 1e-3), (k, rel(res[2][k], res[0][k]), noise)


def test_data_parallel_texture_path_matches_fused(eng_small):
    """world_size 1 with an identity 'all-reduce': the data-parallel update (dense gradient -> reduce -> vhap_tex_apply_grad, gradient slab
    reduce -> Adam) must reproduce the fused single-GPU step (same texture, same mip pyramid seen by the next forward)"""
    from vhap_b200.parallel import DataParallelStep
    e, sc = eng_small
    batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
    res = []
    for mode in ("fused", "dp"):
        e.load_params(sc["params"])
        e.set_stage("rgb_global_tracking")
        e.inject_random(None, None, None)
        e.global_step = 3
        for i in range(1):      # one step: later steps amplify the run-to-run noise of the atomics (see test_graph_replay_matches_eager)
            if mode == "fused":
                e.step(batch)
            else:
                e.zero_grad()
                e.energy(batch, backward=True, training=True, global_B=batch.B, reduce_fn=lambda a, b: b.copy_(a))
                e.adam_step(allreduce_fn=lambda t: None)
                e.global_step += 1
        losses = e.energy(batch, backward=False, training=True).clone()      # forward through the rebuilt pyramid
        torch.cuda.synchronize()
        res.append(({k: v.copy() for k, v in e.get_params().items()}, losses.cpu().numpy()))
    for k in res[0][0]:
        assert rel(res[1][0][k], res[0][0][k]) < _traj_tol(k, 2e-3), (k, rel(res[1][0][k], res[0][0][k]))
    assert abs(res[1][1][0] - res[0][1][0]) < 1e-3 * abs(res[0][1][0])


def test_pipelined_replay_flushes_on_read(eng_small):
    """reading the parameters in the middle of a pipelined replay applies the pending texture update and restarts the pipeline"""
    e, sc = eng_small
    batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
    res = []
    for mode in ("eager", "graph"):
        e.load_params(sc["params"])
        e.set_stage("rgb_global_tracking")
        e.inject_random(None, None, None)
        e.global_step = 7
        mid = None
        if mode == "graph":
            e.graph_begin([batch], pipelined=True)
        for i in range(2):
            e.graph_step(0) if mode == "graph" else e.step(batch)
            if i == 0:
                mid = e.get_params()["tex_extra"].copy()          # graph mode: flush + new prologue on the next step
        if mode == "graph":
            e.graph_end()
        torch.cuda.synchronize()
        res.append((mid, {k: v.copy() for k, v in e.get_params().items()}))
    assert rel(res[1][0], res[0][0]) < 1e-3
    for k in res[0][1]:
        assert rel(res[1][1][k], res[0][1][k]) < _traj_tol(k, 3e-3), (k, rel(res[1][1][k], res[0][1][k]))
