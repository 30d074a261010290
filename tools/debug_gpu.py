import sys, copy
sys.path.insert(0, '.')
import numpy as np, torch
from tests.scene import make_scene
from vhap_b200.engine import Engine
from vhap_b200.config import STAGES, opt_dict_for
from oracle import energy as E

def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)

sc = make_scene(B=3, H=96, W=96, T=256, n_t=4, timesteps=[1, 2, 1])
e = Engine(sc["m"], sc["cfg"], 4, tex_painted=sc["tex_painted"])
dist = dict(w_fg=sc["w_fg"], w_bg=sc["w_bg"], u_rand=sc["u_rand"])
sample = dict(rgb=sc["rgb16"].to(torch.float64), lmk2d=torch.tensor(sc["lmk2d"]), timestep_index=sc["ts"])
tp = torch.tensor(sc["tex_painted"], dtype=torch.float64)

def run(stage_name, mod):
    cfg = copy.deepcopy(sc["cfg"]); mod(cfg.w)
    e.cfg = cfg
    stage = STAGES[stage_name]
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc["params"].items()}
    Et, log, aux = E.compute_energy(P, sample, stage, cfg, sc["m"], sc["model"], disturbance=dist, tex_painted=tp, return_aux=True)
    Et.backward()
    e.load_params(sc["params"]); e.set_stage(stage); e.inject_random(sc["w_fg"], sc["w_bg"], sc["u_rand"])
    batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], sc["ts"])
    e.zero_grad(); e.energy(batch, True, True)
    tg = e.texture_grad_dense().cpu().numpy() if "texture" in stage.optimizable_params else None
    torch.cuda.synchronize()
    got = e.loss_dict()
    print("==", stage_name, {k: (round(got[k], 6), round(float(v), 6)) for k, v in log.items() if k != "total"})
    errs = {}
    for k in ("shape", "expr", "rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose", "lights", "static_offset", "focal_length"):
        if P[k].grad is not None and float(P[k].grad.abs().max()) > 0:
            errs[k] = float("%.3g" % rel(e.g[k].cpu().numpy(), P[k].grad.numpy().reshape(-1)))
    if tg is not None and P["tex_extra"].grad is not None:
        errs["tex"] = float("%.3g" % rel(tg, P["tex_extra"].grad.numpy()))
    print("   grad rel err:", errs)
    return aux, batch, stage

def only_lmk(w):
    w.reg_shape = w.reg_expr = w.reg_neck = w.reg_jaw = w.reg_eyes = 0.0
def only_regs(w):
    w.landmark = None
def nothing(w):
    pass
def no_photo_regs(w):
    w.reg_diffuse = None; w.reg_tex_tv = None; w.reg_tex_res_clusters = None; w.landmark = None
    w.reg_shape = w.reg_expr = w.reg_neck = w.reg_jaw = w.reg_eyes = 0.0
    w.reg_offset = w.reg_offset_lap = w.reg_offset_rigid = None
    w.smooth_trans = w.smooth_rot = w.smooth_neck = w.smooth_jaw = w.smooth_expr = 0.0
run("lmk_init_all", only_lmk)
run("lmk_init_all", only_regs)
aux, batch, stage = run("rgb_global_tracking", no_photo_regs)
run("rgb_global_tracking", nothing)
# pixel analysis
e.cfg = sc["cfg"]
aux, batch, stage = run("rgb_global_tracking", nothing)
planes = e.render_planes(batch, training=True)
ids_ref = aux["rast"][..., 3].detach().numpy().astype(np.int32)
ids_got = planes["cid"][..., 1].cpu().numpy().astype(np.int32)[:, ::-1]
print("id mismatches", (ids_got != ids_ref).sum())
for name in ("rgba", "albedo", "normal", "diffuse"):
    g = planes[name].cpu().numpy()[..., :3]
    r = aux["render"][name].detach().numpy()[..., :3]
    fg = (ids_ref[:, ::-1] > 0)
    err = np.abs(g - r).max(-1)
    if name != "rgba":
        err = err * fg
    bad = err > 2e-4
    # boundary pixels: any 4-neighbour with a different id (image orientation)
    idi = ids_ref[:, ::-1]
    bnd = np.zeros_like(bad)
    bnd[:, 1:] |= idi[:, 1:] != idi[:, :-1]; bnd[:, :-1] |= idi[:, 1:] != idi[:, :-1]
    bnd[:, :, 1:] |= idi[:, :, 1:] != idi[:, :, :-1]; bnd[:, :, :-1] |= idi[:, :, 1:] != idi[:, :, :-1]
    print(name, "max err", err.max(), "bad", bad.sum(), "bad on id-boundaries", (bad & bnd).sum(), "median", np.median(err[fg]))
    if name == "rgba" and bad.sum():
        idx = np.argwhere(bad)[:8]
        for b, y, x in idx:
            print("   ", b, y, x, "got", g[b, y, x], "ref", r[b, y, x], "id", idi[b, y, x], "cid", planes["cid"][b, y, x, 0].item())
