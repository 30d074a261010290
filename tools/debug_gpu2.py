import sys, copy, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
from tests.scene import make_scene
from vhap_b200.engine import Engine
from oracle import energy as E, lbs as L

def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)

sc = make_scene(B=3, H=96, W=96, T=256, n_t=4, timesteps=[1, 2, 1])
e = Engine(sc["m"], sc["cfg"], 4, tex_painted=sc["tex_painted"])
e.load_params(sc["params"])
ts = sc["ts"]; B = 3; H = W = 96; model = sc["model"]
P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in sc["params"].items()}
verts, _, lm = L.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts], P["jaw_pose"][ts],
                               P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
lm.retain_grad()
K, RT = E.fill_cam_params(P, B, H, W)
loss = 10.0 * E.lmk_energy(lm, torch.tensor(sc["lmk2d"]), K, RT, (H, W), True, False)
loss.backward()
g_lm = lm.grad.clone()
batch = e.stage_sample(sc["rgb16"].to(torch.float32), sc["lmk2d"], ts)
cp = e._c_params()
dv = torch.empty(B, e.V, 3, device=e.dev); dl = torch.empty(B, 70, 3, device=e.dev)
e._ck(e.L.vhap_flame_forward(e.ctx, C.byref(cp), C.byref(batch.c), dv.data_ptr(), None, dl.data_ptr(), e._stream()))
print("lmk fwd rel", rel(dl.cpu().numpy(), lm.detach().numpy()))
e.zero_grad()
opt = {k: True for k in ("cam", "pose", "shape", "joints", "expr", "texture", "lights", "static_offset")}; opt["texture"] = False
cg = e._c_grads(opt)
gl = g_lm.to(torch.float32).to(e.dev).contiguous()
e._ck(e.L.vhap_flame_backward(e.ctx, C.byref(cp), C.byref(batch.c), None, gl.data_ptr(), C.byref(cg), e._stream()))
torch.cuda.synchronize()
print("sparse g_lmks through flame_backward:", {k: float("%.3g" % rel(e.g[k].cpu().numpy(), P[k].grad.numpy().reshape(-1)))
      for k in ("shape", "expr", "rotation", "translation", "neck_pose", "jaw_pose", "eyes_pose", "static_offset")})
# ---- energy path, landmark term only
import copy
from vhap_b200.config import STAGES
cfg = copy.deepcopy(sc["cfg"]); w = cfg.w
w.reg_shape = w.reg_expr = w.reg_neck = w.reg_jaw = w.reg_eyes = 0.0
e.cfg = cfg
e.set_stage(STAGES["lmk_init_all"])
e.zero_grad(); e.energy(batch, True, True); torch.cuda.synchronize()
for k in ("focal_length", "translation", "rotation"):
    print(k, "got", e.g[k].cpu().numpy()[:9], "\n   ref", P[k].grad.numpy().reshape(-1)[:9])
print(e.loss_dict()["lmk"], loss.item())
