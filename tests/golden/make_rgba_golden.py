#!/usr/bin/env python
"""Generate tests/golden/rgba_golden.npz by running the REFERENCE's own NVDiffRenderer.render_rgba
(vhap/util/render_nvdiffrast.py:354-484, imported unmodified) end to end, with its four nvdiffrast calls (dr.interpolate x2,
dr.texture, dr.antialias) served by this repo's oracle restatements of those ops, and its random draws (torch.rand_like x2,
torch.randint per cluster) replaced inside the call by the injected Bernoulli / uniform planes the oracle and the CUDA tests use.

What this pins (against the reference's code, values AND gradients): everything in render_rgba that is NOT nvdiffrast --
vertex normals -> interpolate -> safe_normalize, the per-face texc.detach() mask of align_texture_except_fid (:390-396), SH shading
evaluated twice (normal / normal.detach, :402-403), rgb = albedo * diffuse, alpha, constant / image background incl. its vertical
flip (:407-421), the whole cluster-disturbance block (:424-460: pools per cluster over the batch, detached, cluster 0 from the
background, cluster 1 passed through, blend by w), detach_by_indices before the antialias (:463-464), the `aa` mask (:466) and the
output flips (:476-483).  The dr.* ops themselves are the oracle's on both sides and stay UNPINNED.

    PYTHONPATH=/root/reference python tests/golden/make_rgba_golden.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")
from oracle import camera as Cm, energy as OE, lbs as OL, raster as RA, render as RE    # noqa: E402
from tests.scene import make_scene                                                     # noqa: E402

STATE = {}


def dr_interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    out, da = RE.interpolate(attr, rast, tri.long(), rast_db)
    return out, da


def dr_texture(tex, uv, uv_da, filter_mode=None, max_mip_level=None, boundary_mode="wrap"):
    assert filter_mode == "linear-mipmap-linear"
    B, H, W, _ = uv.shape
    mips = RE.build_mips(tex[0])                                     # the tracker expands ONE texture to B copies (tracker.py:234)
    out = RE.texture_sample(mips, uv.reshape(-1, 2), uv_da.reshape(-1, 4))
    return out.reshape(B, H, W, -1)


def dr_antialias(color, rast, pos, tri):
    return RE.antialias(color, rast, pos, tri.long(), STATE["adj_opp"])


stub_t = types.ModuleType("nvdiffrast.torch")
stub_t.RasterizeCudaContext = lambda *a, **k: object()
stub_t.RasterizeGLContext = lambda *a, **k: object()
stub_t.interpolate, stub_t.texture, stub_t.antialias = dr_interpolate, dr_texture, dr_antialias
stub = types.ModuleType("nvdiffrast"); stub.torch = stub_t
sys.modules["nvdiffrast"], sys.modules["nvdiffrast.torch"] = stub, stub_t
from vhap.util.render_nvdiffrast import NVDiffRenderer      # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self
_torch_tensor = torch.tensor


def _tensor_on_cpu(*a, **k):
    if k.get("device") == "cuda":
        k.pop("device")
    return _torch_tensor(*a, **k)


torch.tensor = _tensor_on_cpu


class InjectedRandom:
    """inside render_rgba: rand_like #1 -> w_fg, #2 -> w_bg (as uniforms below / above the rate), randint(0, n) -> min(int(u * n), n - 1)"""

    def __init__(self, w_fg, w_bg, u):
        self.planes = [w_fg, w_bg]
        self.u = u.reshape(-1).to(torch.float32)

    def __enter__(self):
        self._rl, self._ri = torch.rand_like, torch.randint

        def rand_like(x, *a, **k):
            w = self.planes.pop(0)
            return (1.0 - w.to(x.dtype)).reshape(x.shape)            # 0 < rate where the draw says "take a pool sample", 1 >= rate otherwise

        def randint(low, high, size, **k):
            return (self.u * torch.tensor(float(high), dtype=torch.float32)).to(torch.int64).clamp(max=high - 1)

        torch.rand_like, torch.randint = rand_like, randint
        return self

    def __exit__(self, *exc):
        torch.rand_like, torch.randint = self._rl, self._ri


def main():
    sc = make_scene(B=2, H=28, W=36, T=32, n_t=3, timesteps=[0, 2])
    m, model = sc["m"], sc["model"]
    STATE["adj_opp"] = m.face_adjacency_opposite()
    B, H, W = sc["B"], sc["H"], sc["W"]
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in sc["params"].items()}
    ts = torch.as_tensor(sc["ts"]).long()
    with torch.no_grad():
        v64, _, _ = OL.flame_forward(model, P["shape"][None].expand(B, -1), P["expr"][ts], P["rotation"][ts], P["neck_pose"][ts], P["jaw_pose"][ts],
                                     P["eyes_pose"][ts], P["translation"][ts], static_offset=P["static_offset"])
        K, RTm = OE.fill_cam_params(P, B, H, W)
    f32 = torch.float32
    verts0 = v64.to(f32)
    faces = model["faces"]
    faces_uv = model["faces_uv"]
    verts_uv = model["verts_uv"].to(f32).clone(); verts_uv[:, 1] = 1 - verts_uv[:, 1]          # tracker.py:315-316
    clip0 = Cm.world_to_clip(verts0, RTm.to(f32), K.to(f32), (H, W))
    rast, rast_db = RA.rasterize(clip0, faces, (H, W))
    rast, rast_db = rast.detach().to(f32), rast_db.detach().to(f32)
    tex0 = torch.tensor(sc["tex_painted"] + sc["params"]["tex_extra"], dtype=f32)               # [3,T,T]
    lights0 = P["lights"].to(f32)
    bg_img = sc["rgb16"].to(f32).permute(0, 2, 3, 1).contiguous()
    fid2cid = torch.as_tensor(m.fid2cid(sc["cfg"].tex_clusters))                               # [F], padded by the renderer itself
    tex_exc = torch.as_tensor(m.get_fid_by_region(["hair", "boundary", "neck"]))
    bnd_exc = torch.as_tensor(m.get_vid_by_region(["hair", "bottomline"]))
    g = torch.Generator().manual_seed(5)
    w_out = torch.randn(B, H, W, 4, generator=g)
    out = dict(verts=verts0.numpy(), clip=clip0.numpy(), rast=rast.numpy(), rast_db=rast_db.numpy(), tex=tex0.numpy(), lights=lights0.numpy(),
               bg_img=bg_img.numpy(), verts_uv=verts_uv.numpy(), w_fg=sc["w_fg"].numpy(), w_bg=sc["w_bg"].numpy(), u_rand=sc["u_rand"].numpy(),
               tex_exc=tex_exc.numpy(), bnd_exc=bnd_exc.numpy(), w_out=w_out.numpy(), fid2cid=fid2cid.numpy())
    cases = {"plain_imgbg": dict(bg="img", disturb=False, exc=False), "disturb_exc_imgbg": dict(bg="img", disturb=True, exc=True),
             "disturb_white": dict(bg=[1.0, 1.0, 1.0], disturb=True, exc=False)}
    rnd = NVDiffRenderer(use_opengl=False, lighting_type="SH", lighting_space="world", disturb_rate_fg=0.5, disturb_rate_bg=0.5, fid2cid=fid2cid)
    for name, c in cases.items():
        verts = verts0.clone().requires_grad_(True)
        clip = clip0.detach().clone().requires_grad_(True)
        tex = tex0.clone().requires_grad_(True)
        lights = lights0.clone().requires_grad_(True)
        rast_dict = {"rast_out": rast, "rast_out_db": rast_db, "verts": verts, "verts_camera": None, "verts_clip": clip}
        bg = bg_img if c["bg"] == "img" else c["bg"]
        with InjectedRandom(sc["w_fg"], sc["w_bg"], sc["u_rand"]):
            res = rnd.render_rgba(rast_dict, verts, faces, verts_uv, faces_uv, tex[None].expand(B, -1, -1, -1), lights[None], bg,
                                  tex_exc if c["exc"] else None, bnd_exc if c["exc"] else None, c["disturb"])
        (res["rgba"] * w_out).sum().backward()
        for k in ("rgba", "albedo", "normal", "diffuse", "diffuse_detach_normal", "aa"):
            out[f"{name}/{k}"] = res[k].detach().numpy()
        if "cid" in res:
            out[f"{name}/cid"] = res["cid"].numpy()
        out[f"{name}/g_verts"] = verts.grad.numpy(); out[f"{name}/g_clip"] = clip.grad.numpy()
        out[f"{name}/g_tex"] = tex.grad.numpy(); out[f"{name}/g_lights"] = lights.grad.numpy()
        print(name, "fg px", int((rast[..., 3] > 0).sum()), "aa px", int(res["aa"][..., 0].sum()), "|g_clip|", float(clip.grad.abs().sum()))
    path = Path(__file__).with_name("rgba_golden.npz")
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size)


if __name__ == "__main__":
    main()
