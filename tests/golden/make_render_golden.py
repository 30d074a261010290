#!/usr/bin/env python
"""Generate tests/golden/render_golden.npz from the pure-PyTorch METHODS OF THE REFERENCE'S OWN RENDERER
(vhap/util/render_nvdiffrast.py, imported unmodified): camera (projection_from_intrinsics / world_to_camera / camera_to_clip /
world_to_clip / world_to_ndc, :117-214), vertex normals (compute_v_normals, :297-316), SH shading (shade, :19-53,332-347) and
detach_by_indices (:349-352).  The module's only missing dependency is `nvdiffrast`, whose functions these methods never call;
a stub module with a dummy rasteriser context is registered in sys.modules so that the file imports as it is.  (The
rasterise / interpolate / texture / antialias calls themselves stay unpinned: they are nvdiffrast.)

    PYTHONPATH=/root/reference python tests/golden/make_render_golden.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

stub = types.ModuleType("nvdiffrast")
stub_t = types.ModuleType("nvdiffrast.torch")
stub_t.RasterizeCudaContext = lambda *a, **k: object()
stub_t.RasterizeGLContext = lambda *a, **k: object()
stub.torch = stub_t
sys.modules["nvdiffrast"], sys.modules["nvdiffrast.torch"] = stub, stub_t
sys.path.insert(0, "/root/reference")
from vhap.util.render_nvdiffrast import NVDiffRenderer      # noqa: E402

# the reference hard-codes `.cuda()` on the homogeneous-coordinate columns (render_nvdiffrast.py:174,…); this container has no GPU,
# so inside THIS process `.cuda()` is made the identity -- the reference file itself stays untouched
torch.Tensor.cuda = lambda self, *a, **k: self
_torch_tensor = torch.tensor


def _tensor_on_cpu(*a, **k):                      # `torch.tensor(..., device='cuda')` constants (render_nvdiffrast.py:312)
    if k.get("device") == "cuda":
        k.pop("device")
    return _torch_tensor(*a, **k)


torch.tensor = _tensor_on_cpu


def main():
    g = torch.Generator().manual_seed(11)
    rnd = NVDiffRenderer(use_opengl=False, lighting_type="SH", lighting_space="world")
    B, V, F, H, W = 3, 40, 60, 48, 80
    out = {}
    verts = (torch.randn(B, V, 3, generator=g) * 0.3).requires_grad_(True)
    faces = torch.stack([torch.randperm(V, generator=g)[:3] for _ in range(F)]).long()
    faces[5] = torch.tensor([7, 7, 9])                   # degenerate face
    verts_n = verts.detach().clone()
    # a vertex used by no face -> zero normal -> the (0,0,1) fallback of compute_v_normals
    lonely = [v for v in range(V) if v not in set(faces.reshape(-1).tolist())]
    RT = torch.eye(3, 4)[None].repeat(B, 1, 1)
    RT[:, :3, :3] += 0.05 * torch.randn(B, 3, 3, generator=g)
    RT[:, 2, 3] = -1.0 + 0.1 * torch.randn(B, generator=g)
    # (fx, fy, cx, cy): the [1, 4] form the tracker builds for uncalibrated cameras (tracker.py:141-157); the reference's own
    # projection_from_intrinsics only accepts N = 1 in this form (its [N,1] splits do not broadcast into [N])
    K4 = torch.tensor([[1.5 * 80, 1.4 * 80, 39.0, 25.0]])
    K3 = torch.zeros(B, 3, 3); K3[:, 0, 0] = K4[0, 0]; K3[:, 1, 1] = K4[0, 1]; K3[:, 0, 2] = K4[0, 2]; K3[:, 1, 2] = K4[0, 3]; K3[:, 2, 2] = 1
    K3[1, 0, 0] *= 1.1; K3[2, 1, 2] += 3.0                                               # per-frame intrinsics in the [B,3,3] form
    out.update(verts=verts.detach().numpy(), faces=faces.numpy(), RT=RT.numpy(), K4=K4.numpy(), K3=K3.numpy(), image_size=np.array([H, W]),
               lonely=np.array(lonely))
    for name, K in (("K4", K4), ("K3", K3)):
        out[f"proj_{name}"] = rnd.projection_from_intrinsics(K, (H, W)).numpy()
    cam = rnd.world_to_camera(verts, RT)
    clip = rnd.camera_to_clip(cam, K3, (H, W))
    clip2 = rnd.world_to_clip(verts, RT, K4, (H, W))                                     # [1,4] intrinsics shared by the batch
    ndc = rnd.world_to_ndc(verts, RT, K3, (H, W), flip_y=True)
    ndc_nf = rnd.world_to_ndc(verts, RT, K3, (H, W), flip_y=False)
    wts = torch.randn(B, V, 4, generator=g)
    (clip * wts).sum().backward()
    out.update(cam=cam.detach().numpy(), clip=clip.detach().numpy(), clip_w2c=clip2.detach().numpy(), ndc_flip=ndc.detach().numpy(),
               ndc=ndc_nf.detach().numpy(), w_clip=wts.numpy(), g_verts_from_clip=verts.grad.numpy().copy())
    # vertex normals + their gradient
    vn_in = verts_n.clone().requires_grad_(True)
    vn = rnd.compute_v_normals(vn_in, faces)
    wn = torch.randn(B, V, 3, generator=g)
    (vn * wn).sum().backward()
    out.update(v_normals=vn.detach().numpy(), w_vn=wn.numpy(), g_verts_from_vn=vn_in.grad.numpy())
    # SH shading + gradients to normals and lights
    normal = torch.nn.functional.normalize(torch.randn(B, 6, 5, 3, generator=g), dim=-1).requires_grad_(True)
    lights = torch.zeros(1, 9, 3); lights[0, 0] = float(np.sqrt(4 * np.pi)); lights[0, 1:] = 0.2 * torch.randn(8, 3, generator=g)
    lights.requires_grad_(True)
    diffuse = rnd.shade(normal, lights)
    wd = torch.randn(B, 6, 5, 3, generator=g)
    (diffuse * wd).sum().backward()
    out.update(sh_normal=normal.detach().numpy(), sh_lights=lights.detach().numpy(), sh_diffuse=diffuse.detach().numpy(), w_sh=wd.numpy(),
               g_normal=normal.grad.numpy(), g_lights=lights.grad.numpy(), sh_const=rnd.sh_const.numpy())
    # detach_by_indices
    x = torch.randn(B, V, 4, generator=g).requires_grad_(True)
    idx = torch.tensor([1, 4, 9])
    y = rnd.detach_by_indices(x, idx)
    y.sum().backward()
    out.update(dbi_idx=idx.numpy(), dbi_grad=x.grad.numpy())
    path = Path(__file__).with_name("render_golden.npz")
    np.savez_compressed(path, **out)
    print(path, sorted(out))


if __name__ == "__main__":
    main()
