#!/usr/bin/env python
"""Generate tests/golden/lbs_golden.npz by running the REFERENCE's own functions (imported
unmodified from /root/reference) on seeded inputs.  Authoring-container only; the output is committed.

Reference functions exercised (the importable part of the hot path):
  vhap/model/lbs.py: batch_rodrigues :25, vertices2landmarks :60, lbs :101, vertices2joints :198,
                     blend_shapes :218, batch_rigid_transform :254
  vhap/util/mesh.py: normalize_image_points :41
  vhap/util/vector_ops.py: safe_normalize :13
plus the reference's landmark-energy arithmetic (tracker.py:347-389) and joint prior (tracker.py:650-680)
evaluated with those imported functions and reference-defined weights.
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, "/root/reference")
from vhap.model import lbs as R                      # noqa: E402
from vhap.util.mesh import normalize_image_points    # noqa: E402
from vhap.util import vector_ops as VO               # noqa: E402

OUT = Path(__file__).resolve().parent / "lbs_golden.npz"


def main():
    g = torch.Generator().manual_seed(1234)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    B, V, K, L, Fc = 3, 96, 12, 9, 150
    v_template = rn(V, 3) * 0.1
    shapedirs = rn(V, 3, K) * 0.01
    posedirs = rn(36, V * 3) * 0.01
    J_regressor = torch.softmax(rn(5, V), dim=1)
    lbs_weights = torch.softmax(rn(V, 5) * 2, dim=1)
    parents = torch.tensor([-1, 0, 1, 1, 1])
    faces = torch.randint(0, V, (Fc, 3), generator=g)
    lmk_faces_idx = torch.randint(0, Fc, (1, L), generator=g)
    lmk_bary = torch.softmax(rn(1, L, 3), dim=-1)

    betas = rn(B, K)
    pose = rn(B, 15) * 0.3
    pose[0, 3:6] = 0.0                      # exercise the 1e-8 quirk of batch_rodrigues at zero rotation
    transl = rn(B, 3) * 0.05
    offset = rn(1, V, 3) * 0.002

    out = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        c = lambda x: x.to(dt)
        v_shaped = c(v_template)[None] + R.blend_shapes(c(betas), c(shapedirs)) + c(offset)
        verts, J, A1 = R.lbs(c(pose), v_shaped, c(posedirs), c(J_regressor), parents, c(lbs_weights), dtype=dt)
        verts = verts + c(transl)[:, None]
        lm = R.vertices2landmarks(verts, faces, lmk_faces_idx.repeat(B, 1), c(lmk_bary).repeat(B, 1, 1))
        rot = R.batch_rodrigues(c(pose).view(-1, 3), dtype=dt)
        out[f"{name}_v_shaped"] = v_shaped.numpy()
        out[f"{name}_verts"] = verts.numpy()
        out[f"{name}_joints"] = J.numpy()
        out[f"{name}_A1"] = A1.numpy()
        out[f"{name}_lmks"] = lm.numpy()
        out[f"{name}_rot"] = rot.numpy()
    # gradient golden (float64 autograd through the reference functions)
    b = betas.clone().requires_grad_(True)
    p = pose.clone().requires_grad_(True)
    t = transl.clone().requires_grad_(True)
    o = offset.clone().requires_grad_(True)
    v_shaped = v_template[None] + R.blend_shapes(b, shapedirs) + o
    verts, J, _ = R.lbs(p, v_shaped, posedirs, J_regressor, parents, lbs_weights, dtype=torch.float64)
    verts = verts + t[:, None]
    lm = R.vertices2landmarks(verts, faces, lmk_faces_idx.repeat(B, 1), lmk_bary.repeat(B, 1, 1))
    wv = rn(B, V, 3)
    wl = rn(B, L, 3)
    loss = (verts * wv).sum() + (lm * wl).sum()
    loss.backward()
    out.update(g_wv=wv.numpy(), g_wl=wl.numpy(), g_betas=b.grad.numpy(), g_pose=p.grad.numpy(),
               g_transl=t.grad.numpy(), g_offset=o.grad.numpy())

    # normalize_image_points + safe_normalize
    u, v = rn(5, 7) * 100 + 128, rn(5, 7) * 100 + 128
    un, vn = normalize_image_points(u, v, (240, 320))
    out.update(nip_u=u.numpy(), nip_v=v.numpy(), nip_un=un.numpy(), nip_vn=vn.numpy())
    x = rn(6, 3)
    x[0] = 0
    out.update(sn_x=x.numpy(), sn_y=VO.safe_normalize(x).numpy())

    # joint prior (tracker.py:650-680), using the imported batch_rodrigues
    neck, jawp, eyes = rn(B, 3) * 0.2, rn(B, 3) * 0.2, rn(B, 6) * 0.2
    w = {"neck": 3e-1, "jaw": 3e-1, "eyes": 3e-2}      # base.py:135-138
    E = 0
    for nm, ps in (("neck", neck), ("jaw", jawp), ("eyes", eyes[:, :3]), ("eyes", eyes[:, 3:])):
        rm = R.batch_rodrigues(torch.cat([torch.zeros_like(ps), ps], 0), dtype=torch.float64)
        diff = ((rm[[0]] - rm[1:]) ** 2).mean()
        if nm == "jaw":
            diff = diff + F.relu(-ps[:, 0]).mean() * 10 + (ps[:, 1:] ** 2).mean() * 3
        elif nm == "eyes":
            diff = diff + ((eyes[:, :3] - eyes[:, 3:]) ** 2).mean()
        E = E + diff * w[nm]
    out.update(jp_neck=neck.numpy(), jp_jaw=jawp.numpy(), jp_eyes=eyes.numpy(), jp_E=np.asarray(E.item()))

    out.update(v_template=v_template.numpy(), shapedirs=shapedirs.numpy(), posedirs=posedirs.numpy(),
               J_regressor=J_regressor.numpy(), lbs_weights=lbs_weights.numpy(), faces=faces.numpy(),
               lmk_faces_idx=lmk_faces_idx.numpy(), lmk_bary=lmk_bary.numpy(), betas=betas.numpy(),
               pose=pose.numpy(), transl=transl.numpy(), offset=offset.numpy())
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size)


if __name__ == "__main__":
    main()
