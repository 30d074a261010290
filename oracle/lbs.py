"""Oracle restatement of FLAME forward: blend shapes + LBS + landmarks.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/vhap/model/lbs.py and vhap/model/flame.py:571-646 line by line in behaviour
(including the 1e-8 added INSIDE the norm in batch_rodrigues, lbs.py:40) but is written dtype-generic
so it can run in float64 and give autograd reference gradients.  Pinned against the reference's own
lbs.py by tests/golden/lbs_golden.npz (tests/test_oracle_golden.py).
"""
import torch


def batch_rodrigues(rot_vecs):
    """lbs.py:25-57.  rot_vecs [N,3] -> [N,3,3]."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)          # lbs.py:40 (quirk kept)
    rot_dir = rot_vecs / angle
    cos = torch.cos(angle)[:, None]
    sin = torch.sin(angle)[:, None]
    rx, ry, rz = rot_dir[:, 0:1], rot_dir[:, 1:2], rot_dir[:, 2:3]
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(-1, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def blend_shapes(betas, shape_disps):
    """lbs.py:218-239: einsum bl,mkl->bmk."""
    return torch.einsum("bl,mkl->bmk", betas, shape_disps)


def vertices2joints(J_regressor, vertices):
    """lbs.py:198-215."""
    return torch.einsum("bik,ji->bjk", vertices, J_regressor)


def batch_rigid_transform(rot_mats, joints, parents):
    """lbs.py:254-304.  rot_mats [B,J,3,3], joints [B,J,3] -> posed joints [B,J,3], rel transforms [B,J,4,4]."""
    B, J = joints.shape[:2]
    rel = joints.clone()
    rel[:, 1:] = joints[:, 1:] - joints[:, parents[1:]]
    T = torch.zeros(B, J, 4, 4, dtype=joints.dtype, device=joints.device)
    T[:, :, :3, :3] = rot_mats
    T[:, :, :3, 3] = rel
    T[:, :, 3, 3] = 1
    chain = [T[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[int(parents[i])], T[:, i]))
    G = torch.stack(chain, dim=1)
    posed = G[:, :, :3, 3]
    jh = torch.cat([joints, torch.zeros_like(joints[..., :1])], -1)[..., None]     # [B,J,4,1]
    corr = torch.matmul(G, jh)                                                      # [B,J,4,1]
    A = G.clone()
    A[:, :, :, 3] = G[:, :, :, 3] - corr[..., 0]
    return posed, A


def lbs(pose, v_shaped, posedirs, J_regressor, parents, lbs_weights):
    """lbs.py:101-195.  pose [B,15] axis-angle (root, neck, jaw, eye_l, eye_r)."""
    B = pose.shape[0]
    J = vertices2joints(J_regressor, v_shaped)
    rot = batch_rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)
    ident = torch.eye(3, dtype=pose.dtype, device=pose.device)
    pose_feature = (rot[:, 1:] - ident).reshape(B, -1)
    v_posed = torch.matmul(pose_feature, posedirs).view(B, -1, 3) + v_shaped
    J_t, A = batch_rigid_transform(rot, J, parents)
    T = torch.matmul(lbs_weights[None].expand(B, -1, -1), A.view(B, -1, 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones_like(v_posed[..., :1])], -1)
    verts = torch.matmul(T, vh[..., None])[:, :, :3, 0]
    return verts, J_t, A[:, 1]


def vertices2landmarks(vertices, faces, lmk_faces_idx, lmk_bary):
    """lbs.py:60-98 (shared embedding for the whole batch, flame.py:634-640)."""
    tri = faces[lmk_faces_idx.long()].long()            # [L,3]
    lv = vertices[:, tri]                               # [B,L,3,3]
    return torch.einsum("blfi,lf->bli", lv, lmk_bary)


def flame_forward(model, shape, expr, rotation, neck, jaw, eyes, translation,
                  static_offset=None, dynamic_offset=None, zero_centered_at_root_node=False):
    """flame.py:571-646.  `model` = dict of tensors (v_template, shapedirs, posedirs, J_regressor,
    parents, lbs_weights, faces, lmk_faces_idx, lmk_bary).  Returns verts, v_shaped (verts_cano), lmks."""
    betas = torch.cat([shape, expr], dim=1)
    full_pose = torch.cat([rotation, neck, jaw, eyes], dim=1)
    v_shaped = model["v_template"][None] + blend_shapes(betas, model["shapedirs"])
    if static_offset is not None:
        v_shaped = v_shaped + static_offset
    if dynamic_offset is not None:
        v_shaped = v_shaped + dynamic_offset
    verts, J, _ = lbs(full_pose, v_shaped, model["posedirs"], model["J_regressor"], model["parents"], model["lbs_weights"])
    if zero_centered_at_root_node:
        verts = verts - J[:, [0]]
    verts = verts + translation[:, None, :]
    lmks = vertices2landmarks(verts, model["faces"], model["lmk_faces_idx"], model["lmk_bary"])
    return verts, v_shaped, lmks


def model_tensors(m, dtype=torch.float64, device="cpu"):
    """FlameModelData (numpy) -> dict of torch tensors for the oracle."""
    t = lambda a: torch.as_tensor(a, dtype=dtype, device=device)
    return {
        "v_template": t(m.v_template), "shapedirs": t(m.shapedirs), "posedirs": t(m.posedirs),
        "J_regressor": t(m.J_regressor), "parents": torch.as_tensor(m.parents),
        "lbs_weights": t(m.lbs_weights), "faces": torch.as_tensor(m.faces.astype("int64"), device=device),
        "faces_uv": torch.as_tensor(m.faces_uv.astype("int64"), device=device),
        "verts_uv": t(m.verts_uv),
        "lmk_faces_idx": torch.as_tensor(m.lmk_faces_idx.astype("int64"), device=device), "lmk_bary": t(m.lmk_bary),
    }
