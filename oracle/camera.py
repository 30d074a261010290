"""Oracle restatement of the camera transforms.  TEST INFRASTRUCTURE ONLY.
Follows vhap/util/render_nvdiffrast.py:117-214 and vhap/util/mesh.py:41-51."""
import torch


def projection_from_intrinsics(K, image_size, near=0.1, far=10.0):
    """render_nvdiffrast.py:117-160.  K [B,3,3] or [B,4]=(fx,fy,cx,cy); image_size=(h,w) -> [B,4,4]."""
    h, w = image_size
    if K.shape[-2:] == (3, 3):
        fx, fy, cx, cy = K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2]
    elif K.shape[-1] == 4:
        fx, fy, cx, cy = K[..., 0], K[..., 1], K[..., 2], K[..., 3]
    else:
        raise ValueError(f"Expected K to be (N, 3, 3) or (N, 4) but got: {K.shape}")
    B = K.shape[0]
    proj = torch.zeros(B, 4, 4, dtype=K.dtype, device=K.device)
    proj[:, 0, 0] = fx * 2 / w
    proj[:, 1, 1] = fy * 2 / h
    proj[:, 0, 2] = (w - 2 * cx) / w
    proj[:, 1, 2] = (h - 2 * cy) / h
    proj[:, 2, 2] = -(far + near) / (far - near)
    proj[:, 2, 3] = -2 * far * near / (far - near)
    proj[:, 3, 2] = -1
    return proj


def world_to_camera(vtx, RT):
    """render_nvdiffrast.py:162-179.  vtx [B,V,3], RT [B,3,4] -> [B,V,3] (a [B,4,4] RT gives [B,V,4], like the reference)."""
    B = vtx.shape[0]
    RT = RT.to(vtx.dtype).expand(B, -1, -1)
    posw = torch.cat([vtx, torch.ones_like(vtx[..., :1])], -1)
    return torch.bmm(posw, RT.transpose(-1, -2))


def camera_to_clip(vtx_cam, K, image_size):
    """render_nvdiffrast.py:181-197.  vtx_cam [B,V,3] (w = 1 appended) or [B,V,4]."""
    proj = projection_from_intrinsics(K, image_size).to(vtx_cam.dtype)
    posw = torch.cat([vtx_cam, torch.ones_like(vtx_cam[..., :1])], -1) if vtx_cam.shape[-1] == 3 else vtx_cam
    if proj.shape[0] < posw.shape[0]:
        proj = proj.expand(posw.shape[0], -1, -1)
    return torch.bmm(posw, proj.transpose(-1, -2))


def world_to_clip(vtx, RT, K, image_size):
    return camera_to_clip(world_to_camera(vtx, RT), K, image_size)


def world_to_ndc(vtx, RT, K, image_size, flip_y=False):
    """render_nvdiffrast.py:208-214."""
    clip = world_to_clip(vtx, RT, K, image_size)
    ndc = clip[:, :, :3] / clip[:, :, 3:]
    if flip_y:
        ndc = ndc * torch.tensor([1.0, -1.0, 1.0], dtype=ndc.dtype, device=ndc.device)
    return ndc


def normalize_image_points(u, v, resolution):
    """mesh.py:41-51.  resolution = (h, w)."""
    return 2 * (u - resolution[1] / 2.0) / resolution[1], 2 * (v - resolution[0] / 2.0) / resolution[0]
