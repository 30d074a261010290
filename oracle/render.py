"""Oracle renderer: vertex normals, interpolate, mip-mapped texture, SH shading, composite, cluster colour
disturbance, antialias.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against nvdiffrast (oracle/__init__.py).

Follows vhap/util/render_nvdiffrast.py:297-316 (normals), :19-53,:83-96,:332-347 (SH), :354-484 (render_rgba
graph: order of ops, flips, detaches, disturbance) and restates the nvdiffrast ops called at :384,:389 (interpolate),
:399 (texture, linear-mipmap-linear) and :465 (antialias) from the published algorithm.
All tensors here are in nvdiffrast orientation (row 0 = bottom); `render_rgba` flips its outputs like the reference.
"""
import math

import torch

from .raster import shade_pass  # noqa: F401  (re-export)


def dot(x, y):
    return torch.sum(x * y, -1, keepdim=True)


def safe_normalize(x, eps=1e-20):
    """vector_ops.py:10-14."""
    return x / torch.sqrt(torch.clamp(dot(x, x), min=eps))


def compute_v_normals(verts, faces):
    """render_nvdiffrast.py:297-316: area-weighted vertex normals, fallback (0,0,1), safe_normalize."""
    i0, i1, i2 = faces[:, 0].long(), faces[:, 1].long(), faces[:, 2].long()
    v0, v1, v2 = verts[:, i0], verts[:, i1], verts[:, i2]
    fn = torch.cross(v1 - v0, v2 - v0, dim=-1)
    vn = torch.zeros_like(verts)
    vn = vn.index_add(1, i0, fn).index_add(1, i1, fn).index_add(1, i2, fn)
    fb = torch.tensor([0.0, 0.0, 1.0], dtype=verts.dtype, device=verts.device)
    vn = torch.where(dot(vn, vn) > 1e-20, vn, fb)
    return safe_normalize(vn)


SH_CONST = [
    1 / math.sqrt(4 * math.pi),
    ((2 * math.pi) / 3) * math.sqrt(3 / (4 * math.pi)),
    ((2 * math.pi) / 3) * math.sqrt(3 / (4 * math.pi)),
    ((2 * math.pi) / 3) * math.sqrt(3 / (4 * math.pi)),
    (math.pi / 4) * 3 * math.sqrt(5 / (12 * math.pi)),
    (math.pi / 4) * 3 * math.sqrt(5 / (12 * math.pi)),
    (math.pi / 4) * 3 * math.sqrt(5 / (12 * math.pi)),
    (math.pi / 4) * (3 / 2) * math.sqrt(5 / (12 * math.pi)),
    (math.pi / 4) * (1 / 2) * math.sqrt(5 / (4 * math.pi)),
]  # render_nvdiffrast.py:83-96


def sh_shading(normal, lights):
    """render_nvdiffrast.py:19-53.  normal [...,3], lights [9,3] -> [...,3]."""
    n = normal
    x, y, z = n[..., 0], n[..., 1], n[..., 2]
    basis = torch.stack([x * 0 + 1, x, y, z, x * y, x * z, y * z, x * x - y * y, 3 * z * z - 1], -1)
    basis = basis * torch.tensor(SH_CONST, dtype=n.dtype, device=n.device)
    return basis @ lights


def interpolate(attr, rast, tri, rast_db=None):
    """dr.interpolate.  attr [1|B,V,A]; returns out [B,H,W,A] and (if rast_db) out_da [B,H,W,2A] laid out
    (d a_k/dx, d a_k/dy) per attribute k (diff_attrs='all')."""
    B, H, W, _ = rast.shape
    ids = rast[..., 3].long()
    fg = ids > 0
    t = (ids - 1).clamp(min=0)
    vi = tri.long()[t]                                              # [B,H,W,3]
    if attr.shape[0] == 1:
        a = attr[0][vi]                                             # [B,H,W,3,A]
    else:
        bidx = torch.arange(B, device=rast.device)[:, None, None, None]
        a = attr[bidx, vi]
    u, v = rast[..., 0:1], rast[..., 1:2]
    out = u * a[..., 0, :] + v * a[..., 1, :] + (1 - u - v) * a[..., 2, :]
    out = torch.where(fg[..., None], out, torch.zeros_like(out))
    if rast_db is None:
        return out, None
    d0, d1 = a[..., 0, :] - a[..., 2, :], a[..., 1, :] - a[..., 2, :]
    dadx = rast_db[..., 0:1] * d0 + rast_db[..., 2:3] * d1
    dady = rast_db[..., 1:2] * d0 + rast_db[..., 3:4] * d1
    da = torch.stack([dadx, dady], -1).reshape(B, H, W, -1)
    da = torch.where(fg[..., None], da, torch.zeros_like(da))
    return out, da


def build_mips(tex):
    """tex [T,T,C] -> list of levels down to 1x1 by 2x2 box averaging (A.3)."""
    levels = [tex]
    while levels[-1].shape[0] > 1:
        t = levels[-1]
        t = 0.25 * (t[0::2, 0::2] + t[1::2, 0::2] + t[0::2, 1::2] + t[1::2, 1::2])
        levels.append(t)
    return levels


def _bilinear_wrap(tex, u, v):
    """tex [h,w,C], u,v [N] (texture coords, any range; wrap) -> [N,C]; texel centres at i+0.5; tex[row=v, col=u]."""
    h, w = tex.shape[:2]
    u = u - torch.floor(u)
    v = v - torch.floor(v)
    x = u * w - 0.5
    y = v * h - 0.5
    x0f, y0f = torch.floor(x), torch.floor(y)
    fx, fy = (x - x0f)[:, None], (y - y0f)[:, None]
    x0 = x0f.long() % w
    y0 = y0f.long() % h
    x1, y1 = (x0 + 1) % w, (y0 + 1) % h
    t00, t10, t01, t11 = tex[y0, x0], tex[y0, x1], tex[y1, x0], tex[y1, x1]
    return (t00 * (1 - fx) + t10 * fx) * (1 - fy) + (t01 * (1 - fx) + t11 * fx) * fy


def mip_level(uv_da, T, max_level):
    """uv_da [N,4] = (du/dx, du/dy, dv/dx, dv/dy) -> continuous level in [0,max_level] (A.3)."""
    dsdx, dsdy, dtdx, dtdy = uv_da[:, 0] * T, uv_da[:, 1] * T, uv_da[:, 2] * T, uv_da[:, 3] * T
    A = dsdx * dsdx + dtdx * dtdx
    Bq = dsdy * dsdy + dtdy * dtdy
    C = dsdx * dsdy + dtdx * dtdy
    l2b = 0.5 * (A + Bq)
    l2n = 0.25 * (A - Bq) * (A - Bq) + C * C
    l2a = torch.where(l2n > 0, torch.sqrt(l2n.clamp_min(1e-38)), torch.zeros_like(l2n))
    major = l2b + l2a
    lvl = 0.5 * torch.log2(major.clamp_min(1e-30))
    return lvl.clamp(0.0, float(max_level))


def texture_sample(mips, uv, uv_da):
    """dr.texture(..., filter_mode='linear-mipmap-linear', boundary_mode='wrap') on the fg pixels.
    mips: list from build_mips (shared texture), uv [N,2], uv_da [N,4] -> [N,C]."""
    T = mips[0].shape[0]
    max_level = len(mips) - 1
    lvl = mip_level(uv_da, T, max_level)
    l0 = torch.floor(lvl).long().clamp(max=max_level)
    l1 = (l0 + 1).clamp(max=max_level)
    f = (lvl - l0.to(lvl.dtype))[:, None]
    out = torch.zeros(uv.shape[0], mips[0].shape[-1], dtype=uv.dtype, device=uv.device)
    for L in range(max_level + 1):
        m0 = l0 == L
        if m0.any():
            idx = torch.nonzero(m0)[:, 0]
            out = out.index_add(0, idx, _bilinear_wrap(mips[L], uv[idx, 0], uv[idx, 1]) * (1 - f[idx]))
        m1 = l1 == L
        if m1.any():
            idx = torch.nonzero(m1)[:, 0]
            out = out.index_add(0, idx, _bilinear_wrap(mips[L], uv[idx, 0], uv[idx, 1]) * f[idx])
    return out


# ---------------------------------------------------------------------------------------------- antialias
def antialias(color, rast, pos, tri, adj_opp):
    """dr.antialias(color, rast, pos, tri).  color [B,H,W,C], rast [B,H,W,4], pos [B,V,4] clip coords (rows the
    caller detached simply get no gradient), tri [F,3], adj_opp [F,3] = opposite vertex across edge k=(v_k,v_k+1)
    in the neighbouring face (-1 boundary, -2 non-manifold) -- the static replacement of nvdiffrast's topology hash.

    For each horizontally / vertically adjacent pixel pair with different triangle ids: choose the nearer surface
    (non-empty beats empty, otherwise smaller z/w, ties -> second pixel); among that triangle's edges, the first
    (order v0v1, v1v2, v2v0) that is a silhouette edge and crosses the segment between the two pixel centres at
    parameter t in [0,1] measured from the NEAR pixel decides: alpha = t - 0.5; alpha > 0 blends the near colour
    into the far pixel by alpha, otherwise the far colour into the near pixel by -alpha (A.4)."""
    B, H, W, C = color.shape
    dt, dev = color.dtype, color.device
    ids = rast[..., 3].long()
    zw = rast[..., 2]
    tri = tri.long()
    adj = torch.as_tensor(adj_opp, device=dev).long()
    out = color.clone()
    flat = color.reshape(-1, C)
    out_flat = out.reshape(-1, C)
    for d in (0, 1):
        if d == 0:
            id0, id1 = ids[:, :, :-1], ids[:, :, 1:]
            z0, z1 = zw[:, :, :-1], zw[:, :, 1:]
        else:
            id0, id1 = ids[:, :-1, :], ids[:, 1:, :]
            z0, z1 = zw[:, :-1, :], zw[:, 1:, :]
        b, y, x = torch.nonzero(id0 != id1, as_tuple=True)
        if b.numel() == 0:
            continue
        i0, i1 = id0[b, y, x] - 1, id1[b, y, x] - 1
        zz0, zz1 = z0[b, y, x], z1[b, y, x]
        use0 = torch.where((i0 >= 0) & (i1 >= 0), zz0 < zz1, i0 >= 0)       # near surface is pixel 0's
        t_id = torch.where(use0, i0, i1)
        x1p, y1p = (x + 1, y) if d == 0 else (x, y + 1)
        nx, ny = torch.where(use0, x, x1p), torch.where(use0, y, y1p)        # near pixel
        fxp, fyp = torch.where(use0, x1p, x), torch.where(use0, y1p, y)      # far pixel
        sgn = torch.where(use0, 1.0, -1.0).to(dt)                            # direction near -> far along axis d
        fx = (2 * nx + 1).to(dt) / W - 1.0
        fy = (2 * ny + 1).to(dt) / H - 1.0
        vi = tri[t_id]                                                       # [N,3]
        p = pos[b[:, None], vi]                                              # [N,3,4]
        qx = p[:, :, 0] / p[:, :, 3] - fx[:, None]
        qy = p[:, :, 1] / p[:, :, 3] - fy[:, None]
        op = adj[t_id]                                                       # [N,3]
        po = pos[b[:, None], op.clamp(min=0)]                                # [N,3,4]
        ox = po[:, :, 0] / po[:, :, 3] - fx[:, None]
        oy = po[:, :, 1] / po[:, :, 3] - fy[:, None]
        if d == 1:                                                           # make the pair axis the "x" axis
            qx, qy, ox, oy = qy, qx, oy, ox
        pitch = (2.0 / W if d == 0 else 2.0 / H) * sgn                       # signed distance to the far centre
        found = torch.zeros_like(use0)
        alpha = torch.zeros(b.shape[0], dtype=dt, device=dev)
        for k in range(3):
            a_, b_, c_ = k, (k + 1) % 3, (k + 2) % 3
            ax, ay, bx, by = qx[:, a_], qy[:, a_], qx[:, b_], qy[:, b_]
            cross = (ay > 0) != (by > 0)
            den = by - ay
            den_s = torch.where(cross, den, torch.ones_like(den))
            xc = (ax * by - ay * bx) / den_s
            tpar = xc / pitch
            hit = cross & (tpar >= 0) & (tpar <= 1)
            # silhouette test
            ex, ey = bx - ax, by - ay
            side_c = ex * (qy[:, c_] - ay) - ey * (qx[:, c_] - ax)
            side_o = ex * (oy[:, k] - ay) - ey * (ox[:, k] - ax)
            sil = (op[:, k] == -1) | ((op[:, k] >= 0) & (side_c * side_o > 0))
            take = hit & sil & ~found
            alpha = torch.where(take, tpar - 0.5, alpha)
            found = found | take
        sel = torch.nonzero(found)[:, 0]
        if sel.numel() == 0:
            continue
        al = alpha[sel][:, None]
        lin_n = (b[sel] * H + ny[sel]) * W + nx[sel]
        lin_f = (b[sel] * H + fyp[sel]) * W + fxp[sel]
        cn, cf = flat[lin_n], flat[lin_f]
        pos_a = (al > 0).to(dt)
        out_flat = out_flat.index_add(0, lin_f, pos_a * al * (cn - cf))
        out_flat = out_flat.index_add(0, lin_n, (1 - pos_a) * (-al) * (cf - cn))
    return out_flat.reshape(B, H, W, C)


# ---------------------------------------------------------------------------------------------- render_rgba
def disturb(rgba, rgba_bg, cid, num_clusters, w_fg, w_bg, u_rand):
    """render_nvdiffrast.py:424-460 with INJECTED randomness.  cid [B,H,W] long; w_fg/w_bg [B,H,W] bool (the
    Bernoulli draws of :429-435); u_rand [B,H,W] float32 in [0,1): pixel p of cluster i samples pool index
    min(int(fp32(u*n_i)), n_i-1) where the pool is the (detached) cluster-i pixels of the whole batch in (b,y,x)
    order -- the reference draws an independent randint per cluster (:455), identical in distribution."""
    B, H, W, C = rgba.shape
    out = torch.zeros_like(rgba)
    flat_c = cid.reshape(-1)
    u32 = u_rand.reshape(-1).to(torch.float32)
    for i in range(num_clusters):
        c_rgba = rgba_bg if i == 0 else rgba
        w = (w_bg if i == 0 else w_fg).reshape(-1)
        m = flat_c == i
        if i == 1:
            out = out + m.reshape(B, H, W, 1) * c_rgba
            continue
        pool = c_rgba.reshape(-1, C)[m].detach()
        n = pool.shape[0]
        if n == 0:
            continue
        idx = (u32 * torch.tensor(float(n), dtype=torch.float32)).to(torch.int64).clamp(max=n - 1)
        sample = pool[idx]
        wf = w.to(rgba.dtype)[:, None]
        val = sample * wf + c_rgba.reshape(-1, C) * (1 - wf)
        out = out + (m[:, None] * val).reshape(B, H, W, C)
    return out


def render_rgba(rast, rast_db, verts, verts_clip, faces, verts_uv, faces_uv, tex, lights, background,
                adj_opp, fid2cid_padded=None, align_texture_except_fid=None, align_boundary_except_vid=None,
                disturbance=None, v_normal=None):
    """NVDiffRenderer.render_rgba (render_nvdiffrast.py:354-484), lighting_type='SH', lighting_space='world'.
    tex [3,T,T] (single shared texture; the reference expands it to B copies, tracker.py:234), verts_uv already has
    v flipped by the caller (tracker.py:315-316).  background: [B,H,W,3] image tensor (image orientation) or list.
    disturbance: None or dict(w_fg, w_bg, u_rand).  Returns dict of [B,H,W,C] tensors, flipped to image orientation."""
    B, H, W, _ = rast.shape
    dt = rast.dtype
    ids = rast[..., 3].long()
    fg = ids > 0
    if v_normal is None:                                     # (override only used by unit tests of the pixel math)
        v_normal = compute_v_normals(verts, faces)
    normal, _ = interpolate(v_normal, rast, faces)
    normal = safe_normalize(normal)
    texc, texd = interpolate(verts_uv[None], rast, faces_uv, rast_db)
    if align_texture_except_fid is not None:
        mask = torch.zeros(faces.shape[0] + 1, dtype=torch.bool, device=rast.device)
        mask[torch.as_tensor(align_texture_except_fid).long() + 1] = True
        texc = torch.where(mask[ids][..., None], texc.detach(), texc)          # :390-396 (texd is NOT detached)
    mips = build_mips(tex.permute(1, 2, 0))
    bi, yi, xi = torch.nonzero(fg, as_tuple=True)
    alb_fg = texture_sample(mips, texc[bi, yi, xi], texd[bi, yi, xi])
    # nvdiffrast samples every pixel; empty pixels have uv = 0, uv_da = 0 -> level 0 at uv (0,0)
    alb_bg = _bilinear_wrap(mips[0], torch.zeros(1, dtype=dt, device=rast.device), torch.zeros(1, dtype=dt, device=rast.device))[0]
    albedo = alb_bg.expand(B, H, W, 3).clone().index_put((bi, yi, xi), alb_fg)
    diffuse = sh_shading(normal, lights)
    diffuse_detach_normal = sh_shading(normal.detach(), lights)
    rgb = albedo * diffuse
    rgba = torch.cat([rgb, fg[..., None].to(dt)], -1)
    if isinstance(background, (list, tuple)):
        rgba_bg = torch.tensor(list(background) + [0], dtype=dt, device=rast.device).expand(B, H, W, 4)
    elif isinstance(background, torch.Tensor):
        rgba_bg = torch.cat([background.to(dt), torch.zeros_like(background[..., :1], dtype=dt)], -1)
    else:
        raise ValueError(f"Unknown background type: {type(background)}")
    rgba_bg = rgba_bg.flip(1)
    rgba = torch.where(fg[..., None], rgba, rgba_bg)
    out = {}
    if disturbance is not None:
        cid = fid2cid_padded[ids]
        out["cid"] = cid[..., None].flip(1)
        rgba = disturb(rgba, rgba_bg, cid, int(fid2cid_padded.max()) + 1,
                       disturbance["w_fg"], disturbance["w_bg"], disturbance["u_rand"])
    pos = verts_clip
    if align_boundary_except_vid is not None:
        pos = verts_clip.clone()
        vid = torch.as_tensor(align_boundary_except_vid).long()
        pos[:, vid] = verts_clip[:, vid].detach()
    rgba_pre = rgba
    rgba_aa = antialias(rgba, rast, pos, faces, adj_opp)
    out.update(albedo=albedo.flip(1), normal=normal.flip(1), diffuse=diffuse.flip(1),
               diffuse_detach_normal=diffuse_detach_normal.flip(1), rgba=rgba_aa.flip(1), rgba_pre=rgba_pre.flip(1))
    return out


def render_rgba_vis(rast, rast_db, verts, verts_clip, faces, adj_opp, background=(1.0, 1.0, 1.0), verts_uv=None, faces_uv=None, tex=None,
                    lights=None):
    """NVDiffRenderer.render_rgba_vis (render_nvdiffrast.py:486-567) for shade_smooth=True, lighting_space='world': the visualisation
    render of the viewer / editor / NeRF export (flame_viewer.py:327, export_as_nerf_dataset.py:436).  Differences from render_rgba: no
    disturbance and no detaches, `normal` and `diffuse` carry the background outside the mesh (:553-554), the albedo is 1 without a
    texture (:531), and `lights=None` means lighting_type 'constant' (diffuse = 1, :332-333).  The per-vertex colour branch (:526-530) is
    not restated (the engine has no such input).  Returns [B,H,W,C] planes in image orientation."""
    B, H, W, _ = rast.shape
    dt = rast.dtype
    ids = rast[..., 3].long()
    fg = (ids > 0)[..., None]
    normal, _ = interpolate(compute_v_normals(verts, faces), rast, faces)
    normal = safe_normalize(normal)
    if verts_uv is not None and faces_uv is not None and tex is not None:
        texc, texd = interpolate(verts_uv[None], rast, faces_uv, rast_db)
        mips = build_mips(tex.permute(1, 2, 0))
        bi, yi, xi = torch.nonzero(fg[..., 0], as_tuple=True)
        alb_bg = _bilinear_wrap(mips[0], torch.zeros(1, dtype=dt, device=rast.device), torch.zeros(1, dtype=dt, device=rast.device))[0]
        albedo = alb_bg.expand(B, H, W, 3).clone().index_put((bi, yi, xi), texture_sample(mips, texc[bi, yi, xi], texd[bi, yi, xi]))
    else:
        albedo = torch.ones_like(normal)
    diffuse = sh_shading(normal, lights) if lights is not None else torch.ones_like(normal)
    rgba = torch.cat([albedo * diffuse, fg.to(dt)], -1)
    if isinstance(background, (list, tuple)):
        rgba_bg = torch.tensor(list(background) + [0], dtype=dt, device=rast.device).expand(B, H, W, 4)
    elif isinstance(background, torch.Tensor):
        rgba_bg = torch.cat([background.to(dt), torch.zeros_like(background[..., :1], dtype=dt)], -1)
    else:
        raise ValueError(f"Unknown background type: {type(background)}")
    rgba_bg = rgba_bg.flip(1)
    normal = torch.where(fg, normal, rgba_bg[..., :3])
    diffuse = torch.where(fg, diffuse, rgba_bg[..., :3])
    rgba = torch.where(fg, rgba, rgba_bg)
    rgba_aa = antialias(rgba, rast, verts_clip, faces, adj_opp)
    return dict(albedo=albedo.flip(1), normal=normal.flip(1), diffuse=diffuse.flip(1), rgba=rgba_aa.flip(1), verts_clip=verts_clip)
