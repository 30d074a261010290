"""Oracle rasteriser.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against nvdiffrast (see oracle/__init__.py).

Replaces `dr.rasterize` at vhap/util/render_nvdiffrast.py:254 (called from :216-245).  Two parts:

1. `rasterize_ids`  -- visibility, in exact integer / IEEE-fp32 arithmetic.  This is THE specification the
   CUDA rasteriser (vhap_b200/csrc/raster.cu) reproduces bit for bit (DESIGN.md "Rasteriser specification"):
     * vertex valid iff x,y,z,w finite and w > 0 (no near-plane polygon clipping: a triangle with an
       invalid vertex is culled; per-pixel depth outside [-1,1] is rejected instead);
     * snap: X = rint(clamp((x/w) * (W*8), +-2^17)), Y likewise with H  (fp32 div, fp32 mul, round-half-even)
       = 1/16-pixel fixed point relative to the viewport centre (CudaRaster's CR_SUBPIXEL_LOG2 = 4);
     * pixel (px,py), row 0 = BOTTOM, samples C = ((2px+1-W)*8, (2py+1-H)*8);
     * coverage: three int64 edge functions, sign-normalised by the triangle's orientation, sample on an
       edge belongs to the triangle iff the (normalised) edge direction has dy<0 or (dy==0 and dx<0);
       zero-area triangles are culled; optional back-face culling (area2 < 0) for the fork's option;
     * depth: fp32 plane z = (zx*Cx + zy*Cy) + zc through the three z/w values (formulas below, every op
       an individually rounded fp32 op, no FMA); nearest wins, ties go to the LOWER triangle index.
2. `shade_pass` -- perspective-correct barycentrics (u,v), z/w and the screen-space derivative block
   (du/dx,du/dy,dv/dx,dv/dy) from the UNSNAPPED clip coordinates of the winning triangle, differentiable
   torch code following nvdiffrast's published forward shader.
"""
import numpy as np
import torch

GUARD = float(1 << 17)
F32 = np.float32


def snap_vertices(clip, H, W):
    """clip [V,4] float32 -> X,Y int64 [V], zw float32 [V], valid bool [V]."""
    clip = np.ascontiguousarray(clip, dtype=F32)
    x, y, z, w = clip[:, 0], clip[:, 1], clip[:, 2], clip[:, 3]
    valid = np.isfinite(clip).all(1) & (w > 0)
    ws = np.where(valid, w, F32(1))
    with np.errstate(all="ignore"):
        sx = (x / ws) * F32(W * 8)
        sy = (y / ws) * F32(H * 8)
        zw = z / ws
    sx = np.minimum(np.maximum(sx, F32(-GUARD)), F32(GUARD))
    sy = np.minimum(np.maximum(sy, F32(-GUARD)), F32(GUARD))
    X = np.rint(sx).astype(np.int64)
    Y = np.rint(sy).astype(np.int64)
    X[~valid] = 0
    Y[~valid] = 0
    return X, Y, zw.astype(F32), valid


def _floordiv(a, b):
    return a // b


def rasterize_ids_loop(clip, tri, H, W, cull_backface=False):
    """The specification written as one numpy block per triangle (slow: a Python loop over the triangles).  `rasterize_ids` evaluates the same
    arithmetic batched over the triangles; tests/test_oracle_raster_fast.py holds the two to bit-identical ids AND depths.
    clip [B,V,4] float32 numpy, tri [F,3] int -> ids [B,H,W] int32 (triangle index + 1, 0 = empty; row 0 = bottom) and the winning
    depth [B,H,W] float32 (inf where empty)."""
    clip = np.asarray(clip, dtype=F32)
    tri = np.asarray(tri, dtype=np.int64)
    B = clip.shape[0]
    ids = np.zeros((B, H, W), np.int32)
    zbuf = np.full((B, H, W), np.inf, F32)
    for b in range(B):
        X, Y, zw, valid = snap_vertices(clip[b], H, W)
        tv = valid[tri].all(1)
        X3, Y3, Z3 = X[tri], Y[tri], zw[tri]
        d1x, d1y = X3[:, 1] - X3[:, 0], Y3[:, 1] - Y3[:, 0]
        d2x, d2y = X3[:, 2] - X3[:, 0], Y3[:, 2] - Y3[:, 0]
        area2 = d1x * d2y - d2x * d1y
        keep = tv & (area2 != 0)
        if cull_backface:
            keep &= area2 > 0
        mnx, mxx = X3.min(1), X3.max(1)
        mny, mxy = Y3.min(1), Y3.max(1)
        px0 = np.maximum(-_floordiv(-(mnx + W * 8 - 8), 16), 0)
        px1 = np.minimum(_floordiv(mxx + W * 8 - 8, 16), W - 1)
        py0 = np.maximum(-_floordiv(-(mny + H * 8 - 8), 16), 0)
        py1 = np.minimum(_floordiv(mxy + H * 8 - 8, 16), H - 1)
        keep &= (px0 <= px1) & (py0 <= py1)
        idb, zb = ids[b], zbuf[b]
        for t in np.nonzero(keep)[0]:
            x0, x1, y0, y1 = int(px0[t]), int(px1[t]), int(py0[t]), int(py1[t])
            Cx = ((2 * np.arange(x0, x1 + 1, dtype=np.int64) + 1 - W) * 8)[None, :]
            Cy = ((2 * np.arange(y0, y1 + 1, dtype=np.int64) + 1 - H) * 8)[:, None]
            s = 1 if area2[t] > 0 else -1
            inside = None
            for a, c in ((1, 2), (2, 0), (0, 1)):
                ax, ay, bx, by = X3[t, a], Y3[t, a], X3[t, c], Y3[t, c]
                dx, dy = s * (bx - ax), s * (by - ay)
                E = dx * (Cy - ay) - dy * (Cx - ax)
                tl = (dy < 0) or (dy == 0 and dx < 0)
                ins = (E >= 0) if tl else (E > 0)
                inside = ins if inside is None else (inside & ins)
            if not inside.any():
                continue
            # fp32 depth plane (each op rounded individually)
            z0, z1, z2 = Z3[t, 0], Z3[t, 1], Z3[t, 2]
            f1x, f1y, f2x, f2y = F32(d1x[t]), F32(d1y[t]), F32(d2x[t]), F32(d2y[t])
            dz1, dz2 = F32(z1 - z0), F32(z2 - z0)
            af = F32(area2[t])
            with np.errstate(all="ignore"):
                zx = F32(F32(F32(dz1 * f2y) - F32(dz2 * f1y)) / af)
                zy = F32(F32(F32(dz2 * f1x) - F32(dz1 * f2x)) / af)
                zc = F32(F32(z0 - F32(zx * F32(X3[t, 0]))) - F32(zy * F32(Y3[t, 0])))
                zp = ((zx * Cx.astype(F32)) + (zy * Cy.astype(F32))) + zc
            assert zp.dtype == F32
            ok = inside & (zp >= F32(-1)) & (zp <= F32(1))
            zs = zb[y0:y1 + 1, x0:x1 + 1]
            isub = idb[y0:y1 + 1, x0:x1 + 1]
            win = ok & ((zp < zs) | ((zp == zs) & (t + 1 < isub)))
            zs[win] = zp[win]
            isub[win] = t + 1
    return ids, zbuf



_BOXES = (4, 8, 16, 32)     # triangles are evaluated batched on the smallest K x K pixel grid their bounding box fits; larger ones go through the per-triangle block


def _order_key(z):
    """fp32 depth -> uint32 whose unsigned order is the float order (z in [-1, 1], finite)"""
    u = z.view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000))


def rasterize_ids(clip, tri, H, W, cull_backface=False):
    """clip [B,V,4] float32 numpy, tri [F,3] int -> ids [B,H,W] int32 (triangle index + 1, 0 = empty;
    row 0 = bottom) and the winning depth [B,H,W] float32 (inf where empty).

    Same arithmetic as `rasterize_ids_loop` (the specification above), batched: all triangles with a pixel bounding box of at most
    K x K (K in _BOXES) are tested on a [n, K, K] grid at once (int64 edge functions, individually rounded fp32 depth plane), the
    candidates are resolved per pixel by the minimum of the packed key (order-preserving depth bits << 32 | triangle id) = nearest
    depth, ties to the lower triangle index; the few larger triangles are merged with the per-triangle block afterwards."""
    clip = np.asarray(clip, dtype=F32)
    tri = np.asarray(tri, dtype=np.int64)
    B = clip.shape[0]
    ids = np.zeros((B, H, W), np.int32)
    zbuf = np.full((B, H, W), np.inf, F32)
    for b in range(B):
        X, Y, zw, valid = snap_vertices(clip[b], H, W)
        tv = valid[tri].all(1)
        X3, Y3, Z3 = X[tri], Y[tri], zw[tri]
        d1x, d1y = X3[:, 1] - X3[:, 0], Y3[:, 1] - Y3[:, 0]
        d2x, d2y = X3[:, 2] - X3[:, 0], Y3[:, 2] - Y3[:, 0]
        area2 = d1x * d2y - d2x * d1y
        keep = tv & (area2 != 0)
        if cull_backface:
            keep &= area2 > 0
        mnx, mxx = X3.min(1), X3.max(1)
        mny, mxy = Y3.min(1), Y3.max(1)
        px0 = np.maximum(-_floordiv(-(mnx + W * 8 - 8), 16), 0)
        px1 = np.minimum(_floordiv(mxx + W * 8 - 8, 16), W - 1)
        py0 = np.maximum(-_floordiv(-(mny + H * 8 - 8), 16), 0)
        py1 = np.minimum(_floordiv(mxy + H * 8 - 8, 16), H - 1)
        keep &= (px0 <= px1) & (py0 <= py1)
        ext = np.maximum(px1 - px0, py1 - py0)                          # bounding-box extent - 1
        small = keep & (ext < _BOXES[-1])
        key = np.full(H * W, np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
        zflat = np.full(H * W, np.inf, F32)
        work = []
        lo = -1
        for K in _BOXES:
            ts = np.nonzero(keep & (ext > lo) & (ext < K))[0]
            lo = K - 1
            CH = max(64, (1 << 18) // (K * K))
            work += [(K, ts[c0:c0 + CH]) for c0 in range(0, len(ts), CH)]
        for K, t in work:
            oy, ox = np.meshgrid(np.arange(K, dtype=np.int64), np.arange(K, dtype=np.int64), indexing="ij")
            gx = px0[t][:, None, None] + ox[None]                     # [n,K,K] pixel coordinates
            gy = py0[t][:, None, None] + oy[None]
            inb = (gx <= px1[t][:, None, None]) & (gy <= py1[t][:, None, None])
            Cx = (2 * gx + 1 - W) * 8
            Cy = (2 * gy + 1 - H) * 8
            sgn = np.where(area2[t] > 0, 1, -1).astype(np.int64)[:, None, None]
            inside = inb
            for a, c in ((1, 2), (2, 0), (0, 1)):
                ax, ay = X3[t, a][:, None, None], Y3[t, a][:, None, None]
                dx = sgn * (X3[t, c][:, None, None] - ax)
                dy = sgn * (Y3[t, c][:, None, None] - ay)
                E = dx * (Cy - ay) - dy * (Cx - ax)
                tl = (dy < 0) | ((dy == 0) & (dx < 0))
                inside = inside & np.where(tl, E >= 0, E > 0)
            if not inside.any():
                continue
            z0, z1, z2 = Z3[t, 0], Z3[t, 1], Z3[t, 2]
            f1x, f1y, f2x, f2y = d1x[t].astype(F32), d1y[t].astype(F32), d2x[t].astype(F32), d2y[t].astype(F32)
            dz1, dz2 = (z1 - z0).astype(F32), (z2 - z0).astype(F32)
            af = area2[t].astype(F32)
            with np.errstate(all="ignore"):
                zx = ((dz1 * f2y).astype(F32) - (dz2 * f1y).astype(F32)).astype(F32) / af
                zy = ((dz2 * f1x).astype(F32) - (dz1 * f2x).astype(F32)).astype(F32) / af
                zc = ((z0 - (zx * X3[t, 0].astype(F32)).astype(F32)).astype(F32) - (zy * Y3[t, 0].astype(F32)).astype(F32)).astype(F32)
                zp = ((zx[:, None, None] * Cx.astype(F32)) + (zy[:, None, None] * Cy.astype(F32))) + zc[:, None, None]
            assert zp.dtype == F32 and zx.dtype == F32 and zc.dtype == F32
            ok = inside & (zp >= F32(-1)) & (zp <= F32(1))
            n_i, y_i, x_i = np.nonzero(ok)
            if len(n_i) == 0:
                continue
            zz = zp[n_i, y_i, x_i]
            pix = gy[n_i, y_i, x_i] * W + gx[n_i, y_i, x_i]
            k = (_order_key(np.where(zz == 0, F32(0), zz)).astype(np.uint64) << np.uint64(32)) | (t[n_i] + 1).astype(np.uint64)
            np.minimum.at(key, pix, k)
        hit = key != np.uint64(0xFFFFFFFFFFFFFFFF)
        idb = (key & np.uint64(0xFFFFFFFF)).astype(np.int64)
        ub = (key >> np.uint64(32)).astype(np.uint32)
        zb_bits = np.where(ub & np.uint32(0x80000000), ub & np.uint32(0x7FFFFFFF), ~ub)
        zflat[hit] = zb_bits.view(F32)[hit]
        ids[b] = np.where(hit, idb, 0).astype(np.int32).reshape(H, W)
        zbuf[b] = zflat.reshape(H, W)
        big = keep & ~small
        if big.any():                                                  # merged in ascending index order with the same comparison
            sub = np.nonzero(big)[0]
            ids_l, z_l = rasterize_ids_loop(clip[b:b + 1], tri[sub], H, W, cull_backface)
            idl = np.where(ids_l[0] > 0, sub[np.maximum(ids_l[0] - 1, 0)] + 1, 0).astype(np.int32)
            zl = z_l[0]
            win = (idl > 0) & ((zl < zbuf[b]) | ((zl == zbuf[b]) & ((ids[b] == 0) | (idl < ids[b]))))
            ids[b][win] = idl[win]
            zbuf[b][win] = zl[win]
    return ids, zbuf


def shade_pass(clip, tri, ids):
    """Differentiable per-pixel barycentrics.  clip [B,V,4] torch, tri [F,3] long, ids [B,H,W] (0 = empty).
    Returns rast [B,H,W,4] = (u, v, z/w, id) and rast_db [B,H,W,4] = (du/dx, du/dy, dv/dx, dv/dy), zeros where empty.
    (nvdiffrast forward shader semantics; clamps give zero gradient when active.)"""
    B, H, W = ids.shape
    dt, dev = clip.dtype, clip.device
    ids_t = torch.as_tensor(ids, device=dev).long()
    fg = ids_t > 0
    bi, yi, xi = torch.nonzero(fg, as_tuple=True)
    t = ids_t[bi, yi, xi] - 1
    vi = tri[t]                                                     # [N,3]
    p = clip[bi[:, None], vi]                                       # [N,3,4]
    xs, ys = 2.0 / W, 2.0 / H
    fx = (xi.to(dt) * xs) + (1.0 / W - 1.0)
    fy = (yi.to(dt) * ys) + (1.0 / H - 1.0)
    px = p[:, :, 0] - fx[:, None] * p[:, :, 3]
    py = p[:, :, 1] - fy[:, None] * p[:, :, 3]
    a0 = px[:, 1] * py[:, 2] - py[:, 1] * px[:, 2]
    a1 = px[:, 2] * py[:, 0] - py[:, 2] * px[:, 0]
    a2 = px[:, 0] * py[:, 1] - py[:, 0] * px[:, 1]
    iw = 1.0 / (a0 + a1 + a2)
    b0, b1 = a0 * iw, a1 * iw
    z = p[:, 0, 2] * a0 + p[:, 1, 2] * a1 + p[:, 2, 2] * a2
    w = p[:, 0, 3] * a0 + p[:, 1, 3] * a1 + p[:, 2, 3] * a2
    zw = z / w
    X, Y, Wc = p[:, :, 0], p[:, :, 1], p[:, :, 3]
    da0dx = Y[:, 2] * Wc[:, 1] - Y[:, 1] * Wc[:, 2]
    da0dy = X[:, 1] * Wc[:, 2] - X[:, 2] * Wc[:, 1]
    da1dx = Y[:, 0] * Wc[:, 2] - Y[:, 2] * Wc[:, 0]
    da1dy = X[:, 2] * Wc[:, 0] - X[:, 0] * Wc[:, 2]
    da2dx = Y[:, 1] * Wc[:, 0] - Y[:, 0] * Wc[:, 1]
    da2dy = X[:, 0] * Wc[:, 1] - X[:, 1] * Wc[:, 0]
    datdx = da0dx + da1dx + da2dx
    datdy = da0dy + da1dy + da2dy
    dfxdx, dfydy = xs * iw, ys * iw
    # outputs are clamped first; the derivative block uses the clamped barycentrics
    b0c, b1c, zwc = b0.clamp(0, 1), b1.clamp(0, 1), zw.clamp(-1, 1)
    dudx = dfxdx * (b0c * datdx - da0dx)
    dudy = dfydy * (b0c * datdy - da0dy)
    dvdx = dfxdx * (b1c * datdx - da1dx)
    dvdy = dfydy * (b1c * datdy - da1dy)
    rast = torch.zeros(B, H, W, 4, dtype=dt, device=dev)
    rast_db = torch.zeros(B, H, W, 4, dtype=dt, device=dev)
    rast = rast.index_put((bi, yi, xi), torch.stack([b0c, b1c, zwc, (t + 1).to(dt)], -1))
    rast_db = rast_db.index_put((bi, yi, xi), torch.stack([dudx, dudy, dvdx, dvdy], -1))
    return rast, rast_db


def rasterize(clip, tri, image_size, cull_backface=False):
    """Drop-in for `dr.rasterize(ctx, pos, tri, resolution)`: returns (rast, rast_db)."""
    H, W = image_size
    ids, _ = rasterize_ids(clip.detach().to(torch.float32).cpu().numpy(), tri.cpu().numpy(), H, W, cull_backface)
    return shade_pass(clip, tri.long(), ids)
