// Per-pixel bodies of the three fused render passes (shared by render.cu kernels and tests/hostcheck):
//   pass A  shade:           rasterised id -> composite colour of foreground pixels + reg_diffuse partials
//   pass B  disturb+AA+loss: cluster colour disturbance, antialias (gather form), L1 photometric partial sums
//   pass C  backward:        analytic adjoint of B and A -> clip-position / vertex-normal / texel / light gradients
// Reference graph being restated: NVDiffRenderer.render_rgba (vhap/util/render_nvdiffrast.py:354-484) and
// compute_photometric_energy (vhap/model/tracker.py:391-478); oracle: oracle/render.py, oracle/energy.py.
#pragma once
#include "pixel_math.cuh"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

struct PassArgs {
  RenderArgs R;
  const uint16_t* target;     // [B,H,W,4] fp16 bits, IMAGE orientation (row 0 = top); or, target_u8 != 0: [B,H,W,3] uint8 RGB as decoded
  int target_u8;              // vhap_frame_batch::target_format
  f4* pre;                    // [B,H,W] composite colour of fg pixels (rgb,1) in raster orientation
  uint8_t* signs;             // [B,H,W] 2 bits per channel: 0 zero, 1 positive, 2 negative  (sign of gt - pred)
  float* final_rgba;          // optional [B,H,W,4] debug plane (after AA), raster orientation, or NULL
  // cluster pools (render_nvdiffrast.py:445-459)
  const int* pool_list;       // pixel indices grouped by cluster, ascending inside a cluster
  const int* pool_base;       // [16]
  const int* pool_count;      // [16]
  const int* pool_tri;        // rasterised id of pool_list[i] (saves the dependent tri_id[pixel] access in the list-walking passes)
  int disturb;                // enable_disturbance
  float rate_fg, rate_bg;     // <0 = None
  const uint8_t* inj_w;       // injected Bernoulli draws (bit0 fg, bit1 bg) or NULL -> Philox
  const float* inj_u;         // injected uniform for the pool index or NULL -> Philox
  uint64_t seed, step;
  const int* step_ptr;        // optional device-resident step counter ([1] = global step) used instead of `step` (CUDA-graph replay)
  int bg_mode; float bg_color[3];
  // scalars produced by the finalize step, consumed by pass C
  const float* scal;          // [0] photo_scale = w_photo/(3 n_fg)  [1] g_var = w_regdiff/(B_glob*H*W)  [2] g_max (w_regdiff if max>1 else 0)
                              // [3] argmax pixel*3+channel as float bits (int)  [4] fg flag of argmax
  // gradient sinks
  float* g_clip;              // [B,V,4]
  float* g_vnorm;             // [B,V,4]
  float* g_tex;               // gradient pyramid (float4 layout) or NULL
  // optional debug planes (raster orientation), NULL to skip
  f4* plane_albedo; f4* plane_normal; f4* plane_diffuse;
  // antialias pair cache, one float per (pixel 0 of the pair, direction d): 0 = no blend, else sign = near surface is pixel 0's,
  // |code| - 1 = alpha (written by the pair-analysis pass for EVERY pair of adjacent pixels with different ids)
  float* aa_code;             // [B,H,W,2]
  const uint8_t* loss_mask;   // optional test hook [B,H,W], IMAGE orientation: 0 = the pixel is left out of the L1 photometric sum (vhap_set_loss_mask)
};

// ------------------------------------------------------------------------------------------ small utilities
VH_HD float half_bits_to_float(uint16_t h) {
#if defined(__CUDA_ARCH__)
  return __half2float(__ushort_as_half(h));
#else
  uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023;
  float v;
  if (e == 0) v = ldexpf((float)m, -24);
  else if (e == 31) v = m ? NAN : INFINITY;
  else v = ldexpf((float)(m + 1024), (int)e - 25);
  return s ? -v : v;
#endif
}

VH_HD void philox4x32(uint64_t key, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// target pixel (image-orientation pixel index) as float RGB.  uint8 targets are what the reference's dataset decodes; F.to_tensor divides
// by 255 in fp32 (video_dataset.py:256-260) -- the same IEEE division happens here, so no host-side float image is ever materialised
// b / 255 correctly rounded (== the IEEE fp32 division of F.to_tensor for all 256 byte values: tests/test_gpu_staging.py) with one
// multiplication and one Newton correction in FMAs instead of the ~20-instruction division sequence, three times per pixel
VH_HD float u8_unit(uint8_t b) {
#if defined(__CUDA_ARCH__)
  const float r = 1.f / 255.f, x = (float)b;
  const float q = x * r;
  return fmaf(fmaf(-q, 255.f, x), r, q);
#else
  return (float)b / 255.f;
#endif
}
VH_HD f3 load_target(const PassArgs& P, size_t ipix) {
  if (P.target_u8) {
    const uint8_t* t = (const uint8_t*)P.target + ipix * 3;
    return mk3(u8_unit(t[0]), u8_unit(t[1]), u8_unit(t[2]));
  }
  const uint16_t* t = P.target + ipix * 4;
  return mk3(half_bits_to_float(t[0]), half_bits_to_float(t[1]), half_bits_to_float(t[2]));
}

// composite colour before disturbance: fg -> pass-A buffer, bg -> target image (flipped, render_nvdiffrast.py:419) or constant
VH_HD f4 pre_color(const PassArgs& P, int b, int y, int x, int id) {
  const RenderArgs& A = P.R;
  size_t pix = ((size_t)b * A.H + y) * A.W + x;
  if (id > 0) { f4 c = P.pre[pix]; c.w = 1.f; return c; }      // .w of the buffer holds z/w, the colour's alpha is 1
  f4 c; c.w = 0.f;
  if (P.bg_mode == 0) {
    f3 t = load_target(P, ((size_t)b * A.H + (A.H - 1 - y)) * A.W + x);
    c.x = t.x; c.y = t.y; c.z = t.z;
  } else { c.x = P.bg_color[0]; c.y = P.bg_color[1]; c.z = P.bg_color[2]; }
  return c;
}

// Disturbed colour D(q) (render_nvdiffrast.py:445-459).  own_weight = d D(q) / d pre(q) (0 when the pixel took a pool sample).
VH_HD f4 disturbed_color(const PassArgs& P, int b, int y, int x, int id, float* own_weight) {
  const RenderArgs& A = P.R;
  f4 c = pre_color(P, b, y, x, id);
  if (own_weight) *own_weight = 1.f;
  if (!P.disturb) return c;
  int cid = A.fid2cid[id];
  if (cid == 1) return c;
  size_t pix = ((size_t)b * A.H + y) * A.W + x;
  bool w; float u;
  float rate = cid == 0 ? P.rate_bg : P.rate_fg;
  if (P.inj_w) {
    w = cid == 0 ? ((P.inj_w[pix] >> 1) & 1) : (P.inj_w[pix] & 1);
    u = P.inj_u[pix];
  } else {
    // counter-based generator (splitmix64 of (pixel, step, seed)); the reference's torch Philox stream cannot be reproduced
    // anyway (different consumption order), see DESIGN.md "Disturbance randomness"
    uint64_t stp = P.step_ptr ? (uint64_t)P.step_ptr[1] : P.step;
    uint64_t z = (uint64_t)pix * 0x9E3779B97F4A7C15ull + stp * 0xD1B54A32D192ED03ull + P.seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    float ub = (float)((uint32_t)z >> 8) * 5.9604644775390625e-08f;      // 24-bit uniform in [0,1)
    w = rate >= 0.f && ub < rate;
    u = (float)((uint32_t)(z >> 40)) * 5.9604644775390625e-08f;          // independent 24 bits
  }
  if (rate < 0.f) w = false;
  int n = P.pool_count[cid];
  if (!w || n <= 0) return c;
  int idx = (int)(u * (float)n);
  if (idx > n - 1) idx = n - 1;
  int q = P.pool_list[P.pool_base[cid] + idx];
  int qx, qy, qb; vh_unflatten(A, q, qb, qy, qx);
  if (own_weight) *own_weight = 0.f;
  return pre_color(P, qb, qy, qx, cid > 0 ? 1 : 0);     // a pool of cluster cid >= 1 holds foreground pixels only, cluster 0 background only
}

// ------------------------------------------------------------------------------------------ antialias pair analysis
struct AAPair {
  bool found;
  bool near0;          // near surface is pixel 0's
  int tri, va, vb;     // chosen triangle, edge vertex ids
  float alpha;         // t - 0.5
  int d;
  // for the backward
  float ax, ay, bx, by, pitch;   // edge endpoints in the pair frame (axis coordinate, perpendicular), relative to the near centre
  f4 pa, pb;
};

// (x0,y0) is pixel 0; pixel 1 = (x0+1,y0) for d=0, (x0,y0+1) for d=1.  ids are triangle+1 (0 = empty), id0 != id1.
VH_HD void aa_analyze(const RenderArgs& A, int b, int x0, int y0, int d, int id0, int id1, AAPair& r) {
  r.found = false; r.d = d;
  int t0 = id0 - 1, t1 = id1 - 1;
  int x1 = x0 + (d == 0), y1 = y0 + (d == 1);
  bool use0;
  if (t0 >= 0 && t1 >= 0) {
    if (A.zwbuf) use0 = A.zwbuf[(((size_t)b * A.H + y0) * A.W + x0) * 4 + 3] < A.zwbuf[(((size_t)b * A.H + y1) * A.W + x1) * 4 + 3];
    else use0 = tri_zw(A, b, x0, y0, t0) < tri_zw(A, b, x1, y1, t1);
  }
  else use0 = t0 >= 0;
  r.near0 = use0;
  int tri = use0 ? t0 : t1;
  r.tri = tri;
  int nx = use0 ? x0 : x1, ny = use0 ? y0 : y1;
  float sgn = use0 ? 1.f : -1.f;
  float fx = (float)(2 * nx + 1) / A.W - 1.f, fy = (float)(2 * ny + 1) / A.H - 1.f;
  i4 f = A.faces[tri];
  int vi[3] = {f.x, f.y, f.z};
  const f4* cl = A.clip + (size_t)b * A.V;
  const float* nd = A.ndc ? A.ndc + (size_t)b * A.V * 2 : nullptr;
  float sx[3], sy[3];
  for (int k = 0; k < 3; ++k) {
    float qx, qy;
    if (nd) { qx = nd[vi[k] * 2] - fx; qy = nd[vi[k] * 2 + 1] - fy; }
    else { f4 pk = cl[vi[k]]; qx = pk.x / pk.w - fx; qy = pk.y / pk.w - fy; }
    sx[k] = d == 0 ? qx : qy; sy[k] = d == 0 ? qy : qx;
  }
  float pitch = (d == 0 ? 2.f / A.W : 2.f / A.H) * sgn;
  const int* adj = A.adj_opp + (size_t)tri * 4;
  for (int k = 0; k < 3; ++k) {
    int ia = k, ib = (k + 1) % 3, ic = (k + 2) % 3;
    float ax = sx[ia], ay = sy[ia], bx = sx[ib], by = sy[ib];
    bool cross = (ay > 0.f) != (by > 0.f);
    if (!cross) continue;
    float xc = (ax * by - ay * bx) / (by - ay);
    float t = xc / pitch;
    if (!(t >= 0.f && t <= 1.f)) continue;
    int op = adj[k];
    bool sil = op == -1;
    if (op >= 0) {
      float qx, qy;
      if (nd) { qx = nd[op * 2] - fx; qy = nd[op * 2 + 1] - fy; }
      else { f4 po = cl[op]; qx = po.x / po.w - fx; qy = po.y / po.w - fy; }
      float ox = d == 0 ? qx : qy, oy = d == 0 ? qy : qx;
      float ex = bx - ax, ey = by - ay;
      float side_c = ex * (sy[ic] - ay) - ey * (sx[ic] - ax);
      float side_o = ex * (oy - ay) - ey * (ox - ax);
      sil = side_c * side_o > 0.f;
    }
    if (!sil) continue;
    r.found = true; r.alpha = t - 0.5f; r.va = vi[ia]; r.vb = vi[ib];
    r.ax = ax; r.ay = ay; r.bx = bx; r.by = by; r.pitch = pitch; r.pa = cl[vi[ia]]; r.pb = cl[vi[ib]];
    return;
  }
}

// gradient of alpha wrt the clip positions of the two edge vertices
VH_HD void aa_bwd(const PassArgs& P, int b, const AAPair& r, float g_alpha) {
  const RenderArgs& A = P.R;
  float den = r.by - r.ay, num = r.ax * r.by - r.ay * r.bx;
  float g_xc = g_alpha / r.pitch;
  float g_num = g_xc / den, g_den = -g_xc * num / (den * den);
  float g_ax = g_num * r.by, g_by = g_num * r.ax + g_den, g_ay = -g_num * r.bx - g_den, g_bx = -g_num * r.ay;
  const int vid[2] = {r.va, r.vb};
  const f4 pp[2] = {r.pa, r.pb};
  const float gsx[2] = {g_ax, g_bx}, gsy[2] = {g_ay, g_by};
  for (int k = 0; k < 2; ++k) {
    if (A.vert_flags && (A.vert_flags[vid[k]] & 1)) continue;       // detach_by_indices (render_nvdiffrast.py:349-352)
    float g_qx = r.d == 0 ? gsx[k] : gsy[k], g_qy = r.d == 0 ? gsy[k] : gsx[k];
    float iw = 1.f / pp[k].w;
    float* t = P.g_clip + ((size_t)b * A.V + vid[k]) * 4;
    VH_ATOMIC_ADD4(t, g_qx * iw, g_qy * iw, 0.f, -(g_qx * pp[k].x + g_qy * pp[k].y) * iw * iw);
  }
}

// ------------------------------------------------------------------------------------------ pass A
// acc[0] += var_c(diffuse) over fg pixels, acc[1] += 1 per fg pixel; (mx, mx_idx) running max of diffuse (fg only)
VH_HD void passA_body(const PassArgs& P, int b, int y, int x, float* acc, float& mx, int& mx_idx, int id_known = -1) {
  const RenderArgs& A = P.R;
  size_t pix = ((size_t)b * A.H + y) * A.W + x;
  int id = id_known >= 0 ? id_known : A.tri_id[pix];
  if (id <= 0) {
    if (P.plane_albedo) { f4 z = {0, 0, 0, 0}; P.plane_albedo[pix] = z; P.plane_normal[pix] = z; P.plane_diffuse[pix] = z; }
    return;
  }
  PixShade s;
  shade_pixel(A, b, x, y, id - 1, s);
  f4 c; c.x = s.rgb.x; c.y = s.rgb.y; c.z = s.rgb.z; c.w = s.ts.zw;      // .w: z/w for the antialias depth comparison
  P.pre[pix] = c;
  float m = (s.diffuse.x + s.diffuse.y + s.diffuse.z) * (1.f / 3.f);
  float dx = s.diffuse.x - m, dy = s.diffuse.y - m, dz = s.diffuse.z - m;
  acc[0] += 0.5f * (dx * dx + dy * dy + dz * dz);
  acc[1] += 1.f;
  float dm = s.diffuse.x; int ch = 0;
  if (s.diffuse.y > dm) { dm = s.diffuse.y; ch = 1; }
  if (s.diffuse.z > dm) { dm = s.diffuse.z; ch = 2; }
  if (dm > mx) { mx = dm; mx_idx = (int)pix * 3 + ch; }
  if (P.plane_albedo) {
    f4 a = {s.albedo.x, s.albedo.y, s.albedo.z, 0}; P.plane_albedo[pix] = a;
    f4 n = {s.n.x, s.n.y, s.n.z, 0}; P.plane_normal[pix] = n;
    f4 dd = {s.diffuse.x, s.diffuse.y, s.diffuse.z, 0}; P.plane_diffuse[pix] = dd;
  }
}

// ------------------------------------------------------------------------------------------ pair analysis (pass B1)
// one call per pair of adjacent pixels with different ids; pix0 = (b,y0,x0), d = 0 horizontal / 1 vertical
VH_HD void aa_pair_body(const PassArgs& P, int b, int y0, int x0, int d) {
  const RenderArgs& A = P.R;
  size_t p0 = ((size_t)b * A.H + y0) * A.W + x0, p1 = p0 + (d == 0 ? 1 : A.W);
  AAPair r;
  aa_analyze(A, b, x0, y0, d, A.tri_id[p0], A.tri_id[p1], r);
  P.aa_code[p0 * 2 + d] = r.found ? (r.alpha + 1.f) * (r.near0 ? 1.f : -1.f) : 0.f;
}
// cached result of the pair (p, q) seen from pixel p; returns false if the pair produces no blend
VH_HD bool aa_lookup(const PassArgs& P, size_t pix, size_t qpix, int d, bool p_is0, bool& modifies_p, float& a, float& alpha) {
  float code = P.aa_code[(p_is0 ? pix : qpix) * 2 + d];
  if (code == 0.f) return false;
  bool near0 = code > 0.f;
  alpha = fabsf(code) - 1.f;
  bool p_near = (near0 == p_is0);
  modifies_p = (alpha > 0.f) ? !p_near : p_near;      // alpha > 0 modifies the FAR pixel, otherwise the NEAR pixel
  a = fabsf(alpha);
  return true;
}

// ------------------------------------------------------------------------------------------ pass B
// Gather-form antialias: the final colour of p is D(p) plus the blends of the (up to 4) pixel pairs p belongs to for
// which p is the pixel being modified.  acc[0] += sum_c |gt - pred|, acc[1] += (alpha_aa > 0).
VH_HD void passB_body(const PassArgs& P, int b, int y, int x, float* acc) {
  const RenderArgs& A = P.R;
  size_t pix = ((size_t)b * A.H + y) * A.W + x;
  int id = A.tri_id[pix];
  f4 Dp = disturbed_color(P, b, y, x, id, nullptr);
  f4 out = Dp;
  const int nbx[4] = {x - 1, x + 1, x, x}, nby[4] = {y, y, y - 1, y + 1};
  for (int k = 0; k < 4; ++k) {
    int qx = nbx[k], qy = nby[k];
    if (qx < 0 || qx >= A.W || qy < 0 || qy >= A.H) continue;
    int idq = A.tri_id[((size_t)b * A.H + qy) * A.W + qx];
    if (idq == id) continue;
    int d = k >> 1;
    bool p_is0 = (k & 1);                       // k=1: p left of q; k=3: p below q
    bool modifies_p; float a, alpha;
    if (!aa_lookup(P, pix, ((size_t)b * A.H + qy) * A.W + qx, d, p_is0, modifies_p, a, alpha)) continue;
    if (!modifies_p) continue;
    f4 Dq = disturbed_color(P, b, qy, qx, idq, nullptr);
    out.x += a * (Dq.x - Dp.x); out.y += a * (Dq.y - Dp.y); out.z += a * (Dq.z - Dp.z); out.w += a * (Dq.w - Dp.w);
  }
  const f3 t = load_target(P, ((size_t)b * A.H + (A.H - 1 - y)) * A.W + x);
  float e0 = t.x - out.x, e1 = t.y - out.y, e2 = t.z - out.z;
  if (P.loss_mask && !P.loss_mask[((size_t)b * A.H + (A.H - 1 - y)) * A.W + x]) { e0 = 0.f; e1 = 0.f; e2 = 0.f; }
  acc[0] += fabsf(e0) + fabsf(e1) + fabsf(e2);
  if (out.w > 0.f) acc[1] += 1.f;
  uint8_t sg = (uint8_t)((e0 > 0.f ? 1 : (e0 < 0.f ? 2 : 0)) | ((e1 > 0.f ? 1 : (e1 < 0.f ? 2 : 0)) << 2) | ((e2 > 0.f ? 1 : (e2 < 0.f ? 2 : 0)) << 4));
  P.signs[pix] = sg;
  if (P.final_rgba) { float* o = P.final_rgba + pix * 4; o[0] = out.x; o[1] = out.y; o[2] = out.z; o[3] = out.w; }
  // 'aa' plane of the reference's render_out (render_nvdiffrast.py:466: pixels the antialias changed), kept in the free .w of the albedo plane
  if (P.plane_albedo) P.plane_albedo[pix].w = (out.x != Dp.x || out.y != Dp.y || out.z != Dp.z || out.w != Dp.w) ? 1.f : 0.f;
}

VH_HD f3 sign_grad(uint8_t sg, float scale) {
  // d loss / d pred = -sign(gt - pred) * scale
  f3 g;
  int s0 = sg & 3, s1 = (sg >> 2) & 3, s2 = (sg >> 4) & 3;
  g.x = s0 == 1 ? -scale : (s0 == 2 ? scale : 0.f);
  g.y = s1 == 1 ? -scale : (s1 == 2 ? scale : 0.f);
  g.z = s2 == 1 ? -scale : (s2 == 2 ? scale : 0.f);
  return g;
}

// ------------------------------------------------------------------------------------------ pass C
// g_lights_local[27] accumulates the light gradient of this pixel.  ext_grad (optional, [B,H,W,4] raster orientation)
// replaces the L1-loss gradient with a caller-provided d L / d rgba_aa (modular render_rgba backward).
// Split in two so that each kernel stays small (registers, instruction cache): C1 = colour-gradient bookkeeping through the
// antialias / disturbance adjoints (+ the rare silhouette position gradients) -> d L / d rgb of the pixel; C2 = shading adjoint.
VH_HD f3 passC1_body(const PassArgs& P, int b, int y, int x, const float* ext_grad, int id_known = -1) {
  const RenderArgs& A = P.R;
  size_t pix = ((size_t)b * A.H + y) * A.W + x;
  int id = id_known >= 0 ? id_known : A.tri_id[pix];
  // Only FOREGROUND pixels are visited (the kernel walks the compacted foreground list): a background pixel has no
  // parameters behind its colour, and the position gradient of a (foreground, background) pair is owned by its
  // foreground pixel; (foreground, foreground) pairs are owned by their pixel 0.
  if (id <= 0) return mk3(0, 0, 0);
  float scale = P.scal[0];
  f3 gp = ext_grad ? mk3(ext_grad[pix * 4], ext_grad[pix * 4 + 1], ext_grad[pix * 4 + 2]) : sign_grad(P.signs[pix], scale);
  float own_w;
  f4 Dp = disturbed_color(P, b, y, x, id, &own_w);
  float self_w = 1.f;                 // weight of D(p) in final(p)
  f3 gD = mk3(0, 0, 0);               // d L / d D(p)
  const int nbx[4] = {x - 1, x + 1, x, x}, nby[4] = {y, y, y - 1, y + 1};
  for (int k = 0; k < 4; ++k) {
    int qx = nbx[k], qy = nby[k];
    if (qx < 0 || qx >= A.W || qy < 0 || qy >= A.H) continue;
    size_t qpix = ((size_t)b * A.H + qy) * A.W + qx;
    int idq = A.tri_id[qpix];
    if (idq == id) continue;
    int d = k >> 1;
    bool p_is0 = (k & 1);
    bool modifies_p; float a, alpha;
    if (!aa_lookup(P, pix, qpix, d, p_is0, modifies_p, a, alpha)) continue;
    f3 gq = ext_grad ? mk3(ext_grad[qpix * 4], ext_grad[qpix * 4 + 1], ext_grad[qpix * 4 + 2]) : sign_grad(P.signs[qpix], scale);
    if (modifies_p) self_w -= a;                  // final(p) = D(p) + a (D(q) - D(p))
    else { gD.x += a * gq.x; gD.y += a * gq.y; gD.z += a * gq.z; }   // final(q) = D(q) + a (D(p) - D(q))
    // position gradient: handled once per pair (see the ownership rule above)
    if ((p_is0 || idq <= 0) && P.g_clip) {
      f4 Dq = disturbed_color(P, b, qy, qx, idq, nullptr);
      // target of the blend and the "other" colour
      f3 g_t = modifies_p ? gp : gq;
      f3 diff = modifies_p ? mk3(Dq.x - Dp.x, Dq.y - Dp.y, Dq.z - Dp.z) : mk3(Dp.x - Dq.x, Dp.y - Dq.y, Dp.z - Dq.z);
      float g_abs = dot3(g_t, diff);              // d L / d |alpha|
      if (ext_grad) {                             // alpha channel also carries gradient in the modular path
        float ga_t = modifies_p ? ext_grad[pix * 4 + 3] : ext_grad[qpix * 4 + 3];
        g_abs += ga_t * (modifies_p ? (Dq.w - Dp.w) : (Dp.w - Dq.w));
      }
      float g_alpha = alpha > 0.f ? g_abs : -g_abs;
      AAPair r;                                   // geometry of the (rare) blending pair, recomputed for the adjoint
      if (p_is0) aa_analyze(A, b, x, y, d, id, idq, r);
      else aa_analyze(A, b, qx, qy, d, idq, id, r);
      if (r.found) aa_bwd(P, b, r, g_alpha);
    }
  }
  gD.x += self_w * gp.x; gD.y += self_w * gp.y; gD.z += self_w * gp.z;
  return gD * own_w;
}

VH_HD void passC2_body(const PassArgs& P, int b, int y, int x, f3 g_rgb, float* g_lights_local, int id_known = -1, VertGrad* vg = nullptr) {
  const RenderArgs& A = P.R;
  size_t pix = ((size_t)b * A.H + y) * A.W + x;
  int id = id_known >= 0 ? id_known : A.tri_id[pix];
  if (id <= 0) return;
  // reg_diffuse on diffuse_detach_normal (tracker.py:547-550): variance term + global max term
  f3 g_dd = mk3(0, 0, 0);
  PixShade s;
  shade_pixel(A, b, x, y, id - 1, s);
  float g_var = P.scal[1];
  if (g_var != 0.f) {
    float m = (s.diffuse.x + s.diffuse.y + s.diffuse.z) * (1.f / 3.f);
    g_dd = mk3((s.diffuse.x - m) * g_var, (s.diffuse.y - m) * g_var, (s.diffuse.z - m) * g_var);
    float g_max = P.scal[2];
    if (g_max != 0.f) {
      int am = ((const int*)P.scal)[3];
      if (am / 3 == (int)pix) { int ch = am % 3; if (ch == 0) g_dd.x += g_max; else if (ch == 1) g_dd.y += g_max; else g_dd.z += g_max; }
    }
  }
  shade_pixel_bwd(A, b, id - 1, s, g_rgb, g_dd, P.g_clip, P.g_vnorm, P.g_tex, g_lights_local, vg);
}

VH_HD void passC_body(const PassArgs& P, int b, int y, int x, const float* ext_grad, float* g_lights_local) {
  f3 g = passC1_body(P, b, y, x, ext_grad);
  passC2_body(P, b, y, x, g, g_lights_local);
}
