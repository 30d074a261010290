// Per-pixel math of the fused renderer: perspective barycentrics + derivatives, attribute interpolation, mip-mapped
// texture sampling, SH shading -- forward and analytic backward.  Restates, for one pixel, what the reference does
// with dr.rasterize / dr.interpolate / dr.texture / shade() (vhap/util/render_nvdiffrast.py:254,384-405) and autograd.
// VH_HD: shared between the CUDA kernels (render.cu) and the host unit checks (tests/hostcheck).
#pragma once
#include "common.cuh"

#define VH_MAX_MIPS 14

struct RenderArgs {
  int B, H, W, V, F, T, max_level;
  const i4* faces;            // [F] (v0,v1,v2,_)
  const i4* faces_uv;         // [F]
  const float* verts_uv;      // [VT,2], v flipped
  const f4* clip;             // [B,V] clip-space positions
  const f4* vnorm;            // [B,V] unit vertex normals (xyz)
  const float* lights;        // [9,3]
  const f4* mips;             // texture pyramid, float4 texels (rgb_), level l at mip_off[l], size (T>>l)^2
  int mip_off[VH_MAX_MIPS];
  const int* tri_id;          // [B,H,W] triangle+1, 0 = empty; row 0 = bottom
  const uint8_t* face_flags;  // [F] bit0: texture coordinate detached (align_texture_except_fid)
  const uint8_t* vert_flags;  // [V] bit0: clip position detached inside antialias (align_boundary_except_vid)
  const uint8_t* fid2cid;     // [F+1]
  const int* adj_opp;         // [F*4] opposite vertex per edge (x,y,z), -1 boundary, -2 non-manifold
  const float* ndc;           // optional [B,V,2] = clip.xy / clip.w (saves the divisions in the antialias analysis)
  const float* zwbuf;         // optional [B,H,W,4]: .w of foreground pixels holds their z/w (written by pass A)
  int* tex_l0_flag;           // optional [regions]: raised where the backward scatters into level 0 of the texel-gradient pyramid
  const int* geo;             // optional [B]: view sharing, frame b uses the vertex normals of geometry geo[b] (clip positions stay per view)
  int pow2, wshift, hshift;   // pow2 != 0: W = 1 << wshift, H = 1 << hshift (pixel index -> (b,y,x) without integer divisions)
};

// flat pixel index (b*H + y)*W + x -> (b, y, x)
VH_HD void vh_unflatten(const RenderArgs& A, int pix, int& b, int& y, int& x) {
  if (A.pow2) { x = pix & (A.W - 1); y = (pix >> A.wshift) & (A.H - 1); b = pix >> (A.wshift + A.hshift); }
  else { x = pix % A.W; y = (pix / A.W) % A.H; b = pix / (A.W * A.H); }
}

struct TriSetup {
  int vi[3];
  f4 p[3];
  float fx, fy;
  float a0, a1, a2, iw;
  float b0u, b1u;             // unclamped barycentrics
  float b0, b1;               // clamped (outputs u,v)
  float zw;
  float da0dx, da0dy, da1dx, da1dy, da2dx, da2dy, datdx, datdy;
  float dudx, dudy, dvdx, dvdy;
};

VH_HD float vh_sat(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// nvdiffrast forward-shader semantics; oracle: oracle/raster.py shade_pass
VH_HD void tri_setup(const RenderArgs& A, int b, int px, int py, int tri, TriSetup& s) {
  i4 f = A.faces[tri];
  s.vi[0] = f.x; s.vi[1] = f.y; s.vi[2] = f.z;
  const f4* cl = A.clip + (size_t)b * A.V;
  s.p[0] = cl[f.x]; s.p[1] = cl[f.y]; s.p[2] = cl[f.z];
  float xs = 2.f / A.W, ys = 2.f / A.H;
  s.fx = px * xs + (1.f / A.W - 1.f);
  s.fy = py * ys + (1.f / A.H - 1.f);
  float p0x = s.p[0].x - s.fx * s.p[0].w, p0y = s.p[0].y - s.fy * s.p[0].w;
  float p1x = s.p[1].x - s.fx * s.p[1].w, p1y = s.p[1].y - s.fy * s.p[1].w;
  float p2x = s.p[2].x - s.fx * s.p[2].w, p2y = s.p[2].y - s.fy * s.p[2].w;
  s.a0 = p1x * p2y - p1y * p2x;
  s.a1 = p2x * p0y - p2y * p0x;
  s.a2 = p0x * p1y - p0y * p1x;
  s.iw = 1.f / (s.a0 + s.a1 + s.a2);
  s.b0u = s.a0 * s.iw; s.b1u = s.a1 * s.iw;
  s.b0 = vh_sat(s.b0u); s.b1 = vh_sat(s.b1u);
  float z = s.p[0].z * s.a0 + s.p[1].z * s.a1 + s.p[2].z * s.a2;
  float w = s.p[0].w * s.a0 + s.p[1].w * s.a1 + s.p[2].w * s.a2;
  s.zw = fminf(fmaxf(z / w, -1.f), 1.f);
  s.da0dx = s.p[2].y * s.p[1].w - s.p[1].y * s.p[2].w;
  s.da0dy = s.p[1].x * s.p[2].w - s.p[2].x * s.p[1].w;
  s.da1dx = s.p[0].y * s.p[2].w - s.p[2].y * s.p[0].w;
  s.da1dy = s.p[2].x * s.p[0].w - s.p[0].x * s.p[2].w;
  s.da2dx = s.p[1].y * s.p[0].w - s.p[0].y * s.p[1].w;
  s.da2dy = s.p[0].x * s.p[1].w - s.p[1].x * s.p[0].w;
  s.datdx = s.da0dx + s.da1dx + s.da2dx;
  s.datdy = s.da0dy + s.da1dy + s.da2dy;
  float dfxdx = xs * s.iw, dfydy = ys * s.iw;
  s.dudx = dfxdx * (s.b0 * s.datdx - s.da0dx);
  s.dudy = dfydy * (s.b0 * s.datdy - s.da0dy);
  s.dvdx = dfxdx * (s.b1 * s.datdx - s.da1dx);
  s.dvdy = dfydy * (s.b1 * s.datdy - s.da1dy);
}

// z/w only (used by antialias to pick the nearer surface)
VH_HD float tri_zw(const RenderArgs& A, int b, int px, int py, int tri) {
  TriSetup s; tri_setup(A, b, px, py, tri, s); return s.zw;
}

// Backward of tri_setup: gradients wrt (u, v) [clamped outputs] and the derivative block -> clip positions.
// gpos[k] receives d/d(x,y,_,w) of vertex k (z never receives gradient).
VH_HD void tri_setup_bwd(const RenderArgs& A, const TriSetup& s, float g_u, float g_v,
                         float g_dudx, float g_dudy, float g_dvdx, float g_dvdy, f4 gpos[3]) {
  float xs = 2.f / A.W, ys = 2.f / A.H;
  float dfxdx = xs * s.iw, dfydy = ys * s.iw;
  // through the clamped barycentrics
  float Gu = g_u + g_dudx * dfxdx * s.datdx + g_dudy * dfydy * s.datdy;
  float Gv = g_v + g_dvdx * dfxdx * s.datdx + g_dvdy * dfydy * s.datdy;
  if (!(s.b0u > 0.f && s.b0u < 1.f)) Gu = 0.f;
  if (!(s.b1u > 0.f && s.b1u < 1.f)) Gv = 0.f;
  float g_iw = Gu * s.a0 + Gv * s.a1
             + g_dudx * xs * (s.b0 * s.datdx - s.da0dx) + g_dudy * ys * (s.b0 * s.datdy - s.da0dy)
             + g_dvdx * xs * (s.b1 * s.datdx - s.da1dx) + g_dvdy * ys * (s.b1 * s.datdy - s.da1dy);
  float g_at = -g_iw * s.iw * s.iw;
  float g_a0 = Gu * s.iw + g_at, g_a1 = Gv * s.iw + g_at, g_a2 = g_at;
  float g_datdx = (g_dudx * s.b0 + g_dvdx * s.b1) * dfxdx;
  float g_datdy = (g_dudy * s.b0 + g_dvdy * s.b1) * dfydy;
  float g_da0dx = g_datdx - g_dudx * dfxdx, g_da0dy = g_datdy - g_dudy * dfydy;
  float g_da1dx = g_datdx - g_dvdx * dfxdx, g_da1dy = g_datdy - g_dvdy * dfydy;
  float g_da2dx = g_datdx, g_da2dy = g_datdy;
  const f4* p = s.p;
  float gx[3] = {0, 0, 0}, gy[3] = {0, 0, 0}, gw[3] = {0, 0, 0};
  // da0dx = p2.y*p1.w - p1.y*p2.w ; da0dy = p1.x*p2.w - p2.x*p1.w
  gy[2] += g_da0dx * p[1].w; gw[1] += g_da0dx * p[2].y; gy[1] -= g_da0dx * p[2].w; gw[2] -= g_da0dx * p[1].y;
  gx[1] += g_da0dy * p[2].w; gw[2] += g_da0dy * p[1].x; gx[2] -= g_da0dy * p[1].w; gw[1] -= g_da0dy * p[2].x;
  // da1dx = p0.y*p2.w - p2.y*p0.w ; da1dy = p2.x*p0.w - p0.x*p2.w
  gy[0] += g_da1dx * p[2].w; gw[2] += g_da1dx * p[0].y; gy[2] -= g_da1dx * p[0].w; gw[0] -= g_da1dx * p[2].y;
  gx[2] += g_da1dy * p[0].w; gw[0] += g_da1dy * p[2].x; gx[0] -= g_da1dy * p[2].w; gw[2] -= g_da1dy * p[0].x;
  // da2dx = p1.y*p0.w - p0.y*p1.w ; da2dy = p0.x*p1.w - p1.x*p0.w
  gy[1] += g_da2dx * p[0].w; gw[0] += g_da2dx * p[1].y; gy[0] -= g_da2dx * p[1].w; gw[1] -= g_da2dx * p[0].y;
  gx[0] += g_da2dy * p[1].w; gw[1] += g_da2dy * p[0].x; gx[1] -= g_da2dy * p[0].w; gw[0] -= g_da2dy * p[1].x;
  // a0 = p1x*p2y - p1y*p2x (pkx = p_k.x - fx*p_k.w, pky = p_k.y - fy*p_k.w)
  float px_[3], py_[3];
  for (int k = 0; k < 3; ++k) { px_[k] = p[k].x - s.fx * p[k].w; py_[k] = p[k].y - s.fy * p[k].w; }
  float gpx[3] = {0, 0, 0}, gpy[3] = {0, 0, 0};
  gpx[1] += g_a0 * py_[2]; gpy[2] += g_a0 * px_[1]; gpy[1] -= g_a0 * px_[2]; gpx[2] -= g_a0 * py_[1];
  gpx[2] += g_a1 * py_[0]; gpy[0] += g_a1 * px_[2]; gpy[2] -= g_a1 * px_[0]; gpx[0] -= g_a1 * py_[2];
  gpx[0] += g_a2 * py_[1]; gpy[1] += g_a2 * px_[0]; gpy[0] -= g_a2 * px_[1]; gpx[1] -= g_a2 * py_[0];
  for (int k = 0; k < 3; ++k) {
    gpos[k].x = gx[k] + gpx[k];
    gpos[k].y = gy[k] + gpy[k];
    gpos[k].z = 0.f;
    gpos[k].w = gw[k] - s.fx * gpx[k] - s.fy * gpy[k];
  }
}

// ------------------------------------------------------------------------------------------ SH shading
VH_HD void sh_basis(f3 n, float Bk[9]) {
  Bk[0] = VH_SH_C0; Bk[1] = VH_SH_C1 * n.x; Bk[2] = VH_SH_C1 * n.y; Bk[3] = VH_SH_C1 * n.z;
  Bk[4] = VH_SH_C2 * n.x * n.y; Bk[5] = VH_SH_C2 * n.x * n.z; Bk[6] = VH_SH_C2 * n.y * n.z;
  Bk[7] = VH_SH_C3 * (n.x * n.x - n.y * n.y); Bk[8] = VH_SH_C4 * (3.f * n.z * n.z - 1.f);
}
VH_HD f3 sh_eval(const float Bk[9], const float* L) {
  f3 d = mk3(0, 0, 0);
  for (int k = 0; k < 9; ++k) { d.x += Bk[k] * L[k * 3 + 0]; d.y += Bk[k] * L[k * 3 + 1]; d.z += Bk[k] * L[k * 3 + 2]; }
  return d;
}
// d diffuse / d n contracted with g_diffuse: returns g_n
VH_HD f3 sh_bwd_normal(f3 n, const float* L, f3 gd) {
  float m[9];
  for (int k = 0; k < 9; ++k) m[k] = gd.x * L[k * 3] + gd.y * L[k * 3 + 1] + gd.z * L[k * 3 + 2];
  f3 g;
  g.x = VH_SH_C1 * m[1] + VH_SH_C2 * (n.y * m[4] + n.z * m[5]) + VH_SH_C3 * 2.f * n.x * m[7];
  g.y = VH_SH_C1 * m[2] + VH_SH_C2 * (n.x * m[4] + n.z * m[6]) - VH_SH_C3 * 2.f * n.y * m[7];
  g.z = VH_SH_C1 * m[3] + VH_SH_C2 * (n.x * m[5] + n.y * m[6]) + VH_SH_C4 * 6.f * n.z * m[8];
  return g;
}

// ------------------------------------------------------------------------------------------ texture
struct TexSample {
  float lvl;      // clamped continuous level
  int l0, l1;
  float f;        // blend weight of l1
  bool lvl_free;  // level not clamped -> gradient flows to uv_da
  f3 c0, c1;      // bilinear samples of l0 / l1
  f3 out;
  // LOD intermediates for the backward
  float dsdx, dsdy, dtdx, dtdy, l2n, l2a, major;
};

struct Bilin { int i00, i10, i01, i11; float fx, fy; int w; };

VH_HD void bilin_setup(float u, float v, int w, Bilin& q) {
  u = u - floorf(u); v = v - floorf(v);
  float x = u * w - 0.5f, y = v * w - 0.5f;
  float x0f = floorf(x), y0f = floorf(y);
  q.fx = x - x0f; q.fy = y - y0f;
  int x0 = (int)x0f, y0 = (int)y0f;
  if (x0 < 0) x0 += w; if (y0 < 0) y0 += w;
  if (x0 >= w) x0 -= w; if (y0 >= w) y0 -= w;
  int x1 = x0 + 1; if (x1 >= w) x1 -= w;
  int y1 = y0 + 1; if (y1 >= w) y1 -= w;
  q.i00 = y0 * w + x0; q.i10 = y0 * w + x1; q.i01 = y1 * w + x0; q.i11 = y1 * w + x1; q.w = w;
}

VH_HD f3 tex3(const f4* t, int i) { f4 v = t[i]; return mk3(v.x, v.y, v.z); }

VH_HD f3 bilin_fetch(const f4* lvl, const Bilin& q, f3* ddx = nullptr, f3* ddy = nullptr) {
  f3 t00 = tex3(lvl, q.i00), t10 = tex3(lvl, q.i10), t01 = tex3(lvl, q.i01), t11 = tex3(lvl, q.i11);
  f3 top = t00 * (1.f - q.fx) + t10 * q.fx, bot = t01 * (1.f - q.fx) + t11 * q.fx;
  if (ddx) *ddx = ((t10 - t00) * (1.f - q.fy) + (t11 - t01) * q.fy) * (float)q.w;   // d/du
  if (ddy) *ddy = (bot - top) * (float)q.w;                                            // d/dv
  return top * (1.f - q.fy) + bot * q.fy;
}

// linear-mipmap-linear sample (oracle/render.py mip_level / texture_sample)
// level selection only (no texel fetch): lvl, l0, l1, f, lvl_free and the LOD intermediates
VH_HD void tex_level(const RenderArgs& A, float dudx, float dudy, float dvdx, float dvdy, TexSample& s) {
  float T = (float)A.T;
  s.dsdx = dudx * T; s.dsdy = dudy * T; s.dtdx = dvdx * T; s.dtdy = dvdy * T;
  float Aq = s.dsdx * s.dsdx + s.dtdx * s.dtdx;
  float Bq = s.dsdy * s.dsdy + s.dtdy * s.dtdy;
  float Cq = s.dsdx * s.dsdy + s.dtdx * s.dtdy;
  float l2b = 0.5f * (Aq + Bq);
  s.l2n = 0.25f * (Aq - Bq) * (Aq - Bq) + Cq * Cq;
  s.l2a = s.l2n > 0.f ? sqrtf(s.l2n) : 0.f;
  s.major = l2b + s.l2a;
  float raw = 0.5f * log2f(fmaxf(s.major, 1e-30f));
  s.lvl_free = (raw > 0.f) && (raw < (float)A.max_level) && (s.major > 1e-30f);
  s.lvl = fminf(fmaxf(raw, 0.f), (float)A.max_level);
  s.l0 = (int)floorf(s.lvl); if (s.l0 > A.max_level) s.l0 = A.max_level;
  s.l1 = s.l0 + 1 < A.max_level ? s.l0 + 1 : A.max_level;
  s.f = s.lvl - (float)s.l0;
}
VH_HD void tex_sample(const RenderArgs& A, float u, float v, float dudx, float dudy, float dvdx, float dvdy, TexSample& s) {
  tex_level(A, dudx, dudy, dvdx, dvdy, s);
  Bilin q0, q1;
  bilin_setup(u, v, A.T >> s.l0, q0);
  s.c0 = bilin_fetch(A.mips + A.mip_off[s.l0], q0);
  if (s.l1 != s.l0) { bilin_setup(u, v, A.T >> s.l1, q1); s.c1 = bilin_fetch(A.mips + A.mip_off[s.l1], q1); }
  else s.c1 = s.c0;
  s.out = s.c0 * (1.f - s.f) + s.c1 * s.f;
}

VH_HD void bilin_scatter(float* gl, const Bilin& q, f3 g) {
  float w00 = (1.f - q.fx) * (1.f - q.fy), w10 = q.fx * (1.f - q.fy), w01 = (1.f - q.fx) * q.fy, w11 = q.fx * q.fy;
  const int idx[4] = {q.i00, q.i10, q.i01, q.i11};
  const float ww[4] = {w00, w10, w01, w11};
  for (int k = 0; k < 4; ++k) {
    float* t = gl + (size_t)idx[k] * 4;
    VH_ATOMIC_ADD4(t, g.x * ww[k], g.y * ww[k], g.z * ww[k], 0.f);
  }
}

// Backward of tex_sample: scatters texel gradients into the gradient pyramid (same layout as the mips, float4) and
// returns gradients wrt uv and the uv derivative block.
VH_HD void tex_sample_bwd(const RenderArgs& A, float u, float v, const TexSample& s, f3 g_out, float* grad_pyr,
                          float& g_u, float& g_v, float g_da[4]) {
  Bilin q0, q1;
  f3 dx0, dy0, dx1 = mk3(0, 0, 0), dy1 = mk3(0, 0, 0);
  bilin_setup(u, v, A.T >> s.l0, q0);
  bilin_fetch(A.mips + A.mip_off[s.l0], q0, &dx0, &dy0);
  float w0 = 1.f - s.f, w1 = s.f;
  if (s.l1 != s.l0) {
    bilin_setup(u, v, A.T >> s.l1, q1);
    bilin_fetch(A.mips + A.mip_off[s.l1], q1, &dx1, &dy1);
  } else { w0 = 1.f; w1 = 0.f; }
  if (grad_pyr) {
    bilin_scatter(grad_pyr + (size_t)A.mip_off[s.l0] * 4, q0, g_out * w0);
    if (s.l1 != s.l0) bilin_scatter(grad_pyr + (size_t)A.mip_off[s.l1] * 4, q1, g_out * w1);
  }
  g_u = dot3(g_out, dx0) * w0 + dot3(g_out, dx1) * w1;
  g_v = dot3(g_out, dy0) * w0 + dot3(g_out, dy1) * w1;
  g_da[0] = g_da[1] = g_da[2] = g_da[3] = 0.f;
  if (s.lvl_free && s.l1 != s.l0) {
    float g_lvl = dot3(g_out, s.c1 - s.c0);
    float g_major = g_lvl * 0.5f / (s.major * 0.6931471805599453f);
    float g_l2b = g_major, g_l2a = g_major;
    float g_l2n = s.l2n > 0.f ? g_l2a * 0.5f / s.l2a : 0.f;
    float Aq = s.dsdx * s.dsdx + s.dtdx * s.dtdx;
    float Bq = s.dsdy * s.dsdy + s.dtdy * s.dtdy;
    float Cq = s.dsdx * s.dsdy + s.dtdx * s.dtdy;
    float g_A = 0.5f * g_l2b + g_l2n * 0.5f * (Aq - Bq);
    float g_B = 0.5f * g_l2b - g_l2n * 0.5f * (Aq - Bq);
    float g_C = g_l2n * 2.f * Cq;
    float T = (float)A.T;
    g_da[0] = (g_A * 2.f * s.dsdx + g_C * s.dsdy) * T;   // d/d(dudx)
    g_da[1] = (g_B * 2.f * s.dsdy + g_C * s.dsdx) * T;   // d/d(dudy)
    g_da[2] = (g_A * 2.f * s.dtdx + g_C * s.dtdy) * T;   // d/d(dvdx)
    g_da[3] = (g_B * 2.f * s.dtdy + g_C * s.dtdx) * T;   // d/d(dvdy)
  }
}

// ------------------------------------------------------------------------------------------ full fg-pixel shading
struct PixShade {
  TriSetup ts;
  f3 n0, n1, n2, nraw, n;
  float inv_len; bool n_free;
  float t0[2], t1[2], t2[2];
  float u, v, uvda[4];
  TexSample tx;
  float Bk[9];
  f3 diffuse, albedo, rgb;
};

VH_HD void shade_pixel(const RenderArgs& A, int b, int px, int py, int tri, PixShade& s) {
  tri_setup(A, b, px, py, tri, s.ts);
  const f4* vn = A.vnorm + (size_t)(A.geo ? A.geo[b] : b) * A.V;
  f4 a = vn[s.ts.vi[0]], c = vn[s.ts.vi[1]], d = vn[s.ts.vi[2]];
  s.n0 = mk3(a.x, a.y, a.z); s.n1 = mk3(c.x, c.y, c.z); s.n2 = mk3(d.x, d.y, d.z);
  float b0 = s.ts.b0, b1 = s.ts.b1, b2 = 1.f - b0 - b1;
  s.nraw = s.n0 * b0 + s.n1 * b1 + s.n2 * b2;
  float l2 = dot3(s.nraw, s.nraw);
  s.n_free = l2 > 1e-20f;
  s.inv_len = 1.f / sqrtf(fmaxf(l2, 1e-20f));
  s.n = s.nraw * s.inv_len;
  i4 fu = A.faces_uv[tri];
  s.t0[0] = A.verts_uv[fu.x * 2]; s.t0[1] = A.verts_uv[fu.x * 2 + 1];
  s.t1[0] = A.verts_uv[fu.y * 2]; s.t1[1] = A.verts_uv[fu.y * 2 + 1];
  s.t2[0] = A.verts_uv[fu.z * 2]; s.t2[1] = A.verts_uv[fu.z * 2 + 1];
  s.u = b0 * s.t0[0] + b1 * s.t1[0] + b2 * s.t2[0];
  s.v = b0 * s.t0[1] + b1 * s.t1[1] + b2 * s.t2[1];
  float d0u = s.t0[0] - s.t2[0], d1u = s.t1[0] - s.t2[0], d0v = s.t0[1] - s.t2[1], d1v = s.t1[1] - s.t2[1];
  s.uvda[0] = s.ts.dudx * d0u + s.ts.dvdx * d1u;   // du_tex/dx
  s.uvda[1] = s.ts.dudy * d0u + s.ts.dvdy * d1u;   // du_tex/dy
  s.uvda[2] = s.ts.dudx * d0v + s.ts.dvdx * d1v;   // dv_tex/dx
  s.uvda[3] = s.ts.dudy * d0v + s.ts.dvdy * d1v;   // dv_tex/dy
  tex_sample(A, s.u, s.v, s.uvda[0], s.uvda[1], s.uvda[2], s.uvda[3], s.tx);
  s.albedo = s.tx.out;
  sh_basis(s.n, s.Bk);
  s.diffuse = sh_eval(s.Bk, A.lights);
  s.rgb = s.albedo * s.diffuse;
}

// Backward of shade_pixel for an upstream gradient g_rgb on the composite colour of this (foreground) pixel.
//   g_clip  [B,V,4]  atomics          g_vnorm [B,V,4] atomics        grad_pyr texture-gradient pyramid (float4) or NULL
//   g_lights_local[27] per-thread accumulator (reduced by the caller)
//   g_dd: extra gradient on diffuse_detach_normal (reg_diffuse), flows to the lights only.
// raise the level-0 region flags of the four texels of a level-0 bilinear footprint (texture.cu: one flag per 8-row x 256-column region)
VH_HD void tex_flag_l0(const RenderArgs& A, float u, float v) {
  Bilin q0; bilin_setup(u, v, A.T, q0);
  const int tpr = A.T >= 256 ? A.T >> 8 : 1;
  const int idx[4] = {q0.i00, q0.i10, q0.i01, q0.i11};
  for (int k = 0; k < 4; ++k) { int ty = idx[k] / A.T, tx = idx[k] - ty * A.T; A.tex_l0_flag[(ty >> 3) * tpr + (tx >> 8)] = 1; }
}

// The TEXEL-GRADIENT half of the shading adjoint of one foreground pixel, on its own: d L / d albedo = g_rgb * diffuse scattered into the
// gradient pyramid with the trilinear weights.  Needs the barycentrics / uv derivatives (level of detail) and the shaded diffuse term, but
// neither the texel values nor any vertex gradient -- k_passC2<1> runs it beside the geometry half (k_passC2<2>) on a second stream.
VH_HD void shade_pixel_texgrad(const RenderArgs& A, int b, int px, int py, int tri, f3 g_rgb, float* grad_pyr) {
  TriSetup ts;
  tri_setup(A, b, px, py, tri, ts);
  const f4* vn = A.vnorm + (size_t)(A.geo ? A.geo[b] : b) * A.V;
  f4 a = vn[ts.vi[0]], c = vn[ts.vi[1]], d = vn[ts.vi[2]];
  float b0 = ts.b0, b1 = ts.b1, b2 = 1.f - b0 - b1;
  f3 nraw = mk3(a.x, a.y, a.z) * b0 + mk3(c.x, c.y, c.z) * b1 + mk3(d.x, d.y, d.z) * b2;
  f3 n = nraw * (1.f / sqrtf(fmaxf(dot3(nraw, nraw), 1e-20f)));
  float Bk[9];
  sh_basis(n, Bk);
  f3 g_alb = g_rgb * sh_eval(Bk, A.lights);
  i4 fu = A.faces_uv[tri];
  float t0u = A.verts_uv[fu.x * 2], t0v = A.verts_uv[fu.x * 2 + 1], t1u = A.verts_uv[fu.y * 2], t1v = A.verts_uv[fu.y * 2 + 1];
  float t2u = A.verts_uv[fu.z * 2], t2v = A.verts_uv[fu.z * 2 + 1];
  float u = b0 * t0u + b1 * t1u + b2 * t2u, v = b0 * t0v + b1 * t1v + b2 * t2v;
  float d0u = t0u - t2u, d1u = t1u - t2u, d0v = t0v - t2v, d1v = t1v - t2v;
  TexSample tx;
  tex_level(A, ts.dudx * d0u + ts.dvdx * d1u, ts.dudy * d0u + ts.dvdy * d1u, ts.dudx * d0v + ts.dvdx * d1v, ts.dudy * d0v + ts.dvdy * d1v, tx);
  Bilin q0, q1;
  bilin_setup(u, v, A.T >> tx.l0, q0);
  if (tx.l1 != tx.l0) {
    bilin_scatter(grad_pyr + (size_t)A.mip_off[tx.l0] * 4, q0, g_alb * (1.f - tx.f));
    bilin_setup(u, v, A.T >> tx.l1, q1);
    bilin_scatter(grad_pyr + (size_t)A.mip_off[tx.l1] * 4, q1, g_alb * tx.f);
  } else bilin_scatter(grad_pyr + (size_t)A.mip_off[tx.l0] * 4, q0, g_alb);
  if (A.tex_l0_flag && tx.l0 == 0) tex_flag_l0(A, u, v);
}

// vg != NULL: the per-vertex gradients of this pixel are RETURNED (for a warp-level reduction over the pixels of one triangle,
// k_passC2) instead of being added to g_clip / g_vnorm with one vector reduction per vertex per pixel.
struct VertGrad { f3 gn[3]; f3 gc[3]; };      // d/d vertex normal (xyz), d/d clip position (x, y, w) of the triangle's three vertices
VH_HD void shade_pixel_bwd(const RenderArgs& A, int b, int tri, const PixShade& s, f3 g_rgb, f3 g_dd,
                           float* g_clip, float* g_vnorm, float* grad_pyr, float* g_lights_local, VertGrad* vg = nullptr) {
  f3 g_alb = g_rgb * s.diffuse, g_dif = g_rgb * s.albedo;
  for (int k = 0; k < 9; ++k) {
    g_lights_local[k * 3 + 0] += s.Bk[k] * (g_dif.x + g_dd.x);
    g_lights_local[k * 3 + 1] += s.Bk[k] * (g_dif.y + g_dd.y);
    g_lights_local[k * 3 + 2] += s.Bk[k] * (g_dif.z + g_dd.z);
  }
  f3 g_n = sh_bwd_normal(s.n, A.lights, g_dif);
  f3 g_raw = s.n_free ? (g_n - s.n * dot3(s.n, g_n)) * s.inv_len : g_n * s.inv_len;
  float b0 = s.ts.b0, b1 = s.ts.b1, b2 = 1.f - b0 - b1;
  if (vg) { vg->gn[0] = g_raw * b0; vg->gn[1] = g_raw * b1; vg->gn[2] = g_raw * b2; }
  else if (g_vnorm) {
    float* gv = g_vnorm + (size_t)(A.geo ? A.geo[b] : b) * A.V * 4;
    const f3 gg[3] = {g_raw * b0, g_raw * b1, g_raw * b2};
    for (int k = 0; k < 3; ++k) {
      float* t = gv + (size_t)s.ts.vi[k] * 4;
      VH_ATOMIC_ADD4(t, gg[k].x, gg[k].y, gg[k].z, 0.f);
    }
  }
  float g_b0 = dot3(g_raw, s.n0 - s.n2), g_b1 = dot3(g_raw, s.n1 - s.n2);
  // texture
  float g_u, g_v, g_da[4];
  tex_sample_bwd(A, s.u, s.v, s.tx, g_alb, grad_pyr, g_u, g_v, g_da);
  if (grad_pyr && A.tex_l0_flag && s.tx.l0 == 0) tex_flag_l0(A, s.u, s.v);
  float d0u = s.t0[0] - s.t2[0], d1u = s.t1[0] - s.t2[0], d0v = s.t0[1] - s.t2[1], d1v = s.t1[1] - s.t2[1];
  bool detach_uv = A.face_flags && (A.face_flags[tri] & 1);
  if (!detach_uv) { g_b0 += g_u * d0u + g_v * d0v; g_b1 += g_u * d1u + g_v * d1v; }
  float g_dudx = g_da[0] * d0u + g_da[2] * d0v;
  float g_dvdx = g_da[0] * d1u + g_da[2] * d1v;
  float g_dudy = g_da[1] * d0u + g_da[3] * d0v;
  float g_dvdy = g_da[1] * d1u + g_da[3] * d1v;
  f4 gp[3];
  tri_setup_bwd(A, s.ts, g_b0, g_b1, g_dudx, g_dudy, g_dvdx, g_dvdy, gp);
  if (vg) { for (int k = 0; k < 3; ++k) vg->gc[k] = mk3(gp[k].x, gp[k].y, gp[k].w); }
  else if (g_clip) {
    float* gc = g_clip + (size_t)b * A.V * 4;
    for (int k = 0; k < 3; ++k) {
      float* t = gc + (size_t)s.ts.vi[k] * 4;
      VH_ATOMIC_ADD4(t, gp[k].x, gp[k].y, 0.f, gp[k].w);
    }
  }
}
