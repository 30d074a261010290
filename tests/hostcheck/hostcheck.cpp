// Host-compiled unit checks of the device math headers (developer aid; NOT part of libvhap_b200.so, never used by
// the product path).  Built by tests/hostcheck/build.py with g++ -DVH_HOST_CHECK and driven through ctypes by
// tests/test_hostcheck_math.py, which compares the analytic forward/backward formulas with the oracle's autograd.
#define VH_HOST_CHECK 1
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../vhap_b200/csrc/flame_math.cuh"
#include "../../vhap_b200/csrc/render_bodies.cuh"

extern "C" {

void hc_pose(const float* pose, const float* J, const float* gA, const float* gpf,
             float* A_out, float* pf_out, float* g_pose, float* g_J) {
  PoseFwd f;
  memcpy(f.J, J, sizeof(f.J));
  pose_forward(pose, f);
  memcpy(A_out, f.A, sizeof(f.A));
  memcpy(pf_out, f.pf, sizeof(f.pf));
  float gj[VH_NJ][3];
  pose_backward(pose, f, (const float(*)[12])gA, gpf, g_pose, gj);
  memcpy(g_J, gj, sizeof(gj));
}

// Full serial emulation of passes A, pool compaction, B, finalize, C.
// out_sums: [0] sum var  [1] n fg pixels (raster)  [2] abs err sum  [3] n alpha_aa>0  [4] max diffuse
int hc_render(int B, int H, int W, int V, int F, int T, int max_level,
              const int* faces4, const int* faces_uv4, const float* verts_uv, const float* clip, const float* vnorm,
              const float* lights, const float* mips, const int* mip_off, const int* tri_id,
              const uint8_t* face_flags, const uint8_t* vert_flags, const uint8_t* fid2cid, const int* adj_opp4,
              const uint16_t* target, int disturb, const uint8_t* inj_w, const float* inj_u,
              float w_photo, float w_regdiff, int do_backward,
              float* pre, float* final_rgba, float* out_sums,
              float* g_clip, float* g_vnorm, float* g_tex, float* g_lights) {
  PassArgs P;
  memset(&P, 0, sizeof(P));
  RenderArgs& A = P.R;
  A.B = B; A.H = H; A.W = W; A.V = V; A.F = F; A.T = T; A.max_level = max_level;
  if ((W & (W - 1)) == 0 && (H & (H - 1)) == 0) {          // like fill_render_args (render.cu): shift/mask pixel unflatten for power-of-two images
    A.pow2 = 1;
    while ((1 << A.wshift) < W) ++A.wshift;
    while ((1 << A.hshift) < H) ++A.hshift;
  }
  A.faces = (const i4*)faces4; A.faces_uv = (const i4*)faces_uv4; A.verts_uv = verts_uv;
  A.clip = (const f4*)clip; A.vnorm = (const f4*)vnorm; A.lights = lights; A.mips = (const f4*)mips;
  for (int i = 0; i <= max_level; ++i) A.mip_off[i] = mip_off[i];
  A.tri_id = tri_id; A.face_flags = face_flags; A.vert_flags = vert_flags; A.fid2cid = fid2cid; A.adj_opp = adj_opp4;
  P.target = target; P.pre = (f4*)pre; P.final_rgba = final_rgba;
  A.zwbuf = pre;
  std::vector<uint8_t> signs((size_t)B * H * W);
  P.signs = signs.data();
  P.disturb = disturb; P.rate_fg = 0.5f; P.rate_bg = 0.5f; P.inj_w = inj_w; P.inj_u = inj_u;
  P.bg_mode = 0;
  // pass A
  float accA[2] = {0, 0}; float mx = -1e30f; int mx_idx = -1;
  for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) passA_body(P, b, y, x, accA, mx, mx_idx);
  // pools
  int N = B * H * W;
  std::vector<int> list(N), base(16, 0), count(16, 0);
  for (int i = 0; i < N; ++i) count[fid2cid[tri_id[i]]]++;
  for (int c = 1; c < 16; ++c) base[c] = base[c - 1] + count[c - 1];
  { std::vector<int> cur(base); for (int i = 0; i < N; ++i) list[cur[fid2cid[tri_id[i]]]++] = i; }
  P.pool_list = list.data(); P.pool_base = base.data(); P.pool_count = count.data();
  // pair analysis
  std::vector<float> aa_code((size_t)N * 2, 0.f);
  P.aa_code = aa_code.data();
  for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
    int i = (b * H + y) * W + x;
    if (x + 1 < W && tri_id[i] != tri_id[i + 1]) aa_pair_body(P, b, y, x, 0);
    if (y + 1 < H && tri_id[i] != tri_id[i + W]) aa_pair_body(P, b, y, x, 1);
  }
  // pass B
  float accB[2] = {0, 0};
  for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) passB_body(P, b, y, x, accB);
  // bg diffuse (normal = 0): SH basis {C0,0,0,0,0,0,0,0,-C4}
  float dbg[3];
  for (int c = 0; c < 3; ++c) dbg[c] = VH_SH_C0 * lights[c] - VH_SH_C4 * lights[24 + c];
  float mbg = (dbg[0] + dbg[1] + dbg[2]) / 3.f, vbg = 0;
  for (int c = 0; c < 3; ++c) vbg += 0.5f * (dbg[c] - mbg) * (dbg[c] - mbg);
  float n_bg = (float)N - accA[1];
  float mxbg = std::max(dbg[0], std::max(dbg[1], dbg[2]));
  bool bg_is_max = n_bg > 0 && mxbg > mx;
  out_sums[0] = accA[0] + n_bg * vbg; out_sums[1] = accA[1]; out_sums[2] = accB[0]; out_sums[3] = accB[1];
  out_sums[4] = bg_is_max ? mxbg : mx;
  if (!do_backward) return 0;
  float scal[8] = {0};
  scal[0] = w_photo / (3.f * accB[1]);
  scal[1] = w_regdiff >= 0 ? w_regdiff / (float)N : 0.f;
  float gmax = (w_regdiff >= 0 && out_sums[4] > 1.f) ? w_regdiff : 0.f;
  scal[2] = bg_is_max ? 0.f : gmax;
  ((int*)scal)[3] = mx_idx;
  P.scal = scal;
  P.g_clip = g_clip; P.g_vnorm = g_vnorm; P.g_tex = g_tex;
  for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) passC_body(P, b, y, x, nullptr, g_lights);
  // background pixels' contribution to the light gradient (all share the basis of n = 0)
  if (scal[1] != 0.f) {
    int chm = dbg[0] >= dbg[1] ? (dbg[0] >= dbg[2] ? 0 : 2) : (dbg[1] >= dbg[2] ? 1 : 2);
    for (int c = 0; c < 3; ++c) {
      float gd = n_bg * (dbg[c] - mbg) * scal[1] + ((bg_is_max && c == chm) ? gmax : 0.f);
      g_lights[0 * 3 + c] += VH_SH_C0 * gd;
      g_lights[8 * 3 + c] += -VH_SH_C4 * gd;
    }
  }
  return 0;
}

}  // extern "C"
