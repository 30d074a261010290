// C-ABI entry points of libvhap_b200.so (include/vhap_b200.h) and context management.
#include <stdlib.h>
#include <vector>
#include "engine.h"
#include "accum.h"

void launch_forward_slab(vhap_ctx* c, const PassArgs& P, const float* lights, float* slab, cudaStream_t s);
void launch_finalize(vhap_ctx* c, const PassArgs& P, const vhap_stage_cfg* cfg, const float* slab_global, const float* slab_local, int global_B,
                     const float* lights, float* g_lights, cudaStream_t s);
void launch_flip_plane(vhap_ctx* c, const float* in, float* out, int B, int H, int W, cudaStream_t s);
void launch_cid_plane(vhap_ctx* c, float* out, cudaStream_t s);

static char g_err[512] = "";

void vh_set_error(vhap_ctx* ctx, const char* what, const char* msg) {
  char* dst = ctx ? ctx->err : g_err;
  snprintf(dst, 512, "%s: %s", what, msg);
}

#define CK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { vh_set_error(ctx, #expr, cudaGetErrorString(_e)); return -2; } } while (0)

template <typename T>
static int upload(vhap_ctx* ctx, T** dst, const T* src, size_t n) {
  CK(cudaMalloc((void**)dst, n * sizeof(T)));
  if (src) CK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  else CK(cudaMemset(*dst, 0, n * sizeof(T)));
  return 0;
}
#define UP(dst, src, n) do { if (upload(ctx, &(dst), (src), (n))) return -2; } while (0)

extern "C" int vhap_abi_version(void) { return 2; }   // 2: vhap_frame_batch::target_format (took the padding after W), loss mask, lr scale

static const char* const KID_NAMES[KID_COUNT] = {
    "cam_setup", "pose_fwd", "blend_fwd", "skin_fwd", "landmarks", "vnormals", "vnormals_bwd", "skin_bwd", "pose_bwd", "joff_bwd", "blend_bwd",
    "betas_scatter", "regs", "snap", "bin", "scan", "fine_raster", "rast_out", "passA_shade", "pool_count", "pool_scan", "pool_scatter",
    "aa_pairs", "passB_disturb_aa_loss", "reduce_partials", "forward_slab", "finalize", "passC1_color_adjoint", "passC_backward", "lights_reduce", "tex_level0", "mip_down",
    "tex_fold_reg_adam", "tex_loss_reduce", "adam", "assemble_losses", "misc"};

// per-kernel device time, measured with CUDA events on the launching stream (see LAUNCH in engine.h)
extern "C" int vhap_profile_enable(vhap_ctx* ctx, int32_t on) {
  VhProf* p = ctx->prof;
  if (on && !p->ev[0][0][0])
    for (int k = 0; k < KID_COUNT; ++k) for (int i = 0; i < VH_PROF_SLOTS; ++i) { CK(cudaEventCreate(&p->ev[k][i][0])); CK(cudaEventCreate(&p->ev[k][i][1])); }
  p->on = on;
  if (on >= 0) { for (int k = 0; k < KID_COUNT; ++k) { p->n[k] = 0; p->launches[k] = 0; } p->first = nullptr; }
  else p->on = 0;                                  // on < 0: stop recording but keep the recorded slots (graph timeline)
  return 0;
}
// synchronises; start / end (ms after the first recorded event) of every recorded launch, in recording order per kernel.
// With vhap_profile_enable(ctx, 2) during stream capture the events are nodes of the graph: after a replay this is the
// device timeline of that replay (all streams), the tool for critical-path analysis of the overlapped step.
extern "C" int vhap_profile_timeline(vhap_ctx* ctx, int32_t* kid_out, float* t0_ms, float* t1_ms, int32_t max_n) {
  VhProf* p = ctx->prof;
  CK(cudaDeviceSynchronize());
  int n = 0;
  if (!p->first) return 0;
  for (int k = 0; k < KID_COUNT; ++k)
    for (int i = 0; i < p->n[k] && n < max_n; ++i) {
      float a = 0, b = 0;
      if (cudaEventElapsedTime(&a, p->first, p->ev[k][i][0]) != cudaSuccess || cudaEventElapsedTime(&b, p->first, p->ev[k][i][1]) != cudaSuccess) { cudaGetLastError(); continue; }
      kid_out[n] = k; t0_ms[n] = a; t1_ms[n] = b; ++n;
    }
  return n;
}
// 0: run every kernel chain on the caller's stream (per-kernel timing without co-running kernels); 1 (default): fork/join
extern "C" int vhap_set_overlap(vhap_ctx* ctx, int32_t on) { ctx->no_overlap = on ? 0 : 1; return 0; }
extern "C" int vhap_profile_kernel_count(void) { return KID_COUNT; }
extern "C" const char* vhap_profile_kernel_name(int32_t kid) { return (kid >= 0 && kid < KID_COUNT) ? KID_NAMES[kid] : ""; }
// synchronises; avg_ms[k] = mean device time of the recorded launches of kernel k, launches[k] = launches since enable/reset
extern "C" int vhap_profile_read(vhap_ctx* ctx, float* avg_ms_host, uint64_t* launches_host) {
  VhProf* p = ctx->prof;
  CK(cudaDeviceSynchronize());
  for (int k = 0; k < KID_COUNT; ++k) {
    double s = 0;
    for (int i = 0; i < p->n[k]; ++i) { float ms = 0; cudaEventElapsedTime(&ms, p->ev[k][i][0], p->ev[k][i][1]); s += ms; }
    avg_ms_host[k] = p->n[k] ? (float)(s / p->n[k]) : 0.f;
    launches_host[k] = p->launches[k];
  }
  return 0;
}

extern "C" const char* vhap_last_error(const vhap_ctx* ctx) { return ctx ? ctx->err : g_err; }

extern "C" int vhap_ctx_create(vhap_ctx** out, const vhap_mesh_desc* m, int32_t tex_size, int32_t device) {
  vhap_ctx* ctx = (vhap_ctx*)calloc(1, sizeof(vhap_ctx));
  if (!ctx) return -1;
  *out = ctx;
  ctx->device = device;
  ctx->prof = (VhProf*)calloc(1, sizeof(VhProf));
  CK(cudaSetDevice(device));
  int V = m->V, F = m->F, K = m->K;
  ctx->V = V; ctx->F = F; ctx->VT = m->VT; ctx->K = K; ctx->n_shape = m->n_shape; ctx->n_expr = K - m->n_shape; ctx->n_lmk = m->n_lmk;
  ctx->n_clusters = m->n_clusters; ctx->T = tex_size;
  if (K > 512) { vh_set_error(ctx, "vhap_ctx_create", "K > 512 unsupported"); return -3; }
  if (tex_size & (tex_size - 1)) { vh_set_error(ctx, "vhap_ctx_create", "texture size must be a power of two"); return -3; }
  size_t M = (size_t)3 * V;
  UP(ctx->v_template, m->v_template_host, M);
  UP(ctx->S_bwd, m->shapedirs_host, M * K);                    // [3V][K] as given
  {
    std::vector<float> st(M * K);
    for (size_t r = 0; r < M; ++r) for (int k = 0; k < K; ++k) st[(size_t)k * M + r] = m->shapedirs_host[r * K + k];
    UP(ctx->S_fwd, st.data(), M * K);                          // [K][3V]
    ctx->Mpad = (int)((M + 3) & ~(size_t)3);
    std::vector<float> stp((size_t)K * ctx->Mpad, 0.f);
    for (int k = 0; k < K; ++k) memcpy(&stp[(size_t)k * ctx->Mpad], &st[(size_t)k * M], M * sizeof(float));
    UP(ctx->S_fwd_pad, stp.data(), (size_t)K * ctx->Mpad);
    { const char* e = getenv("VHAP_B200_BLEND"); ctx->use_tc_blend = !(e && strcmp(e, "simt") == 0); }
    { const char* e = getenv("VHAP_B200_TEXFOLD"); ctx->tex_fold_reg = (e && strcmp(e, "reg") == 0); }
    // JS[k][j*3+c] = sum_v Jreg[j][v] S[v][c][k] ; Jt = Jreg template
    std::vector<float> js((size_t)K * 15, 0.f), jt(15, 0.f);
    for (int j = 0; j < 5; ++j)
      for (int v = 0; v < V; ++v) {
        float w = m->J_regressor_host[(size_t)j * V + v];
        if (w == 0.f) continue;
        for (int c = 0; c < 3; ++c) {
          jt[j * 3 + c] += w * m->v_template_host[(size_t)v * 3 + c];
          const float* row = m->shapedirs_host + ((size_t)v * 3 + c) * K;
          for (int k = 0; k < K; ++k) js[(size_t)k * 15 + j * 3 + c] += w * row[k];
        }
      }
    UP(ctx->JS, js.data(), (size_t)K * 15);
    UP(ctx->Jt, jt.data(), (size_t)15);
  }
  UP(ctx->posedirs, m->posedirs_host, (size_t)36 * M);
  UP(ctx->Jreg, m->J_regressor_host, (size_t)5 * V);
  UP(ctx->lbs_w, m->lbs_weights_host, (size_t)V * 5);
  {
    std::vector<i4> f4v(F), fu(F);
    std::vector<int> adj((size_t)F * 4);
    for (int i = 0; i < F; ++i) {
      f4v[i] = {m->faces_host[i * 3], m->faces_host[i * 3 + 1], m->faces_host[i * 3 + 2], 0};
      fu[i] = {m->faces_uv_host[i * 3], m->faces_uv_host[i * 3 + 1], m->faces_uv_host[i * 3 + 2], 0};
      for (int k = 0; k < 3; ++k) adj[(size_t)i * 4 + k] = m->adj_opp_host[i * 3 + k];
      adj[(size_t)i * 4 + 3] = 0;
    }
    UP(ctx->faces, f4v.data(), (size_t)F);
    UP(ctx->faces_uv, fu.data(), (size_t)F);
    UP(ctx->adj_opp, adj.data(), (size_t)F * 4);
  }
  UP(ctx->verts_uv, m->verts_uv_host, (size_t)m->VT * 2);
  UP(ctx->lmk_faces, m->lmk_faces_host, (size_t)m->n_lmk);
  UP(ctx->lmk_bary, m->lmk_bary_host, (size_t)m->n_lmk * 3);
  UP(ctx->fid2cid, m->fid2cid_host, (size_t)F + 1);
  UP(ctx->vf_indptr, m->vf_indptr_host, (size_t)V + 1);
  UP(ctx->vf_faces, m->vf_faces_host, (size_t)3 * F);
  ctx->lap_nnz = m->lap_indptr_host[V];
  UP(ctx->lap_indptr, m->lap_indptr_host, (size_t)V + 1);
  UP(ctx->lap_idx, m->lap_indices_host, (size_t)ctx->lap_nnz);
  UP(ctx->lap_val, m->lap_values_host, (size_t)ctx->lap_nnz);
  UP(ctx->lap_y, (const float*)nullptr, M);
  UP(ctx->face_flags, (const uint8_t*)nullptr, (size_t)F);
  UP(ctx->vert_flags, (const uint8_t*)nullptr, (size_t)V);
  // texture pyramid
  int T = tex_size, l = 0; size_t off = 0;
  for (int s = T; s >= 1; s >>= 1, ++l) { ctx->mip_off[l] = (int)off; off += (size_t)s * s; }
  ctx->max_level = l - 1; ctx->mip_total = off;
  UP(ctx->mips[0], (const f4*)nullptr, off);
  UP(ctx->mips[1], (const f4*)nullptr, off);
  UP(ctx->g_tex, (const float*)nullptr, off * 4);
  ctx->tv_nblocks = (int)(((size_t)T * T + 255) / 256) + T;     // >= CTAs of the texture fold kernel for any T
  UP(ctx->tv_partials, (const float*)nullptr, (size_t)ctx->tv_nblocks * 2);
  UP(ctx->scal, (const float*)nullptr, (size_t)16);
  UP(ctx->acc, (const float*)nullptr, (size_t)ACC_COUNT);
  UP(ctx->maxslot, (const unsigned long long*)nullptr, (size_t)1);
  UP(ctx->overflow_flag, (const int*)nullptr, (size_t)1);
  UP(ctx->pool_base, (const int*)nullptr, (size_t)16);
  UP(ctx->pool_count, (const int*)nullptr, (size_t)16);
  UP(ctx->scan_aux, (const int*)nullptr, (size_t)1024);
  UP(ctx->scan_total, (const int*)nullptr, (size_t)1);
  UP(ctx->pair_count, (const int*)nullptr, (size_t)1);
  UP(ctx->dev_step, (const int*)nullptr, (size_t)2);
  { float one = 1.f; UP(ctx->dev_lr_scale, &one, (size_t)1); }
  // aux[0] carries small latency-critical side chains (pose chain, vertex normals, regularisers): highest priority, like hp[];
  // aux[1] carries the bulk texture pass: default (lowest) priority
  { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&ctx->aux[0], cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithFlags(&ctx->aux[1], cudaStreamNonBlocking)); }
  {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int i = 0; i < 2; ++i) CK(cudaStreamCreateWithPriority(&ctx->hp[i], cudaStreamNonBlocking, hi));
  }
  for (int i = 0; i < EV_COUNT; ++i) CK(cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
  ctx->n_l0_regions = (tex_size >= 8 ? tex_size / 8 : 1) * (tex_size >= 256 ? tex_size / 256 : 1);
  { std::vector<int> ones((size_t)ctx->n_l0_regions, 1); UP(ctx->tex_l0_flag, ones.data(), (size_t)ctx->n_l0_regions); }   // conservative first fold
  CK(cudaMalloc(&ctx->scan_state, (VH_SCAN_MAX_BLOCKS + 2) * sizeof(unsigned long long))); CK(cudaMemset(ctx->scan_state, 0, (VH_SCAN_MAX_BLOCKS + 2) * sizeof(unsigned long long)));
  CK(cudaMalloc(&ctx->tex_loss, 4 * sizeof(float))); CK(cudaMemset(ctx->tex_loss, 0, 4 * sizeof(float)));
  CK(cudaMalloc(&ctx->tex_counter, sizeof(unsigned))); CK(cudaMemset(ctx->tex_counter, 0, sizeof(unsigned)));
  return 0;
}

#define FREE(p) do { if (p) { cudaFree(p); p = nullptr; } } while (0)

static void free_batch(vhap_ctx* c) {
  FREE(c->v_shaped); FREE(c->v_shaped_part); FREE(c->v_posed); FREE(c->g_vshaped); FREE(c->verts); FREE(c->clip); FREE(c->vnorm); FREE(c->vnraw); FREE(c->snap); FREE(c->ndc);
  FREE(c->g_clip); FREE(c->g_vnorm); FREE(c->g_verts); FREE(c->posebuf); FREE(c->poses); FREE(c->gA); FREE(c->gpf); FREE(c->gJ); FREE(c->gbetas);
  FREE(c->betas); FREE(c->cam); FREE(c->tri_id); FREE(c->pre); FREE(c->signs); FREE(c->pool_list); FREE(c->pool_tri); FREE(c->final_rgba); FREE(c->plane_albedo);
  FREE(c->plane_normal); FREE(c->plane_diffuse); FREE(c->tile_count); FREE(c->tile_off); FREE(c->tile_cursor); FREE(c->tile_list);
  FREE(c->pool_blk_count); FREE(c->pool_blk_off); FREE(c->partials); FREE(c->aa_code); FREE(c->pair_list); FREE(c->grgb);
}

extern "C" int vhap_ctx_reserve(vhap_ctx* ctx, int32_t B, int32_t H, int32_t W) {
  if (B <= ctx->maxB && H <= ctx->maxH && W <= ctx->maxW && (size_t)B * H * W <= (size_t)ctx->maxB * ctx->maxH * ctx->maxW) return 0;
  CK(cudaSetDevice(ctx->device));
  free_batch(ctx);
  if (B < ctx->maxB) B = ctx->maxB;            // never shrink a dimension: callers alternate between small and large shapes
  if (H < ctx->maxH) H = ctx->maxH;
  if (W < ctx->maxW) W = ctx->maxW;
  size_t V = ctx->V, M = 3 * V, n = (size_t)B * H * W;
  UP(ctx->v_shaped, (const float*)nullptr, B * M); UP(ctx->v_shaped_part, (const float*)nullptr, 8 * B * M); UP(ctx->v_posed, (const float*)nullptr, B * M); UP(ctx->g_vshaped, (const float*)nullptr, (size_t)B * ctx->Mpad);
  UP(ctx->verts, (const f4*)nullptr, B * V); UP(ctx->clip, (const f4*)nullptr, B * V); UP(ctx->vnorm, (const f4*)nullptr, B * V);
  UP(ctx->vnraw, (const f4*)nullptr, B * V); UP(ctx->snap, (const i4*)nullptr, B * V); UP(ctx->ndc, (const float*)nullptr, B * V * 2);
  UP(ctx->g_clip, (const float*)nullptr, B * V * 4); UP(ctx->g_vnorm, (const float*)nullptr, B * V * 4); UP(ctx->g_verts, (const float*)nullptr, B * V * 4);
  UP(ctx->posebuf, (const PoseFwd*)nullptr, (size_t)B); UP(ctx->poses, (const float*)nullptr, (size_t)B * 15);
  UP(ctx->gA, (const float*)nullptr, (size_t)B * 60); UP(ctx->gpf, (const float*)nullptr, (size_t)B * 36); UP(ctx->gJ, (const float*)nullptr, (size_t)B * 15);
  UP(ctx->gbetas, (const float*)nullptr, (size_t)B * ctx->K); UP(ctx->betas, (const float*)nullptr, (size_t)B * ctx->K);
  UP(ctx->cam, (const CamParams*)nullptr, (size_t)B);
  UP(ctx->tri_id, (const int*)nullptr, n); UP(ctx->pre, (const f4*)nullptr, n); UP(ctx->signs, (const uint8_t*)nullptr, n);
  UP(ctx->pool_list, (const int*)nullptr, n); UP(ctx->pool_tri, (const int*)nullptr, n); UP(ctx->aa_code, (const float*)nullptr, n * 2); UP(ctx->pair_list, (const int*)nullptr, n * 2); UP(ctx->grgb, (const f4*)nullptr, n);
  UP(ctx->final_rgba, (const float*)nullptr, n * 4); UP(ctx->plane_albedo, (const f4*)nullptr, n); UP(ctx->plane_normal, (const f4*)nullptr, n);
  UP(ctx->plane_diffuse, (const f4*)nullptr, n);
  int tiles = B * ((H + VH_TILE - 1) / VH_TILE) * ((W + VH_TILE - 1) / VH_TILE);
  UP(ctx->tile_count, (const int*)nullptr, (size_t)tiles); UP(ctx->tile_off, (const int*)nullptr, (size_t)tiles); UP(ctx->tile_cursor, (const int*)nullptr, (size_t)tiles);
  ctx->tile_cap = (int)std::min<size_t>((size_t)B * ctx->F * 16, (size_t)1 << 30);
  UP(ctx->tile_list, (const int*)nullptr, (size_t)ctx->tile_cap + 8);     // +slack: TMA copies round list chunks up to 16 bytes
  int nblk = (int)((n + 255) / 256);
  ctx->pool_nblk = nblk;
  UP(ctx->pool_blk_count, (const int*)nullptr, (size_t)16 * nblk); UP(ctx->pool_blk_off, (const int*)nullptr, (size_t)16 * nblk);
  ctx->n_partials_rows = nblk;
  UP(ctx->partials, (const float*)nullptr, (size_t)nblk * 4 * VH_NPART);        // x4: kernels with smaller CTAs use proportionally more rows
  ctx->maxB = B; ctx->maxH = H; ctx->maxW = W;
  return 0;
}

extern "C" void vhap_ctx_destroy(vhap_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  free_batch(c);
  FREE(c->v_template); FREE(c->S_fwd); FREE(c->S_fwd_pad); FREE(c->S_bwd); FREE(c->posedirs); FREE(c->Jreg); FREE(c->lbs_w); FREE(c->JS); FREE(c->Jt);
  FREE(c->faces); FREE(c->faces_uv); FREE(c->verts_uv); FREE(c->lmk_faces); FREE(c->lmk_bary); FREE(c->adj_opp); FREE(c->fid2cid);
  FREE(c->vf_indptr); FREE(c->vf_faces); FREE(c->lap_indptr); FREE(c->lap_idx); FREE(c->lap_val); FREE(c->lap_y);
  FREE(c->face_flags); FREE(c->vert_flags); FREE(c->w_off); FREE(c->w_off_lap); FREE(c->rigid_indptr); FREE(c->rigid_vids); FREE(c->uvmask_res);
  FREE(c->mips[0]); FREE(c->mips[1]); FREE(c->tex_painted); FREE(c->g_tex); FREE(c->tv_partials); FREE(c->tex_counter); FREE(c->tex_loss); FREE(c->scan_state); FREE(c->scal); FREE(c->acc); FREE(c->maxslot);
  for (int r = 0; r < VH_DP_MAX; ++r) if (c->dp_peers_host[r]) cudaIpcCloseMemHandle(c->dp_peers_host[r]);
  FREE(c->dp_box); FREE(c->dp_epoch); FREE(c->dp_wait); FREE(c->dp_peers_dev); FREE(c->dp_grm_peers_dev); FREE(c->dp_exrm_peers_dev); FREE(c->dp_gband); FREE(c->dp_exband); FREE(c->dp_counter);
  FREE(c->overflow_flag); FREE(c->pool_base); FREE(c->pool_count); FREE(c->dev_lr_scale); FREE(c->dev_step);
  free(c);
}

extern "C" int vhap_set_stage_masks(vhap_ctx* ctx, const uint8_t* face_tex_detach, const uint8_t* vert_aa_detach, const float* w_off,
                                    const float* w_off_lap, const int32_t* rigid_indptr, const int32_t* rigid_vids, int32_t n_rigid,
                                    const uint8_t* uvmask_res) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemset(ctx->tex_loss, 0, 4 * sizeof(float)));          // a new stage starts without texture-regulariser loss values
  if (face_tex_detach) CK(cudaMemcpy(ctx->face_flags, face_tex_detach, ctx->F, cudaMemcpyHostToDevice)); else CK(cudaMemset(ctx->face_flags, 0, ctx->F));
  if (vert_aa_detach) CK(cudaMemcpy(ctx->vert_flags, vert_aa_detach, ctx->V, cudaMemcpyHostToDevice)); else CK(cudaMemset(ctx->vert_flags, 0, ctx->V));
  FREE(ctx->w_off); FREE(ctx->w_off_lap); FREE(ctx->rigid_indptr); FREE(ctx->rigid_vids);
  if (w_off) UP(ctx->w_off, w_off, (size_t)ctx->V);
  if (w_off_lap) UP(ctx->w_off_lap, w_off_lap, (size_t)ctx->V);
  ctx->n_rigid = 0;
  if (rigid_indptr && n_rigid > 0) {
    UP(ctx->rigid_indptr, rigid_indptr, (size_t)n_rigid + 1);
    UP(ctx->rigid_vids, rigid_vids, (size_t)rigid_indptr[n_rigid]);
    ctx->n_rigid = n_rigid;
  }
  if (uvmask_res) { FREE(ctx->uvmask_res); UP(ctx->uvmask_res, uvmask_res, (size_t)ctx->T * ctx->T); }
  return 0;
}

extern "C" int vhap_set_injected_random(vhap_ctx* ctx, const uint8_t* w_bits, const float* u) { ctx->inj_w = w_bits; ctx->inj_u = u; return 0; }
extern "C" int vhap_set_want_planes(vhap_ctx* ctx, int32_t on) { ctx->want_planes = on; return 0; }
// test hook: per-pixel switch of the L1 photometric term ([B,H,W] bytes, IMAGE orientation, 0 = pixel left out); NULL = all pixels
extern "C" int vhap_set_loss_mask(vhap_ctx* ctx, const uint8_t* mask) { ctx->loss_mask = mask; return 0; }
extern "C" float* vhap_tex_grad_ptr(vhap_ctx* ctx) { return ctx->g_tex; }

static int check_batch(vhap_ctx* ctx, const vhap_frame_batch* fb) {
  if (fb->B > ctx->maxB || (size_t)fb->B * fb->H * fb->W > (size_t)ctx->maxB * ctx->maxH * ctx->maxW || fb->H > ctx->maxH || fb->W > ctx->maxW) {
    vh_set_error(ctx, "batch", "exceeds vhap_ctx_reserve() shape");
    return -4;
  }
  if (fb->target_format != 0 && fb->target_format != 1) { vh_set_error(ctx, "batch", "unknown target_format"); return -4; }
  if (fb->geo && (fb->n_geo < 1 || fb->n_geo > fb->B || !fb->geo_timesteps)) { vh_set_error(ctx, "batch", "view sharing: n_geo / geo_timesteps inconsistent"); return -4; }
  ctx->curB = fb->B; ctx->curH = fb->H; ctx->curW = fb->W;
  return 0;
}
#define LAST() do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { vh_set_error(ctx, "kernel launch", cudaGetErrorString(_e)); return -5; } } while (0)

__global__ void k_copy_verts(const f4* __restrict__ v4, float* __restrict__ out, size_t n) { VH_PDL_SYNC();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  f4 v = v4[i]; out[i * 3] = v.x; out[i * 3 + 1] = v.y; out[i * 3 + 2] = v.z;
}
__global__ void k_load_gverts(const float* __restrict__ g3, float* __restrict__ g4, size_t n) { VH_PDL_SYNC();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  g4[i * 4] = g3[i * 3]; g4[i * 4 + 1] = g3[i * 3 + 1]; g4[i * 4 + 2] = g3[i * 3 + 2]; g4[i * 4 + 3] = 0.f;
}
__global__ void k_project_only(const float* __restrict__ verts, const CamParams* __restrict__ cam, int V, int H, int W, float* __restrict__ clip) { VH_PDL_SYNC();
  int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (v >= V) return;
  const float* p = verts + ((size_t)b * V + v) * 3;
  CamParams c = cam[b];
  float x = p[0], y = p[1], z = p[2];
  float cx_ = c.RT[0] * x + c.RT[1] * y + c.RT[2] * z + c.RT[3];
  float cy_ = c.RT[4] * x + c.RT[5] * y + c.RT[6] * z + c.RT[7];
  float cz_ = c.RT[8] * x + c.RT[9] * y + c.RT[10] * z + c.RT[11];
  float p00 = c.fx * 2.f / W, p11 = c.fy * 2.f / H, p02 = (W - 2.f * c.cx) / W, p12 = (H - 2.f * c.cy) / H;
  float p22 = -(10.f + 0.1f) / (10.f - 0.1f), p23 = -2.f * 10.f * 0.1f / (10.f - 0.1f);
  float* o = clip + ((size_t)b * V + v) * 4;
  o[0] = p00 * cx_ + p02 * cz_; o[1] = p11 * cy_ + p12 * cz_; o[2] = p22 * cz_ + p23; o[3] = -cz_;
}

__global__ void __launch_bounds__(256) k_zero_multi(VhZeroSegs z) { VH_PDL_SYNC();
  const int sgm = blockIdx.y;
  if (sgm >= z.n) return;
  const size_t n4 = z.bytes[sgm] >> 2;
  unsigned* p = (unsigned*)z.p[sgm];
  if ((((size_t)p) & 15) == 0) {
    uint4* p16 = (uint4*)p;
    const size_t n16 = n4 >> 2;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) p16[i] = make_uint4(0, 0, 0, 0);
    for (size_t i = (n16 << 2) + (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) p[i] = 0u;
  } else {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) p[i] = 0u;
  }
}
void vh_zero_multi(vhap_ctx* c, const VhZeroSegs& z, cudaStream_t s) {
  size_t mx = 0;
  for (int i = 0; i < z.n; ++i) mx = z.bytes[i] > mx ? z.bytes[i] : mx;
  int nb = (int)((mx / 16 + 255) / 256);
  nb = nb < 1 ? 1 : (nb > 148 * 4 ? 148 * 4 : nb);
  LAUNCH(c, KID_MISC, s, vh_launch(k_zero_multi, dim3(nb, z.n), 256, 0, s, z));
}

// per-step scratch of the backward pass + the loss accumulators, one launch
static void zero_backward_scratch(vhap_ctx* c, int B, cudaStream_t s, bool with_acc = false, int ntiles = 0) {
  size_t V = c->V;
  VhZeroSegs z; z.n = 0;
  auto add = [&](void* p, size_t bytes) { z.p[z.n] = p; z.bytes[z.n] = bytes; ++z.n; };
  add(c->g_verts, B * V * 4 * sizeof(float)); add(c->g_clip, B * V * 4 * sizeof(float)); add(c->g_vnorm, B * V * 4 * sizeof(float));
  add(c->gA, (size_t)B * 60 * sizeof(float)); add(c->gpf, (size_t)B * 36 * sizeof(float)); add(c->gbetas, (size_t)B * c->K * sizeof(float));
  if (with_acc) { add(c->acc, ACC_COUNT * sizeof(float)); add(c->maxslot, sizeof(unsigned long long)); add(c->pair_count, sizeof(int)); }
  if (ntiles) { add(c->tile_count, sizeof(int) * ntiles); add(c->tile_cursor, sizeof(int) * ntiles); }      // the rasteriser's bins (12 segments max)
  vh_zero_multi(c, z, s);
}

extern "C" int vhap_flame_forward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, float* verts, float* verts_cano, float* lmks, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (check_batch(ctx, fb)) return -4;
  launch_cam_setup(ctx, p, fb, s);
  launch_flame_forward(ctx, p, fb, s);
  size_t n = (size_t)fb->B * ctx->V;
  if (verts) LAUNCH(ctx, KID_MISC, s, vh_launch(k_copy_verts, (unsigned)((n + 255) / 256), 256, 0, s, ctx->verts, verts, n));
  if (verts_cano) cudaMemcpyAsync(verts_cano, ctx->v_shaped, n * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s);
  if (lmks) launch_landmarks(ctx, fb, 0.f, 0, lmks, nullptr, 0, 0, fb->B, s);
  LAST();
  return 0;
}

extern "C" int vhap_flame_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const float* g_verts, const float* g_lmks,
                                   const vhap_grads* g, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (check_batch(ctx, fb)) return -4;
  zero_backward_scratch(ctx, fb->B, s);
  size_t n = (size_t)fb->B * ctx->V;
  if (g_verts) LAUNCH(ctx, KID_MISC, s, vh_launch(k_load_gverts, (unsigned)((n + 255) / 256), 256, 0, s, g_verts, ctx->g_verts, n));
  if (g_lmks) launch_landmarks(ctx, fb, 0.f, 0, nullptr, (float*)g_lmks, 0, 0, fb->B, s);
  launch_flame_backward(ctx, p, fb, g, 0, s);
  LAST();
  return 0;
}

extern "C" int vhap_project(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const float* verts, float* verts_clip, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (check_batch(ctx, fb)) return -4;
  launch_cam_setup(ctx, p, fb, s);
  dim3 g((ctx->V + 127) / 128, fb->B);
  LAUNCH(ctx, KID_MISC, s, vh_launch(k_project_only, g, 128, 0, s, verts, ctx->cam, ctx->V, fb->H, fb->W, verts_clip));
  LAST();
  return 0;
}

extern "C" int vhap_rasterize(vhap_ctx* ctx, const float* verts_clip, int32_t B, int32_t H, int32_t W, int32_t* tri_id, float* rast, float* rast_db,
                              int32_t cull_backface, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  vhap_frame_batch fb; memset(&fb, 0, sizeof(fb)); fb.B = B; fb.H = H; fb.W = W;
  if (check_batch(ctx, &fb)) return -4;
  int* ids = tri_id ? tri_id : ctx->tri_id;
  launch_raster(ctx, (const f4*)verts_clip, ctx->snap, B, H, W, ids, cull_backface, 1, s);
  if (rast || rast_db) launch_rast_out(ctx, (const f4*)verts_clip, B, H, W, ids, rast, rast_db, s);
  LAST();
  return 0;
}

extern "C" int vhap_set_tex_painted(vhap_ctx* ctx, const float* tex_painted, void* stream) {
  size_t n = (size_t)3 * ctx->T * ctx->T;
  if (!ctx->tex_painted) CK(cudaMalloc((void**)&ctx->tex_painted, n * sizeof(float)));
  CK(cudaMemcpyAsync(ctx->tex_painted, tex_painted, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

extern "C" int vhap_tex_rebuild(vhap_ctx* ctx, const float* tex_extra, void* stream) {
  launch_tex_rebuild(ctx, tex_extra, (cudaStream_t)stream);
  LAST();
  return 0;
}

__global__ void k_assemble_losses(const float* __restrict__ acc, vhap_stage_cfg cfg, float max_hw, float* __restrict__ g_focal, int add_focal,
                                  float* __restrict__ out, const float* __restrict__ tex_loss = nullptr) { VH_PDL_SYNC();
  if (threadIdx.x != 0) return;
  if (add_focal && g_focal) g_focal[0] += (acc[ACC_GFX] + acc[ACC_GFY]) * max_hw;       // fx = fy = focal * max(h,w) (tracker.py:153)
  if (!out) return;
  for (int i = 0; i < VHAP_N_LOSS; ++i) out[i] = 0.f;
  out[VHAP_L_LMK] = acc[ACC_LMK]; out[VHAP_L_PHOTO] = acc[ACC_PHOTO]; out[VHAP_L_REG_SHAPE] = acc[ACC_REG_SHAPE];
  out[VHAP_L_REG_EXPR] = acc[ACC_REG_EXPR]; out[VHAP_L_REG_JOINT] = acc[ACC_REG_JOINT]; out[VHAP_L_SMOOTH_POSE] = acc[ACC_SMOOTH_POSE];
  out[VHAP_L_SMOOTH_JOINT] = acc[ACC_SMOOTH_JOINT]; out[VHAP_L_SMOOTH_EXPR] = acc[ACC_SMOOTH_EXPR]; out[VHAP_L_REG_TEX_TV] = tex_loss ? tex_loss[0] : 0.f;
  out[VHAP_L_REG_TEX_RES] = tex_loss ? tex_loss[1] : 0.f; out[VHAP_L_REG_DIFFUSE] = acc[ACC_REG_DIFFUSE]; out[VHAP_L_REG_LIGHT] = acc[ACC_REG_LIGHT];
  out[VHAP_L_REG_OFFSET] = acc[ACC_REG_OFFSET]; out[VHAP_L_REG_OFFSET_LAP] = acc[ACC_REG_OFFSET_LAP]; out[VHAP_L_REG_OFFSET_RIGID] = acc[ACC_REG_OFFSET_RIGID];
  out[VHAP_L_NFG] = acc[ACC_NFG]; out[VHAP_L_ABSERR] = acc[ACC_ABSERR];
  float t = 0.f;
  for (int i = VHAP_L_LMK; i <= VHAP_L_REG_OFFSET_RIGID; ++i) t += out[i];
  out[VHAP_L_TOTAL] = t;
}

// forward half: FLAME, landmarks, rasterise, shade, disturb + AA + L1 partial sums; fills reduce_slab (see render.cu)
extern "C" int vhap_energy_forward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg, float* reduce_slab,
                                   void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (check_batch(ctx, fb)) return -4;
  const bool photo = cfg->photometric && cfg->w_photo >= 0.f;
  zero_backward_scratch(ctx, fb->B, s, true, photo ? fb->B * ((fb->H + VH_TILE - 1) / VH_TILE) * ((fb->W + VH_TILE - 1) / VH_TILE) : 0);
  launch_flame_forward(ctx, p, fb, s, true);            // (+ camera set-up)
  if (photo) {
    // fork: vertex normals (needed only by the shading pass) run on aux stream 0 while the rasteriser runs on the main stream
    const bool vn_overlap = !ctx->no_overlap;
    if (vn_overlap) {
      cudaEventRecord(ctx->ev[EV_VN_FORK], s);
      cudaStreamWaitEvent(ctx->aux[0], ctx->ev[EV_VN_FORK], 0);
      launch_vnormals(ctx, fb->geo ? fb->n_geo : fb->B, ctx->aux[0]);
      cudaEventRecord(ctx->ev[EV_VN_DONE], ctx->aux[0]);
    } else launch_vnormals(ctx, fb->geo ? fb->n_geo : fb->B, s);
    launch_raster(ctx, ctx->clip, ctx->snap, fb->B, fb->H, fb->W, ctx->tri_id, 0, 0, s, true);      // (bins cleared with the step's scratch above)
    if (vn_overlap) cudaStreamWaitEvent(s, ctx->ev[EV_VN_DONE], 0);
    PassArgs P;
    fill_render_args(ctx, P, fb, cfg, p->lights);
    launch_render_forward(ctx, P, s, true);               // (maxslot / pair_count were cleared with the step's scratch above)
    launch_forward_slab(ctx, P, p->lights, reduce_slab, s);
  } else {
    cudaMemsetAsync(reduce_slab, 0, 8 * sizeof(float), s);
  }
  LAST();
  return 0;
}

extern "C" int vhap_energy_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg, const float* reduce_slab,
                                    const float* local_slab, int32_t global_B, const vhap_grads* g, float* losses_out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (check_batch(ctx, fb)) return -4;
  bool photo = cfg->photometric && cfg->w_photo >= 0.f;
  vhap_grads none; memset(&none, 0, sizeof(none));
  const vhap_grads* gg = g ? g : &none;
  int opt_cam = (gg->focal_length != nullptr) && (fb->K == nullptr);
  // landmark energy (tracker.py:712-719): mean over global_B * n landmarks
  // fork: the parameter-space regularisers only need the parameters and the (already zeroed) gradient slab -> aux stream 0,
  // concurrent with the render backward (works eagerly and inside CUDA-graph capture: fork/join through events)
  // The landmark energy (tracker.py:712-719) rides on the same side chain, ahead of the regularisers: its vertex gradients are only
  // needed by the skinning adjoint (joined there through EV_LMK_DONE).
  const bool lmk = cfg->w_landmark >= 0.f;
  const bool regs_forked = (cfg->training || lmk) && !ctx->no_overlap;
  cudaStream_t ss = regs_forked ? ctx->aux[0] : s;
  if (regs_forked) { cudaEventRecord(ctx->ev[EV_REGS_FORK], s); cudaStreamWaitEvent(ss, ctx->ev[EV_REGS_FORK], 0); }
  if (lmk) {
    int nl = cfg->jawline_off ? 51 : 68;
    launch_landmarks(ctx, fb, cfg->w_landmark / ((float)global_B * nl), cfg->jawline_off, nullptr, nullptr, 1, opt_cam, global_B, ss);
    if (regs_forked) cudaEventRecord(ctx->ev[EV_LMK_DONE], ss);
  }
  if (cfg->training) launch_regs(ctx, p, fb, cfg, g, global_B, ss);
  if (regs_forked) cudaEventRecord(ctx->ev[EV_REGS_DONE], ss);
  ctx->tex_fork_pending = 0;
  bool lights_joined = true;
  cudaStream_t gs = s;                         // stream of the geometry backward
  if (photo) {
    PassArgs P;
    fill_render_args(ctx, P, fb, cfg, p->lights);
    launch_finalize(ctx, P, cfg, reduce_slab, local_slab, global_B, p->lights, gg->lights, s);
    if (g) {
      P.g_tex = gg->tex_grad_pyramid;
      const bool lights_side = regs_forked && gg->lights != nullptr;
      launch_render_backward(ctx, P, cfg, p->lights, gg->lights, nullptr, s, lights_side ? ctx->aux[0] : nullptr);
      if (!lights_side) cudaEventRecord(ctx->ev[EV_TEXGRAD_READY], s);   // texel gradients complete: the texture update may start (see vhap_tex_reg_fold_adam)
      lights_joined = !lights_side;
      ctx->tex_fork_pending = 1;
      // the geometry backward is a chain of small latency-bound kernels; the texture update that runs concurrently is one
      // machine-filling streaming kernel.  Put the chain on a highest-priority stream so its CTAs are scheduled ahead of the
      // bulk kernel's remaining CTAs instead of queueing behind all of them.
      if (!ctx->no_overlap) { gs = ctx->hp[0]; cudaStreamWaitEvent(gs, ctx->ev[EV_TEXGRAD_READY], 0); }
      launch_vnormals_bwd(ctx, fb->geo ? fb->n_geo : fb->B, gs);
    }
  }
  if (g && lmk && regs_forked) cudaStreamWaitEvent(gs, ctx->ev[EV_LMK_DONE], 0);      // the skinning adjoint reads the landmark term's vertex gradients
  if (g) launch_flame_backward(ctx, p, fb, gg, opt_cam, gs);
  if (gs != s) { cudaEventRecord(ctx->ev[EV_GEOM_DONE], gs); cudaStreamWaitEvent(s, ctx->ev[EV_GEOM_DONE], 0); }
  if (regs_forked) cudaStreamWaitEvent(s, ctx->ev[EV_REGS_DONE], 0);        // join the side chain (landmarks, regularisers)
  if (!lights_joined) cudaStreamWaitEvent(s, ctx->ev[EV_LIGHTS_DONE], 0);   // ... and the light-gradient reduction
  float max_hw = (float)(fb->H > fb->W ? fb->H : fb->W);
  LAUNCH(ctx, KID_ASSEMBLE, s, vh_launch(k_assemble_losses, 1, 32, 0, s, ctx->acc, *cfg, max_hw, gg->focal_length, opt_cam, losses_out, ctx->tex_loss));
  LAST();
  return 0;
}

extern "C" int vhap_energy_forward_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg,
                                            const vhap_grads* g, float* losses_out, void* stream) {
  float* slab = ctx->scal + 8;     // scal[8..15] doubles as the single-GPU reduce slab
  int r = vhap_energy_forward(ctx, p, fb, cfg, slab, stream);
  if (r) return r;
  return vhap_energy_backward(ctx, p, fb, cfg, slab, slab, fb->B, g, losses_out, stream);
}

extern "C" int vhap_get_plane(vhap_ctx* ctx, int32_t which, float* out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int B = ctx->curB, H = ctx->curH, W = ctx->curW;
  const float* src = nullptr;
  switch (which) {
    case 0: src = ctx->final_rgba; break;
    case 1: src = (const float*)ctx->pre; break;
    case 2: src = (const float*)ctx->plane_albedo; break;
    case 3: src = (const float*)ctx->plane_normal; break;
    case 4: src = (const float*)ctx->plane_diffuse; break;
    case 5: launch_cid_plane(ctx, out, s); LAST(); return 0;
    default: vh_set_error(ctx, "vhap_get_plane", "unknown plane"); return -3;
  }
  launch_flip_plane(ctx, src, out, B, H, W, s);
  LAST();
  return 0;
}

extern "C" int vhap_tex_reg_fold_adam(vhap_ctx* ctx, float* tex_extra, float* g_out, float* adam_m, float* adam_v, float lr, int32_t step,
                                      const vhap_stage_cfg* cfg, float photo_scale, float* losses_out, void* stream) {
  (void)photo_scale;
  cudaStream_t s = (cudaStream_t)stream;
  if (ctx->tex_fork_pending && (g_out == nullptr || ctx->tex_gout_persistent) && !ctx->no_overlap) {
    // (with a caller-owned dense output the call stays on the caller's stream -- ordering w.r.t. the caller's own work on g_out --
    // unless the caller declared the buffer persistent and untouched outside this library's calls, vhap_set_tex_grad_persistent)
    // the texture fold / Adam / mip rebuild only depends on the texel gradients (event 2, recorded right after the fused
    // backward): run it on aux stream 1 concurrently with the geometry backward that was enqueued after that event, then join
    ctx->tex_fork_pending = 0;
    cudaStreamWaitEvent(ctx->aux[1], ctx->ev[EV_TEXGRAD_READY], 0);
    launch_tex_fold(ctx, tex_extra, g_out, adam_m, adam_v, lr, step, cfg, losses_out, ctx->aux[1]);
    cudaEventRecord(ctx->ev[EV_TEX_DONE], ctx->aux[1]);
    cudaStreamWaitEvent(s, ctx->ev[EV_TEX_DONE], 0);
  } else {
    ctx->tex_fork_pending = 0;
    launch_tex_fold(ctx, tex_extra, g_out, adam_m, adam_v, lr, step, cfg, losses_out, s);
  }
  if (losses_out) {
    // add the two texture terms to the loss vector produced by vhap_energy_backward
    LAUNCH(ctx, KID_ASSEMBLE, (cudaStream_t)stream, vh_launch(k_assemble_losses, 1, 32, 0, (cudaStream_t)stream, ctx->acc, *cfg, 0.f, nullptr, 0, losses_out, ctx->tex_loss));
  }
  LAST();
  return 0;
}

// Data-parallel texture update: Adam from a dense, already folded + regularised + all-reduced gradient (the g_out of
// vhap_tex_reg_fold_adam), then level 0 / level 1 of the other pyramid in the same pass and the remaining mip levels.
extern "C" int vhap_tex_apply_grad(vhap_ctx* ctx, float* tex_extra, const float* g_dense, float* adam_m, float* adam_v, float lr, int32_t step,
                                   const vhap_stage_cfg* cfg, void* stream) {
  ctx->tex_apply_grad = g_dense;
  launch_tex_fold(ctx, tex_extra, nullptr, adam_m, adam_v, lr, step, cfg, nullptr, (cudaStream_t)stream);
  ctx->tex_apply_grad = nullptr;
  LAST();
  return 0;
}
// ---- data-parallel peer exchange of the forward slab (render.cu k_forward_slab / k_finalize): CUDA IPC mailboxes, one per rank, mapped by
// every rank of the node.  vhap_dp_init allocates this rank's mailbox and returns its IPC handle (64 bytes); the caller all-gathers the
// handles (any transport) and hands all of them to vhap_dp_connect.  From then on vhap_energy_forward / _backward exchange the batch-global
// scalars themselves (the reduce_slab arguments are ignored) -- no collective call and no host glue between the two halves of a step.
extern "C" int vhap_dp_init(vhap_ctx* ctx, int32_t rank, int32_t world, unsigned char* handle_out_host /*64*/) {
  if (world < 1 || world > VH_DP_MAX || rank < 0 || rank >= world) { vh_set_error(ctx, "vhap_dp_init", "rank / world out of range (max 16 ranks)"); return -3; }
  CK(cudaSetDevice(ctx->device));
  if (!ctx->dp_box) {
    CK(cudaMalloc((void**)&ctx->dp_box, VH_DP_BOX_FLOATS * sizeof(float)));
    CK(cudaMalloc((void**)&ctx->dp_epoch, 4 * sizeof(int)));       // [0] slab epoch, [1] error flag, [2] [3] epochs of the texture barriers A / B
    CK(cudaMalloc((void**)&ctx->dp_wait, 4 * sizeof(unsigned long long))); CK(cudaMemset(ctx->dp_wait, 0, 4 * sizeof(unsigned long long)));
    ctx->dp_err = ctx->dp_epoch + 1;
  }
  CK(cudaMemset(ctx->dp_box, 0, VH_DP_BOX_FLOATS * sizeof(float)));
  { int init[4] = {1, 0, 0, 0}; CK(cudaMemcpy(ctx->dp_epoch, init, sizeof(init), cudaMemcpyHostToDevice)); }
  ctx->dp_rank = rank; ctx->dp_world = world;
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->dp_box));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out_host, &h, 64);
  return 0;
}
extern "C" int vhap_dp_connect(vhap_ctx* ctx, const unsigned char* handles_host /*[world][64]*/) {
  if (!ctx->dp_box) { vh_set_error(ctx, "vhap_dp_connect", "call vhap_dp_init first"); return -3; }
  CK(cudaSetDevice(ctx->device));
  float* table[VH_DP_MAX] = {nullptr};
  for (int r = 0; r < ctx->dp_world; ++r) {
    if (r == ctx->dp_rank) { table[r] = ctx->dp_box; continue; }
    cudaIpcMemHandle_t h; memcpy(&h, handles_host + (size_t)r * 64, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->dp_peers_host[r] = p; table[r] = (float*)p;
  }
  if (!ctx->dp_peers_dev) CK(cudaMalloc((void**)&ctx->dp_peers_dev, VH_DP_MAX * sizeof(float*)));
  CK(cudaMemcpy(ctx->dp_peers_dev, table, sizeof(table), cudaMemcpyHostToDevice));
  return 0;
}
// Peer-memory texture update (dp_tex.cu).  The caller allocates g_rm and ex_rm ([T][3][T] floats each) as SYMMETRIC memory on every rank
// (torch.distributed._symmetric_memory) and passes: the table of every rank's device pointer to it (HOST arrays of `world` pointers, own entry
// included) and the multicast (NVLS) address of the buffer, or NULL when the fabric has no multicast -- then the kernels loop over the peers.
extern "C" int vhap_dp_tex_connect(vhap_ctx* ctx, void* const* grm_ptrs_host, void* grm_multicast, void* const* exrm_ptrs_host, void* exrm_multicast) {
  if (!ctx->dp_peers_dev || ctx->dp_world < 2) { vh_set_error(ctx, "vhap_dp_tex_connect", "call vhap_dp_init / vhap_dp_connect first (world >= 2)"); return -3; }
  if (ctx->T % (ctx->dp_world * 8)) { vh_set_error(ctx, "vhap_dp_tex_connect", "texture size must be a multiple of 8 * world"); return -3; }
  CK(cudaSetDevice(ctx->device));
  const size_t nb = (size_t)3 * ctx->T * ctx->T / ctx->dp_world;
  if (!ctx->dp_grm_peers_dev) {
    CK(cudaMalloc((void**)&ctx->dp_grm_peers_dev, VH_DP_MAX * sizeof(float*)));
    CK(cudaMalloc((void**)&ctx->dp_exrm_peers_dev, VH_DP_MAX * sizeof(float*)));
    CK(cudaMalloc((void**)&ctx->dp_counter, sizeof(unsigned))); CK(cudaMemset(ctx->dp_counter, 0, sizeof(unsigned)));
    CK(cudaMalloc((void**)&ctx->dp_gband, nb * sizeof(float)));
    CK(cudaMalloc((void**)&ctx->dp_exband, nb * sizeof(float)));
  }
  CK(cudaMemcpy(ctx->dp_grm_peers_dev, grm_ptrs_host, ctx->dp_world * sizeof(void*), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->dp_exrm_peers_dev, exrm_ptrs_host, ctx->dp_world * sizeof(void*), cudaMemcpyHostToDevice));
  ctx->dp_grm = (float*)grm_ptrs_host[ctx->dp_rank]; ctx->dp_exrm = (float*)exrm_ptrs_host[ctx->dp_rank];
  ctx->dp_grm_mc = (float*)grm_multicast; ctx->dp_exrm_mc = (float*)exrm_multicast;
  return 0;
}
// The chain of dp_tex.cu in two parts, so that a pipelined step can run the first beside its own geometry backward and the second at the
// start of the next step (Engine._step_body):
//   part 1: fold -> barrier A -> in-switch reduction of this rank's band.  Right after vhap_energy_backward (texel gradients just completed)
//           it forks onto the library's bulk stream behind the fused backward; vhap_dp_tex_join makes `stream` wait for it.
//   part 2: regularisers + Adam on the band -> multicast store -> barrier B -> pyramid rebuild.
extern "C" int vhap_dp_tex_part1(vhap_ctx* ctx, float* tex_extra, void* stream) {
  if (!ctx->dp_grm) { vh_set_error(ctx, "vhap_dp_tex_part1", "call vhap_dp_tex_connect first"); return -3; }
  cudaStream_t s = (cudaStream_t)stream;
  ctx->dp_tex_forked = 0;
  if (ctx->tex_fork_pending && !ctx->no_overlap) {
    cudaStreamWaitEvent(ctx->aux[1], ctx->ev[EV_TEXGRAD_READY], 0);
    s = ctx->aux[1];
    ctx->dp_tex_forked = 1;
  }
  ctx->tex_fork_pending = 0;
  launch_tex_fold_grad_rm(ctx, tex_extra, ctx->dp_grm, s);
  launch_dp_barrier(ctx, 0, s);
  launch_dp_reduce_band(ctx, ctx->dp_gband, s);
  if (ctx->dp_tex_forked) cudaEventRecord(ctx->ev[EV_TEX_DONE], s);
  LAST();
  return 0;
}
extern "C" int vhap_dp_tex_join(vhap_ctx* ctx, void* stream) {
  if (ctx->dp_tex_forked) { cudaStreamWaitEvent((cudaStream_t)stream, ctx->ev[EV_TEX_DONE], 0); ctx->dp_tex_forked = 0; }
  return 0;
}
extern "C" int vhap_dp_tex_part2(vhap_ctx* ctx, float* tex_extra, float* adam_m, float* adam_v, float lr, int32_t step, const vhap_stage_cfg* cfg, void* stream) {
  if (!ctx->dp_grm) { vh_set_error(ctx, "vhap_dp_tex_part2", "call vhap_dp_tex_connect first"); return -3; }
  cudaStream_t s = (cudaStream_t)stream;
  const int rows = ctx->T / ctx->dp_world, y0 = ctx->dp_rank * rows;
  if (launch_tex_band_adam(ctx, tex_extra, ctx->dp_gband, y0, y0 + rows, adam_m, adam_v, lr, step, cfg, ctx->dp_exband, s)) { vh_set_error(ctx, "vhap_dp_tex_part2", "bad band"); return -3; }
  launch_dp_bcast_band(ctx, ctx->dp_exband, s);
  launch_dp_barrier(ctx, 1, s);
  launch_tex_rebuild_rm(ctx, tex_extra, ctx->dp_exrm, s);
  LAST();
  return 0;
}
// one call = both parts on `stream`
extern "C" int vhap_dp_tex_update(vhap_ctx* ctx, float* tex_extra, float* adam_m, float* adam_v, float lr, int32_t step, const vhap_stage_cfg* cfg, void* stream) {
  ctx->tex_fork_pending = 0;
  int r = vhap_dp_tex_part1(ctx, tex_extra, stream);
  if (r) return r;
  return vhap_dp_tex_part2(ctx, tex_extra, adam_m, adam_v, lr, step, cfg, stream);
}
// 0 = fine, 1 = a peer's flag did not arrive within the spin budget (synchronises)
extern "C" int vhap_dp_status(vhap_ctx* ctx, int32_t* out_host) {
  *out_host = 0;
  if (ctx->dp_err) CK(cudaMemcpy(out_host, ctx->dp_err, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

// time this rank spent waiting for its peers since the last reset: out[0] ns in the mid-step slab exchange (k_finalize), out[1] / out[2] ns
// in the texture barriers A / B, out[3] number of slab exchanges.  Synchronises the device.
extern "C" int vhap_dp_wait_stats(vhap_ctx* ctx, uint64_t* out_host, int32_t reset) {
  for (int i = 0; i < 4; ++i) out_host[i] = 0;
  if (!ctx->dp_wait) return 0;
  CK(cudaMemcpy(out_host, ctx->dp_wait, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) CK(cudaMemset(ctx->dp_wait, 0, 4 * sizeof(unsigned long long)));
  return 0;
}

// ---- sharded texture update for data-parallel runs (texture.cu "sharded texture update"): the three compute pieces around the caller's
// reduce-scatter / all-gather.  All run on the stream they are given (no internal fork).
extern "C" int vhap_tex_fold_grad_rm(vhap_ctx* ctx, float* tex_extra, float* g_rm, void* stream) {
  ctx->tex_fork_pending = 0;
  launch_tex_fold_grad_rm(ctx, tex_extra, g_rm, (cudaStream_t)stream);
  LAST();
  return 0;
}
extern "C" int vhap_tex_band_adam(vhap_ctx* ctx, float* tex_extra, const float* g_band, int32_t y_begin, int32_t y_end, float* adam_m, float* adam_v,
                                  float lr, int32_t step, const vhap_stage_cfg* cfg, float* ex_band_out, void* stream) {
  if (launch_tex_band_adam(ctx, tex_extra, g_band, y_begin, y_end, adam_m, adam_v, lr, step, cfg, ex_band_out, (cudaStream_t)stream)) {
    vh_set_error(ctx, "vhap_tex_band_adam", "row band must be a non-empty multiple of 8 rows inside the texture"); return -3;
  }
  LAST();
  return 0;
}
extern "C" int vhap_tex_rebuild_rm(vhap_ctx* ctx, float* tex_extra, const float* ex_rm, void* stream) {
  launch_tex_rebuild_rm(ctx, tex_extra, ex_rm, (cudaStream_t)stream);
  LAST();
  return 0;
}
// Deferred texture update (the update of step k runs at the start of step k+1, beside FLAME / rasteriser / pools, and is joined right
// before the shading pass): cancel the in-call fork of the next vhap_tex_reg_fold_adam (it then runs on the stream it is given) and
// bias the device Adam step it reads (-1: the counter was already advanced).  bias 0 restores the normal behaviour.
extern "C" int vhap_tex_defer(vhap_ctx* ctx, int32_t step_bias) { ctx->tex_fork_pending = 0; ctx->tex_step_bias = step_bias; return 0; }
// loss VALUES of the texture regularisers for the texture currently in use (complete loss vectors with a deferred update)
extern "C" int vhap_tex_reg_loss(vhap_ctx* ctx, const float* tex_extra, const vhap_stage_cfg* cfg, void* stream) {
  launch_tex_reg_loss(ctx, tex_extra, cfg, (cudaStream_t)stream);
  LAST();
  return 0;
}
// re-assemble the loss vector of the step that just ran (accumulators + the current texture-regulariser loss values)
extern "C" int vhap_assemble_losses(vhap_ctx* ctx, const vhap_stage_cfg* cfg, float* losses_out, void* stream) {
  LAUNCH(ctx, KID_ASSEMBLE, (cudaStream_t)stream, vh_launch(k_assemble_losses, 1, 32, 0, (cudaStream_t)stream, ctx->acc, *cfg, 0.f, nullptr, 0, losses_out, ctx->tex_loss));
  LAST();
  return 0;
}
// one-shot: the next render forward waits for `event` (a cudaEvent_t recorded by the caller) after the pixel pools are built and
// before the first kernel that reads the texture
extern "C" int vhap_set_render_wait_event(vhap_ctx* ctx, void* event) { ctx->render_wait_ev = (cudaEvent_t)event; return 0; }

// on != 0: the g_out buffer passed to vhap_tex_reg_fold_adam is persistent and only read by the caller AFTER that call returns
// (on the same stream): the fold may then run on the aux stream beside the geometry backward, like the fused single-GPU update
extern "C" int vhap_set_tex_grad_persistent(vhap_ctx* ctx, int32_t on) { ctx->tex_gout_persistent = on; return 0; }

extern "C" int vhap_adam(vhap_ctx* ctx, float* param, const float* grad, float* m, float* v, int64_t n, float lr, int32_t step, void* stream) {
  launch_adam(ctx, param, grad, m, v, n, lr, step, (cudaStream_t)stream);
  LAST();
  return 0;
}

// ---------------------------------------------------------------------------------------------- modular entry points
__global__ void k_3to4(const float* __restrict__ a, float* __restrict__ b, size_t n, float w) { VH_PDL_SYNC();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  b[i * 4] = a[i * 3]; b[i * 4 + 1] = a[i * 3 + 1]; b[i * 4 + 2] = a[i * 3 + 2]; b[i * 4 + 3] = w;
}
__global__ void k_4to3_add(const float* __restrict__ a, float* __restrict__ b, size_t n) { VH_PDL_SYNC();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  b[i * 3] += a[i * 4]; b[i * 3 + 1] += a[i * 4 + 1]; b[i * 3 + 2] += a[i * 4 + 2];
}
__global__ void k_4to3(const float* __restrict__ a, float* __restrict__ b, size_t n) { VH_PDL_SYNC();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  b[i * 3] = a[i * 4]; b[i * 3 + 1] = a[i * 4 + 1]; b[i * 3 + 2] = a[i * 4 + 2];
}
#define GRID1(n) (unsigned)(((n) + 255) / 256), 256

// compute_v_normals (render_nvdiffrast.py:297-316)
extern "C" int vhap_vertex_normals(vhap_ctx* ctx, const float* verts, int32_t B, float* vnorm, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  vhap_frame_batch fb; memset(&fb, 0, sizeof(fb)); fb.B = B; fb.H = ctx->maxH; fb.W = ctx->maxW;
  if (check_batch(ctx, &fb)) return -4;
  size_t n = (size_t)B * ctx->V;
  LAUNCH(ctx, KID_MISC, s, vh_launch(k_3to4, GRID1(n), 0, s, verts, (float*)ctx->verts, n, 1.f));
  launch_vnormals(ctx, B, s);
  LAUNCH(ctx, KID_MISC, s, vh_launch(k_4to3, GRID1(n), 0, s, (const float*)ctx->vnorm, vnorm, n));
  LAST();
  return 0;
}
// adjoint: g_verts += d(vnorm)/d(verts)^T g_vnorm
extern "C" int vhap_vertex_normals_backward(vhap_ctx* ctx, const float* verts, const float* g_vnorm, int32_t B, float* g_verts, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  size_t n = (size_t)B * ctx->V;
  LAUNCH(ctx, KID_MISC, s, vh_launch(k_3to4, GRID1(n), 0, s, verts, (float*)ctx->verts, n, 1.f));
  launch_vnormals(ctx, B, s);
  LAUNCH(ctx, KID_MISC, s, vh_launch(k_3to4, GRID1(n), 0, s, g_vnorm, ctx->g_vnorm, n, 0.f));
  cudaMemsetAsync(ctx->g_verts, 0, n * 4 * sizeof(float), s);
  launch_vnormals_bwd(ctx, B, s);
  LAUNCH(ctx, KID_MISC, s, vh_launch(k_4to3_add, GRID1(n), 0, s, ctx->g_verts, g_verts, n));
  LAST();
  return 0;
}

__global__ void k_project_bwd(const float* __restrict__ verts, const float* __restrict__ g_clip, const CamParams* __restrict__ cam, int V, int H, int W,
                              float* __restrict__ g_verts, float* __restrict__ g_fxfy /* [2] accumulators */) { VH_PDL_SYNC();
  int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (v >= V) return;
  const float* p = verts + ((size_t)b * V + v) * 3;
  const float* g = g_clip + ((size_t)b * V + v) * 4;
  CamParams c = cam[b];
  float p00 = c.fx * 2.f / W, p11 = c.fy * 2.f / H, p02 = (W - 2.f * c.cx) / W, p12 = (H - 2.f * c.cy) / H, p22 = -(10.f + 0.1f) / (10.f - 0.1f);
  float cx_ = c.RT[0] * p[0] + c.RT[1] * p[1] + c.RT[2] * p[2] + c.RT[3];
  float cy_ = c.RT[4] * p[0] + c.RT[5] * p[1] + c.RT[6] * p[2] + c.RT[7];
  float g_cx = p00 * g[0], g_cy = p11 * g[1], g_cz = p02 * g[0] + p12 * g[1] + p22 * g[2] - g[3];
  float* o = g_verts + ((size_t)b * V + v) * 3;
  o[0] += c.RT[0] * g_cx + c.RT[4] * g_cy + c.RT[8] * g_cz;
  o[1] += c.RT[1] * g_cx + c.RT[5] * g_cy + c.RT[9] * g_cz;
  o[2] += c.RT[2] * g_cx + c.RT[6] * g_cy + c.RT[10] * g_cz;
  if (g_fxfy) { atomicAdd(g_fxfy, g[0] * cx_ * (2.f / W)); atomicAdd(g_fxfy + 1, g[1] * cy_ * (2.f / H)); }
}
// adjoint of vhap_project: g_verts += ..., g_focal += ... (uncalibrated camera, tracker.py:141-157)
extern "C" int vhap_project_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const float* verts, const float* g_clip,
                                     float* g_verts, float* g_focal, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (check_batch(ctx, fb)) return -4;
  launch_cam_setup(ctx, p, fb, s);
  cudaMemsetAsync(ctx->acc, 0, ACC_COUNT * sizeof(float), s);
  dim3 g((ctx->V + 127) / 128, fb->B);
  LAUNCH(ctx, KID_MISC, s, vh_launch(k_project_bwd, g, 128, 0, s, verts, g_clip, ctx->cam, ctx->V, fb->H, fb->W, g_verts, g_focal ? ctx->acc + ACC_GFX : nullptr));
  if (g_focal) {
    vhap_stage_cfg dummy; memset(&dummy, 0, sizeof(dummy));
    LAUNCH(ctx, KID_ASSEMBLE, s, vh_launch(k_assemble_losses, 1, 32, 0, s, ctx->acc, dummy, (float)(fb->H > fb->W ? fb->H : fb->W), g_focal, 1, nullptr, nullptr));
  }
  LAST();
  return 0;
}

// rasterise + render_rgba + photometric energy (+ reg_diffuse) from caller-provided clip positions and vertex normals, with the
// analytic backward to them: the modular form of NVDiffRenderer.rasterize/render_rgba (render_nvdiffrast.py:216-245,354-484)
// + compute_photometric_energy (tracker.py:391-478).  g_* outputs are overwritten (not accumulated); any may be NULL.
extern "C" int vhap_render_photometric(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg,
                                       const float* verts_clip, const float* vnorm, float* losses_out, float* g_clip, float* g_vnorm,
                                       float* g_lights, float* g_tex_pyramid, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (check_batch(ctx, fb)) return -4;
  size_t n = (size_t)fb->B * ctx->V;
  cudaMemsetAsync(ctx->acc, 0, ACC_COUNT * sizeof(float), s);
  zero_backward_scratch(ctx, fb->B, s);
  cudaMemcpyAsync(ctx->clip, verts_clip, n * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s);
  LAUNCH(ctx, KID_MISC, s, vh_launch(k_3to4, GRID1(n), 0, s, vnorm, (float*)ctx->vnorm, n, 0.f));
  launch_raster(ctx, ctx->clip, ctx->snap, fb->B, fb->H, fb->W, ctx->tri_id, 0, 1, s);
  PassArgs P;
  fill_render_args(ctx, P, fb, cfg, p->lights);
  launch_render_forward(ctx, P, s);
  float* slab = ctx->scal + 8;
  launch_forward_slab(ctx, P, p->lights, slab, s);
  if (g_lights) cudaMemsetAsync(g_lights, 0, 27 * sizeof(float), s);
  launch_finalize(ctx, P, cfg, slab, slab, fb->B, p->lights, g_lights, s);
  if (g_clip || g_vnorm || g_lights || g_tex_pyramid) {
    P.g_tex = g_tex_pyramid;
    launch_render_backward(ctx, P, cfg, p->lights, g_lights, nullptr, s);
    if (g_clip) cudaMemcpyAsync(g_clip, ctx->g_clip, n * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s);
    if (g_vnorm) LAUNCH(ctx, KID_MISC, s, vh_launch(k_4to3, GRID1(n), 0, s, ctx->g_vnorm, g_vnorm, n));
  }
  LAUNCH(ctx, KID_ASSEMBLE, s, vh_launch(k_assemble_losses, 1, 32, 0, s, ctx->acc, *cfg, 0.f, nullptr, 0, losses_out, ctx->tex_loss));
  LAST();
  return 0;
}

// backward of the last vhap_render_photometric forward for an EXTERNAL upstream gradient d L / d rgba_aa ([B,H,W,4], image
// orientation, e.g. from autograd): the adjoint of NVDiffRenderer.render_rgba's 'rgba' output (render_nvdiffrast.py:476-483)
extern "C" int vhap_render_rgba_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg,
                                         const float* g_rgba, float* g_clip, float* g_vnorm, float* g_lights, float* g_tex_pyramid, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (check_batch(ctx, fb)) return -4;
  size_t n = (size_t)fb->B * ctx->V;
  zero_backward_scratch(ctx, fb->B, s);
  launch_flip_plane(ctx, g_rgba, ctx->final_rgba, fb->B, fb->H, fb->W, s);      // image -> raster orientation (scratch plane)
  cudaMemsetAsync(ctx->scal, 0, 8 * sizeof(float), s);                           // no photometric scale / reg_diffuse terms
  PassArgs P;
  fill_render_args(ctx, P, fb, cfg, p->lights);
  P.g_tex = g_tex_pyramid;
  if (g_lights) cudaMemsetAsync(g_lights, 0, 27 * sizeof(float), s);
  launch_render_backward(ctx, P, cfg, p->lights, g_lights, ctx->final_rgba, s);
  if (g_clip) cudaMemcpyAsync(g_clip, ctx->g_clip, n * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s);
  if (g_vnorm) vh_launch(k_4to3, GRID1(n), 0, s, ctx->g_vnorm, g_vnorm, n);
  LAST();
  return 0;
}

// ---- device-resident step counters: make a whole optimisation step replayable as a CUDA graph (no host-varying kernel arguments)
__global__ void k_step_set(int* d, int adam_step, int global_step) { VH_PDL_SYNC(); d[0] = adam_step; d[1] = global_step; }
__global__ void k_step_advance(int* d) { VH_PDL_SYNC(); d[0] += 1; d[1] += 1; }
// on != 0: kernels read the Adam step ([0], 1-based) and the global RNG step ([1]) from device memory instead of their arguments
extern "C" int vhap_step_counters(vhap_ctx* ctx, int32_t on, int32_t adam_step, int32_t global_step, void* stream) {
  ctx->use_dev_step = on;
  LAUNCH(ctx, KID_MISC, (cudaStream_t)stream, vh_launch(k_step_set, 1, 1, 0, (cudaStream_t)stream, ctx->dev_step, adam_step, global_step));
  LAST();
  return 0;
}
__global__ void k_set_float(float* d, float v) { VH_PDL_SYNC(); d[0] = v; }
// learning-rate scale of the captured Adam kernels (torch's ExponentialLR between epochs, tracker.py:1407-1412): with device step
// counters on, every Adam kernel multiplies its learning rate by this device scalar, so a captured step graph follows the schedule
extern "C" int vhap_set_lr_scale(vhap_ctx* ctx, float scale, void* stream) {
  LAUNCH(ctx, KID_MISC, (cudaStream_t)stream, vh_launch(k_set_float, 1, 1, 0, (cudaStream_t)stream, ctx->dev_lr_scale, scale));
  LAST();
  return 0;
}
extern "C" int vhap_step_advance(vhap_ctx* ctx, void* stream) {
  LAUNCH(ctx, KID_MISC, (cudaStream_t)stream, vh_launch(k_step_advance, 1, 1, 0, (cudaStream_t)stream, ctx->dev_step));
  LAST();
  return 0;
}
// texture pyramid ping-pong index (0/1): which level-0 buffer is current; needed when graphs captured per parity are replayed
extern "C" int vhap_get_cur_mip(vhap_ctx* ctx) { return ctx->cur_mip; }
extern "C" int vhap_set_cur_mip(vhap_ctx* ctx, int32_t v) { ctx->cur_mip = v & 1; return 0; }

// copies the engine's own geometry of the last forward: which = 0 world vertices [B,V,4], 1 clip positions [B,V,4], 2 vertex normals [B,V,4]
extern "C" int vhap_get_geometry(vhap_ctx* ctx, int32_t which, float* out, void* stream) {
  const void* src = which == 0 ? (const void*)ctx->verts : (which == 1 ? (const void*)ctx->clip : (const void*)ctx->vnorm);
  CK(cudaMemcpyAsync(out, src, (size_t)ctx->curB * ctx->V * 4 * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

void launch_adam_multi(vhap_ctx* c, float* p, const float* g, float* m, float* v, int n_seg, const int64_t* off, const int64_t* len, const float* lr,
                       int step, cudaStream_t s);
// one launch for all parameter groups of a slab: segment k covers [off[k], off[k]+len[k]) with learning rate lr[k] (HOST arrays)
extern "C" int vhap_adam_multi(vhap_ctx* ctx, float* param, const float* grad, float* m, float* v, int32_t n_seg, const int64_t* off_host,
                               const int64_t* len_host, const float* lr_host, int32_t step, void* stream) {
  launch_adam_multi(ctx, param, grad, m, v, n_seg, off_host, len_host, lr_host, step, (cudaStream_t)stream);
  LAST();
  return 0;
}

extern "C" int vhap_overflow_flag(vhap_ctx* ctx, int32_t* out_host) {
  CK(cudaMemcpy(out_host, ctx->overflow_flag, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}
