// FLAME forward/backward kernels: blend shapes (the one dense contraction), pose-corrective + linear-blend skinning fused
// with translation, camera projection and raster snapping; landmark energy; vertex normals; small regularisers.
// Reference path being replaced: FlameHead.forward (vhap/model/flame.py:571-646), vhap/model/lbs.py, NVDiffRenderer
// world_to_camera / camera_to_clip (vhap/util/render_nvdiffrast.py:162-197), compute_v_normals (:297-316),
// compute_lmk_energy (vhap/model/tracker.py:347-389), compute_regularization_energy (:480-605) and their autograd.
#include "engine.h"
#include "accum.h"

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in thread 0 (and broadcast through smem to all)
__device__ float block_sum(float v, float* sh /*[33]*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? sh[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// ------------------------------------------------------------------------------------------------ camera
__device__ __forceinline__ CamParams cam_from(const float* RT, const float* K, const float* focal, int b, int H, int W) {
  CamParams c;
  if (RT) for (int i = 0; i < 12; ++i) c.RT[i] = RT[b * 12 + i];
  else { for (int i = 0; i < 12; ++i) c.RT[i] = 0.f; c.RT[0] = c.RT[5] = c.RT[10] = 1.f; c.RT[11] = -1.f; }   // tracker.py:1335-1337
  if (K) { c.fx = K[b * 4]; c.fy = K[b * 4 + 1]; c.cx = K[b * 4 + 2]; c.cy = K[b * 4 + 3]; }
  else { float f = focal[0] * (float)max(H, W); c.fx = f; c.fy = f; c.cx = 0.5f * W; c.cy = 0.5f * H; }       // tracker.py:141-157
  return c;
}
__global__ void k_cam_setup(CamParams* cam, const float* RT, const float* K, const float* focal, int B, int H, int W) { VH_PDL_SYNC();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) cam[b] = cam_from(RT, K, focal, b, H, W);
}

void launch_cam_setup(vhap_ctx* c, const vhap_params* p, const vhap_frame_batch* fb, cudaStream_t s) {
  LAUNCH(c, KID_CAM, s, vh_launch(k_cam_setup, (fb->B + 63) / 64, 64, 0, s, c->cam, fb->RT, fb->K, p->focal_length, fb->B, fb->H, fb->W));
}

#define VH_NEAR 0.1f
#define VH_FAR 10.0f
struct Proj { float p00, p11, p02, p12, p22, p23; };
__device__ __forceinline__ Proj make_proj(const CamParams& c, int H, int W) {   // render_nvdiffrast.py:117-160
  Proj P;
  P.p00 = c.fx * 2.f / W; P.p11 = c.fy * 2.f / H;
  P.p02 = (W - 2.f * c.cx) / W; P.p12 = (H - 2.f * c.cy) / H;
  P.p22 = -(VH_FAR + VH_NEAR) / (VH_FAR - VH_NEAR); P.p23 = -2.f * VH_FAR * VH_NEAR / (VH_FAR - VH_NEAR);
  return P;
}

// ------------------------------------------------------------------------------------------------ pose / joints
// J = J_regressor (template + S beta + offset) evaluated through the precomputed JS = J_regressor S  (lbs.py:154,198-215)
__global__ void __launch_bounds__(256) k_pose_fwd(
    const float* __restrict__ shape, const float* __restrict__ expr, const float* __restrict__ rot, const float* __restrict__ neck,
    const float* __restrict__ jaw, const float* __restrict__ eyes, const float* __restrict__ offset, const int* __restrict__ ts,
    const float* __restrict__ JS, const float* __restrict__ Jt, const float* __restrict__ Jreg,
    int V, int K, int n_shape, float* __restrict__ betas_out, PoseFwd* __restrict__ posebuf, float* __restrict__ poses) { VH_PDL_SYNC();
  __shared__ float sh[33];
  __shared__ float shJ[15];
  int b = blockIdx.x, t = ts[b], n_expr = K - n_shape;
  float acc[15];
  for (int e = 0; e < 15; ++e) acc[e] = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float be = k < n_shape ? shape[k] : expr[(size_t)t * n_expr + (k - n_shape)];
    if (betas_out) betas_out[(size_t)b * K + k] = be;
    const float* js = JS + (size_t)k * 15;
    for (int e = 0; e < 15; ++e) acc[e] += js[e] * be;
  }
  if (offset)
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      float ox = offset[v * 3], oy = offset[v * 3 + 1], oz = offset[v * 3 + 2];
      for (int j = 0; j < 5; ++j) { float w = Jreg[(size_t)j * V + v]; acc[j * 3] += w * ox; acc[j * 3 + 1] += w * oy; acc[j * 3 + 2] += w * oz; }
    }
  for (int e = 0; e < 15; ++e) { float s = block_sum(acc[e], sh); if (threadIdx.x == 0) shJ[e] = s + Jt[e]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float pose[15];
    for (int c = 0; c < 3; ++c) { pose[c] = rot[t * 3 + c]; pose[3 + c] = neck[t * 3 + c]; pose[6 + c] = jaw[t * 3 + c]; }
    for (int c = 0; c < 6; ++c) pose[9 + c] = eyes[t * 6 + c];
    PoseFwd f;
    for (int e = 0; e < 15; ++e) f.J[e / 3][e % 3] = shJ[e];
    pose_forward(pose, f);
    posebuf[b] = f;
    for (int c = 0; c < 15; ++c) poses[b * 15 + c] = pose[c];
  }
}

// ------------------------------------------------------------------------------------------------ blend shapes
// v_shaped[b][m] = template[m] + offset[m] + sum_k S[k][m] beta[b][k]   (lbs.py:218-239, flame.py:602-608)
// S_fwd layout [K][3V]: a thread owns row m, every load is coalesced across the warp, betas are broadcast from smem.
// The shape part (n_shape columns) is identical for all frames and accumulated once.
#define BLEND_KS 8
template <int NB>
__global__ void __launch_bounds__(128) k_blend_fwd(const float* __restrict__ S, const float* __restrict__ tmpl, const float* __restrict__ offset,
                                                   const float* __restrict__ betas, int M, int K, int n_shape, int B, float* __restrict__ out) { VH_PDL_SYNC();
  extern __shared__ float shb[];      // [NB][K]
  int b0 = blockIdx.y * NB, nb = min(NB, B - b0);
  for (int i = threadIdx.x; i < nb * K; i += blockDim.x) shb[i] = betas[(size_t)b0 * K + i];
  __syncthreads();
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  // K is split over gridDim.z slices (more loads in flight per SM); slice 0 also adds template + offset
  int kper = (K + gridDim.z - 1) / gridDim.z, k0 = blockIdx.z * kper, k1 = min(k0 + kper, K);
  float common = blockIdx.z == 0 ? tmpl[m] + (offset ? offset[m] : 0.f) : 0.f;
  const float* s = S + m;
  int ks = min(k1, n_shape);
#pragma unroll 8
  for (int k = k0; k < ks; ++k) common += s[(size_t)k * M] * shb[k];
  float acc[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) acc[i] = 0.f;
#pragma unroll 4
  for (int k = max(k0, n_shape); k < k1; ++k) {
    float sv = s[(size_t)k * M];
#pragma unroll
    for (int i = 0; i < NB; ++i) acc[i] += sv * shb[i * K + k];
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) if (i < nb) out[((size_t)blockIdx.z * B + b0 + i) * M + m] = common + acc[i];   // per-slice partial, summed by k_skin_fwd
}

// ------------------------------------------------------------------------------------------------ skin + project + snap
// pose correctives (lbs.py:161-173), skinning (lbs.py:181-193), translation (flame.py:624), world->camera->clip
// (render_nvdiffrast.py:162-197) and the rasteriser's fixed-point snap (oracle/raster.py snap_vertices), one pass.
#define SNAP_GUARD 131072.f
template <int NB>
__global__ void __launch_bounds__(128) k_skin_fwd(const float* __restrict__ v_part, int KS, float* __restrict__ v_shaped, const float* __restrict__ posedirs, const float* __restrict__ lbs_w,
                                                  const PoseFwd* __restrict__ posebuf, const float* __restrict__ transl, const int* __restrict__ ts,
                                                  const CamParams* __restrict__ cam, int V, int B, int H, int W,
                                                  float* __restrict__ v_posed, f4* __restrict__ verts, f4* __restrict__ clip, i4* __restrict__ snap, float* __restrict__ ndc) { VH_PDL_SYNC();
  __shared__ float shpf[NB][36];
  __shared__ float shA[NB][60];
  int b0 = blockIdx.y * NB, nb = min(NB, B - b0);
  for (int i = threadIdx.x; i < nb * 36; i += blockDim.x) shpf[i / 36][i % 36] = posebuf[b0 + i / 36].pf[i % 36];
  for (int i = threadIdx.x; i < nb * 60; i += blockDim.x) shA[i / 60][i % 60] = ((const float*)posebuf[b0 + i / 60].A)[i % 60];
  __syncthreads();
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  int M = 3 * V;
  float vp[NB][3];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    if (i < nb) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      for (int z = 0; z < KS; ++z) { const float* s = v_part + ((size_t)z * B + b0 + i) * M + 3 * v; a0 += s[0]; a1 += s[1]; a2 += s[2]; }
      float* o = v_shaped + (size_t)(b0 + i) * M + 3 * v; o[0] = a0; o[1] = a1; o[2] = a2;
      vp[i][0] = a0; vp[i][1] = a1; vp[i][2] = a2;
    }
    else vp[i][0] = vp[i][1] = vp[i][2] = 0.f;
  }
  for (int p = 0; p < 36; ++p) {
    const float* pd = posedirs + (size_t)p * M + 3 * v;
    float d0 = pd[0], d1 = pd[1], d2 = pd[2];
#pragma unroll
    for (int i = 0; i < NB; ++i) { float f = shpf[i][p]; vp[i][0] += f * d0; vp[i][1] += f * d1; vp[i][2] += f * d2; }
  }
  float w[5];
  for (int j = 0; j < 5; ++j) w[j] = lbs_w[(size_t)v * 5 + j];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    if (i >= nb) continue;
    int b = b0 + i;
    float T[12];
    for (int e = 0; e < 12; ++e) T[e] = w[0] * shA[i][e] + w[1] * shA[i][12 + e] + w[2] * shA[i][24 + e] + w[3] * shA[i][36 + e] + w[4] * shA[i][48 + e];
    int t = ts[b];
    float x = T[0] * vp[i][0] + T[1] * vp[i][1] + T[2] * vp[i][2] + T[3] + transl[t * 3];
    float y = T[4] * vp[i][0] + T[5] * vp[i][1] + T[6] * vp[i][2] + T[7] + transl[t * 3 + 1];
    float z = T[8] * vp[i][0] + T[9] * vp[i][1] + T[10] * vp[i][2] + T[11] + transl[t * 3 + 2];
    float* o = v_posed + (size_t)b * M + 3 * v; o[0] = vp[i][0]; o[1] = vp[i][1]; o[2] = vp[i][2];
    f4 vw = {x, y, z, 1.f};
    verts[(size_t)b * V + v] = vw;
    if (!cam) continue;                           // view sharing: projected per view by k_project_views
    CamParams c = cam[b];
    float cx_ = c.RT[0] * x + c.RT[1] * y + c.RT[2] * z + c.RT[3];
    float cy_ = c.RT[4] * x + c.RT[5] * y + c.RT[6] * z + c.RT[7];
    float cz_ = c.RT[8] * x + c.RT[9] * y + c.RT[10] * z + c.RT[11];
    Proj P = make_proj(c, H, W);
    f4 cl = {P.p00 * cx_ + P.p02 * cz_, P.p11 * cy_ + P.p12 * cz_, P.p22 * cz_ + P.p23, -cz_};
    clip[(size_t)b * V + v] = cl;
    bool valid = isfinite(cl.x) && isfinite(cl.y) && isfinite(cl.z) && isfinite(cl.w) && cl.w > 0.f;
    i4 sn = {0, 0, 0, 0};
    if (valid) {
      float sx = __fmul_rn(__fdiv_rn(cl.x, cl.w), (float)(W * 8));
      float sy = __fmul_rn(__fdiv_rn(cl.y, cl.w), (float)(H * 8));
      sx = fminf(fmaxf(sx, -SNAP_GUARD), SNAP_GUARD); sy = fminf(fmaxf(sy, -SNAP_GUARD), SNAP_GUARD);
      sn.x = __float2int_rn(sx); sn.y = __float2int_rn(sy);
      sn.z = __float_as_int(__fdiv_rn(cl.z, cl.w)); sn.w = 1;
    }
    snap[(size_t)b * V + v] = sn;
    ndc[((size_t)b * V + v) * 2] = cl.x / cl.w; ndc[((size_t)b * V + v) * 2 + 1] = cl.y / cl.w;
  }
}

// view sharing: clip positions, raster snap and NDC of view b from the world vertices of its geometry geo[b] (the projection half of
// k_skin_fwd, identical arithmetic)
__global__ void __launch_bounds__(128) k_project_views(const f4* __restrict__ verts, const int* __restrict__ geo, const CamParams* __restrict__ cam, int V, int H, int W,
                                                        f4* __restrict__ clip, i4* __restrict__ snap, float* __restrict__ ndc) { VH_PDL_SYNC();
  int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (v >= V) return;
  f4 vw = verts[(size_t)geo[b] * V + v];
  float x = vw.x, y = vw.y, z = vw.z;
  CamParams c = cam[b];
  float cx_ = c.RT[0] * x + c.RT[1] * y + c.RT[2] * z + c.RT[3];
  float cy_ = c.RT[4] * x + c.RT[5] * y + c.RT[6] * z + c.RT[7];
  float cz_ = c.RT[8] * x + c.RT[9] * y + c.RT[10] * z + c.RT[11];
  Proj P = make_proj(c, H, W);
  f4 cl = {P.p00 * cx_ + P.p02 * cz_, P.p11 * cy_ + P.p12 * cz_, P.p22 * cz_ + P.p23, -cz_};
  clip[(size_t)b * V + v] = cl;
  bool valid = isfinite(cl.x) && isfinite(cl.y) && isfinite(cl.z) && isfinite(cl.w) && cl.w > 0.f;
  i4 sn = {0, 0, 0, 0};
  if (valid) {
    float sx = __fmul_rn(__fdiv_rn(cl.x, cl.w), (float)(W * 8));
    float sy = __fmul_rn(__fdiv_rn(cl.y, cl.w), (float)(H * 8));
    sx = fminf(fmaxf(sx, -SNAP_GUARD), SNAP_GUARD); sy = fminf(fmaxf(sy, -SNAP_GUARD), SNAP_GUARD);
    sn.x = __float2int_rn(sx); sn.y = __float2int_rn(sy);
    sn.z = __float_as_int(__fdiv_rn(cl.z, cl.w)); sn.w = 1;
  }
  snap[(size_t)b * V + v] = sn;
  ndc[((size_t)b * V + v) * 2] = cl.x / cl.w; ndc[((size_t)b * V + v) * 2 + 1] = cl.y / cl.w;
}
// adjoint: g_verts[geo[b]] += (d clip_b / d world)^T g_clip[b]; focal-length terms into acc (uncalibrated cameras)
__global__ void __launch_bounds__(128) k_project_views_bwd(const f4* __restrict__ verts, const int* __restrict__ geo, const CamParams* __restrict__ cam, const float* __restrict__ g_clip,
                                                            int V, int H, int W, int opt_cam, float* __restrict__ g_verts, float* __restrict__ acc) { VH_PDL_SYNC();
  __shared__ float shr[33];
  int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  float gfx = 0.f, gfy = 0.f;
  if (v < V) {
    const float* gc = g_clip + ((size_t)b * V + v) * 4;
    float gcx = gc[0], gcy = gc[1], gcz = gc[2], gcw = gc[3];
    if (gcx != 0.f || gcy != 0.f || gcz != 0.f || gcw != 0.f) {
      CamParams c = cam[b];
      Proj P = make_proj(c, H, W);
      const int g = geo[b];
      f4 x = verts[(size_t)g * V + v];
      float cx_ = c.RT[0] * x.x + c.RT[1] * x.y + c.RT[2] * x.z + c.RT[3];
      float cy_ = c.RT[4] * x.x + c.RT[5] * x.y + c.RT[6] * x.z + c.RT[7];
      float g_cx = P.p00 * gcx, g_cy = P.p11 * gcy, g_cz = P.p02 * gcx + P.p12 * gcy + P.p22 * gcz - gcw;
      gfx = gcx * cx_ * (2.f / W); gfy = gcy * cy_ * (2.f / H);
      float* o = g_verts + ((size_t)g * V + v) * 4;
      atomicAdd(o + 0, c.RT[0] * g_cx + c.RT[4] * g_cy + c.RT[8] * g_cz);
      atomicAdd(o + 1, c.RT[1] * g_cx + c.RT[5] * g_cy + c.RT[9] * g_cz);
      atomicAdd(o + 2, c.RT[2] * g_cx + c.RT[6] * g_cy + c.RT[10] * g_cz);
    }
  }
  if (opt_cam) {
    float s1 = block_sum(gfx, shr), s2 = block_sum(gfy, shr);
    if (threadIdx.x == 0) { atomicAdd(acc + ACC_GFX, s1); atomicAdd(acc + ACC_GFY, s2); }
  }
}

#ifndef VH_SKIN_NB
#define VH_SKIN_NB 2          // frames per CTA in the skinning kernel: 2 fills the machine at B=16 (8 left 82 CTAs for 148 SMs)
#endif
// blend-shape coefficients of the batch: beta[b] = [shape | expr[timestep[b]]]
// (+ the camera set-up of the step as one extra block when cam != NULL: saves a launch at the head of the chain)
__global__ void k_betas_gather(const float* __restrict__ shape, const float* __restrict__ expr, const int* __restrict__ ts, int K, int n_shape,
                               float* __restrict__ betas, int n_geo, CamParams* cam, const float* RT, const float* Kc, const float* focal, int n_views, int H, int W) { VH_PDL_SYNC();
  if ((int)blockIdx.x == n_geo) {
    for (int b = threadIdx.x; b < n_views; b += blockDim.x) cam[b] = cam_from(RT, Kc, focal, b, H, W);
    return;
  }
  int b = blockIdx.x, t = ts[b], n_expr = K - n_shape;
  for (int k = threadIdx.x; k < K; k += blockDim.x) betas[(size_t)b * K + k] = k < n_shape ? shape[k] : expr[(size_t)t * n_expr + (k - n_shape)];
}

void launch_flame_forward(vhap_ctx* c, const vhap_params* p, const vhap_frame_batch* fb, cudaStream_t s, bool with_cam) {
  // view sharing (fb->geo): FLAME runs once per distinct timestep (G geometries), the B views are only projected
  const bool shared = fb->geo != nullptr;
  const int nviews = fb->B;
  const int* ts = shared ? fb->geo_timesteps : fb->timesteps;
  int B = shared ? fb->n_geo : fb->B, V = c->V, M = 3 * V;
  // the joint / pose chain (latency-bound, one block per frame) only meets the blend shapes again in the skinning kernel: it runs
  // on an aux stream beside the tensor-core contraction
  const bool fork = !c->no_overlap;
  cudaStream_t sp = fork ? c->aux[0] : s;
  LAUNCH(c, KID_POSE_FWD, s, vh_launch(k_betas_gather, B + (with_cam ? 1 : 0), 128, 0, s, p->shape, p->expr, ts, c->K, c->n_shape, c->betas, B, c->cam, fb->RT, fb->K, p->focal_length,
                                       nviews, fb->H, fb->W));
  if (fork) { cudaEventRecord(c->ev[EV_POSE_FORK], s); cudaStreamWaitEvent(sp, c->ev[EV_POSE_FORK], 0); }
  LAUNCH(c, KID_POSE_FWD, sp, vh_launch(k_pose_fwd, B, 256, 0, sp, p->shape, p->expr, p->rotation, p->neck_pose, p->jaw_pose, p->eyes_pose, p->static_offset, ts,
                               c->JS, c->Jt, c->Jreg, V, c->K, c->n_shape, nullptr, c->posebuf, c->poses));
  if (fork) cudaEventRecord(c->ev[EV_POSE_DONE], sp);
  int ks = BLEND_KS;
  const float* vpart = c->v_shaped_part;
  if (c->use_tc_blend) {                 // tcgen05 contraction writes the finished v_shaped
    launch_blend_tc_fwd(c, p->static_offset, B, s);
    ks = 1; vpart = c->v_shaped;
  } else {
    dim3 g1((M + 127) / 128, (B + VH_MAXB_CHUNK - 1) / VH_MAXB_CHUNK, BLEND_KS);
    LAUNCH(c, KID_BLEND_FWD, s, vh_launch(k_blend_fwd<VH_MAXB_CHUNK>, g1, 128, VH_MAXB_CHUNK * c->K * sizeof(float), s, c->S_fwd, c->v_template, p->static_offset, c->betas, M, c->K,
                                                                                    c->n_shape, B, c->v_shaped_part));
  }
  if (fork) cudaStreamWaitEvent(s, c->ev[EV_POSE_DONE], 0);
  dim3 g2((V + 127) / 128, (B + VH_SKIN_NB - 1) / VH_SKIN_NB);
  LAUNCH(c, KID_SKIN_FWD, s, vh_launch(k_skin_fwd<VH_SKIN_NB>, g2, 128, 0, s, vpart, ks, c->v_shaped, c->posedirs, c->lbs_w, c->posebuf, p->translation, ts, shared ? nullptr : c->cam, V, B, fb->H, fb->W,
                                   c->v_posed, c->verts, c->clip, c->snap, c->ndc));
  if (shared) {
    dim3 g3((V + 127) / 128, nviews);
    LAUNCH(c, KID_SKIN_FWD, s, vh_launch(k_project_views, g3, 128, 0, s, c->verts, fb->geo, c->cam, V, fb->H, fb->W, c->clip, c->snap, c->ndc));
  }
}

// ------------------------------------------------------------------------------------------------ landmarks
// vertices2landmarks (lbs.py:60-98) + compute_lmk_energy (tracker.py:347-389): loss partial into acc[ACC_LMK] and the
// gradient scattered straight into g_verts / focal accumulators.  lmks_out optional; g_lmk_in = external upstream gradient.
__global__ void __launch_bounds__(96) k_landmarks(const f4* __restrict__ verts, const i4* __restrict__ faces, const int* __restrict__ lmk_faces,
                                                  const float* __restrict__ lmk_bary, const float* __restrict__ lmk2d, const CamParams* __restrict__ cam,
                                                  int V, int n_lmk, int H, int W, float w_scale, int jawline_off, int compute_loss, int opt_cam,
                                                  float* __restrict__ lmks_out, const float* __restrict__ g_lmk_in, float* __restrict__ g_verts,
                                                  float* __restrict__ acc, const int* __restrict__ geo) { VH_PDL_SYNC();
  __shared__ float sh[33];
  int b = blockIdx.x, l = threadIdx.x;
  const int gb = geo ? geo[b] : b;               // geometry of this view
  float loss = 0.f, gfx = 0.f, gfy = 0.f;
  if (l < n_lmk) {
    i4 f = faces[lmk_faces[l]];
    const f4* vb = verts + (size_t)gb * V;
    f4 a = vb[f.x], c1 = vb[f.y], d = vb[f.z];
    float b0 = lmk_bary[l * 3], b1 = lmk_bary[l * 3 + 1], b2 = lmk_bary[l * 3 + 2];
    float x = a.x * b0 + c1.x * b1 + d.x * b2, y = a.y * b0 + c1.y * b1 + d.y * b2, z = a.z * b0 + c1.z * b1 + d.z * b2;
    if (lmks_out) { float* o = lmks_out + ((size_t)b * n_lmk + l) * 3; o[0] = x; o[1] = y; o[2] = z; }
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (g_lmk_in) { const float* gi = g_lmk_in + ((size_t)b * n_lmk + l) * 3; gx = gi[0]; gy = gi[1]; gz = gi[2]; }
    int first = jawline_off ? 17 : 0;
    if (compute_loss && l >= first && l < 68) {
      CamParams c = cam[b];
      float cx_ = c.RT[0] * x + c.RT[1] * y + c.RT[2] * z + c.RT[3];
      float cy_ = c.RT[4] * x + c.RT[5] * y + c.RT[6] * z + c.RT[7];
      float cz_ = c.RT[8] * x + c.RT[9] * y + c.RT[10] * z + c.RT[11];
      Proj P = make_proj(c, H, W);
      float clx = P.p00 * cx_ + P.p02 * cz_, cly = P.p11 * cy_ + P.p12 * cz_, clw = -cz_;
      float px = clx / clw, py = -(cly / clw);                                   // world_to_ndc(flip_y=True)
      const float* gt = lmk2d + ((size_t)b * 68 + l) * 3;
      float u = 2.f * (gt[0] - W * 0.5f) / W, vv = 2.f * (gt[1] - H * 0.5f) / H;  // mesh.py:41-51
      float conf = gt[2];
      if (!jawline_off && l >= 27 && l < 36) conf *= 10.f;                        // tracker.py:379
      float dx = u - px, dy = vv - py;
      loss = (fabsf(dx) + fabsf(dy)) * conf * w_scale;
      float gpx = -(dx > 0.f ? 1.f : (dx < 0.f ? -1.f : 0.f)) * conf * w_scale;
      float gpy = -(dy > 0.f ? 1.f : (dy < 0.f ? -1.f : 0.f)) * conf * w_scale;
      float g_ndcx = gpx, g_ndcy = -gpy;
      float iw = 1.f / clw;
      float g_clx = g_ndcx * iw, g_cly = g_ndcy * iw, g_clw = -(g_ndcx * clx + g_ndcy * cly) * iw * iw;
      float g_cx = P.p00 * g_clx, g_cy = P.p11 * g_cly, g_cz = P.p02 * g_clx + P.p12 * g_cly - g_clw;
      if (opt_cam) { gfx = g_clx * cx_ * (2.f / W); gfy = g_cly * cy_ * (2.f / H); }
      gx += c.RT[0] * g_cx + c.RT[4] * g_cy + c.RT[8] * g_cz;
      gy += c.RT[1] * g_cx + c.RT[5] * g_cy + c.RT[9] * g_cz;
      gz += c.RT[2] * g_cx + c.RT[6] * g_cy + c.RT[10] * g_cz;
    }
    if (g_verts && (gx != 0.f || gy != 0.f || gz != 0.f)) {
      float* gv = g_verts + (size_t)gb * V * 4;
      const int vid[3] = {f.x, f.y, f.z};
      const float bb[3] = {b0, b1, b2};
      for (int k = 0; k < 3; ++k) {
        atomicAdd(gv + (size_t)vid[k] * 4 + 0, gx * bb[k]); atomicAdd(gv + (size_t)vid[k] * 4 + 1, gy * bb[k]); atomicAdd(gv + (size_t)vid[k] * 4 + 2, gz * bb[k]);
      }
    }
  }
  if (compute_loss) {
    float s0 = block_sum(loss, sh), s1 = block_sum(gfx, sh), s2 = block_sum(gfy, sh);
    if (threadIdx.x == 0) { atomicAdd(acc + ACC_LMK, s0); if (opt_cam) { atomicAdd(acc + ACC_GFX, s1); atomicAdd(acc + ACC_GFY, s2); } }
  }
}

void launch_landmarks(vhap_ctx* c, const vhap_frame_batch* fb, float w_scale, int jawline_off, float* lmks_out, float* g_lmk_in,
                      int compute_loss, int opt_cam, int global_B, cudaStream_t s) {
  (void)global_B;
  LAUNCH(c, KID_LMK, s, vh_launch(k_landmarks, fb->B, 96, 0, s, c->verts, c->faces, c->lmk_faces, c->lmk_bary, fb->lmk2d, c->cam, c->V, c->n_lmk, fb->H, fb->W, w_scale, jawline_off,
                                   compute_loss, opt_cam, lmks_out, g_lmk_in, c->g_verts, c->acc, fb->geo));
}

// ------------------------------------------------------------------------------------------------ vertex normals
__global__ void __launch_bounds__(128) k_vnormals(const f4* __restrict__ verts, const i4* __restrict__ faces, const int* __restrict__ indptr,
                                                  const int* __restrict__ vfaces, int V, f4* __restrict__ vnraw, f4* __restrict__ vnorm) { VH_PDL_SYNC();
  int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (v >= V) return;
  const f4* vb = verts + (size_t)b * V;
  f3 n = mk3(0, 0, 0);
  for (int i = indptr[v]; i < indptr[v + 1]; ++i) {
    i4 f = faces[vfaces[i]];
    f4 a = vb[f.x], c = vb[f.y], d = vb[f.z];
    n = n + cross3(mk3(c.x - a.x, c.y - a.y, c.z - a.z), mk3(d.x - a.x, d.y - a.y, d.z - a.z));
  }
  f4 raw = {n.x, n.y, n.z, 0.f};
  vnraw[(size_t)b * V + v] = raw;
  float d2 = dot3(n, n);
  if (!(d2 > 1e-20f)) { n = mk3(0, 0, 1); d2 = 1.f; }                    // render_nvdiffrast.py:312
  float il = 1.f / sqrtf(fmaxf(d2, 1e-20f));
  f4 o = {n.x * il, n.y * il, n.z * il, 0.f};
  vnorm[(size_t)b * V + v] = o;
}
void launch_vnormals(vhap_ctx* c, int B, cudaStream_t s) {
  dim3 g((c->V + 127) / 128, B);
  LAUNCH(c, KID_VNORM, s, vh_launch(k_vnormals, g, 128, 0, s, c->verts, c->faces, c->vf_indptr, c->vf_faces, c->V, c->vnraw, c->vnorm));
}

__global__ void __launch_bounds__(128) k_vnormals_bwd(const f4* __restrict__ verts, const i4* __restrict__ faces, const f4* __restrict__ vnraw,
                                                      const float* __restrict__ g_vnorm, int V, int F, float* __restrict__ g_verts) { VH_PDL_SYNC();
  int fi = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (fi >= F) return;
  i4 f = faces[fi];
  const int vid[3] = {f.x, f.y, f.z};
  f3 gfn = mk3(0, 0, 0);
  for (int k = 0; k < 3; ++k) {
    f4 r = vnraw[(size_t)b * V + vid[k]];
    f3 raw = mk3(r.x, r.y, r.z);
    float d2 = dot3(raw, raw);
    if (!(d2 > 1e-20f)) continue;                                          // constant fallback normal
    const float* gp = g_vnorm + ((size_t)b * V + vid[k]) * 4;
    f3 g = mk3(gp[0], gp[1], gp[2]);
    float il = 1.f / sqrtf(d2);
    f3 n = raw * il;
    gfn = gfn + (g - n * dot3(n, g)) * il;
  }
  const f4* vb = verts + (size_t)b * V;
  f4 a = vb[f.x], c = vb[f.y], d = vb[f.z];
  f3 e1 = mk3(c.x - a.x, c.y - a.y, c.z - a.z), e2 = mk3(d.x - a.x, d.y - a.y, d.z - a.z);
  f3 g1 = cross3(e2, gfn), g2 = cross3(gfn, e1);
  float* gv = g_verts + (size_t)b * V * 4;
  atomicAdd(gv + (size_t)f.y * 4, g1.x); atomicAdd(gv + (size_t)f.y * 4 + 1, g1.y); atomicAdd(gv + (size_t)f.y * 4 + 2, g1.z);
  atomicAdd(gv + (size_t)f.z * 4, g2.x); atomicAdd(gv + (size_t)f.z * 4 + 1, g2.y); atomicAdd(gv + (size_t)f.z * 4 + 2, g2.z);
  atomicAdd(gv + (size_t)f.x * 4, -g1.x - g2.x); atomicAdd(gv + (size_t)f.x * 4 + 1, -g1.y - g2.y); atomicAdd(gv + (size_t)f.x * 4 + 2, -g1.z - g2.z);
}
void launch_vnormals_bwd(vhap_ctx* c, int B, cudaStream_t s) {
  dim3 g((c->F + 127) / 128, B);
  LAUNCH(c, KID_VNORM_BWD, s, vh_launch(k_vnormals_bwd, g, 128, 0, s, c->verts, c->faces, c->vnraw, c->g_vnorm, c->V, c->F, c->g_verts));
}

// ------------------------------------------------------------------------------------------------ skinning backward
template <int NB>
__global__ void __launch_bounds__(128) k_skin_bwd(const float* __restrict__ v_posed, const f4* __restrict__ verts, const float* __restrict__ posedirs,
                                                  const float* __restrict__ lbs_w, const PoseFwd* __restrict__ posebuf, const CamParams* __restrict__ cam,
                                                  const int* __restrict__ ts, const float* __restrict__ g_verts, const float* __restrict__ g_clip,
                                                  int V, int B, int H, int W, int opt_cam, int Mp,
                                                  float* __restrict__ g_vshaped, float* __restrict__ g_offset, float* __restrict__ g_transl,
                                                  float* __restrict__ gA, float* __restrict__ gpf, float* __restrict__ acc) { VH_PDL_SYNC();
  __shared__ float shA[NB][60];
  __shared__ float shr[33];
  __shared__ float red[60][128];
  int b0 = blockIdx.y * NB, nb = min(NB, B - b0);
  for (int i = threadIdx.x; i < nb * 60; i += blockDim.x) shA[i / 60][i % 60] = ((const float*)posebuf[b0 + i / 60].A)[i % 60];
  int v = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  bool on = v < V;
  int M = 3 * V;
  float w[5] = {0, 0, 0, 0, 0};
  if (on) for (int j = 0; j < 5; ++j) w[j] = lbs_w[(size_t)v * 5 + j];
  float goff[3] = {0, 0, 0};
  float gfx = 0.f, gfy = 0.f;
  __syncthreads();
  for (int i = 0; i < nb; ++i) {
    int b = b0 + i;
    float g[3] = {0, 0, 0}, vp[3] = {0, 0, 0}, gvp[3] = {0, 0, 0};
    if (on) {
      const float* gv = g_verts + ((size_t)b * V + v) * 4;
      g[0] = gv[0]; g[1] = gv[1]; g[2] = gv[2];
      float gcx = 0.f, gcy = 0.f, gcz = 0.f, gcw = 0.f;
      if (g_clip) { const float* gc = g_clip + ((size_t)b * V + v) * 4; gcx = gc[0]; gcy = gc[1]; gcz = gc[2]; gcw = gc[3]; }   // (view sharing: already in g_verts)
      if (gcx != 0.f || gcy != 0.f || gcz != 0.f || gcw != 0.f) {
        CamParams c = cam[b];
        Proj P = make_proj(c, H, W);
        f4 x = verts[(size_t)b * V + v];
        float cx_ = c.RT[0] * x.x + c.RT[1] * x.y + c.RT[2] * x.z + c.RT[3];
        float cy_ = c.RT[4] * x.x + c.RT[5] * x.y + c.RT[6] * x.z + c.RT[7];
        float g_cx = P.p00 * gcx, g_cy = P.p11 * gcy, g_cz = P.p02 * gcx + P.p12 * gcy + P.p22 * gcz - gcw;
        gfx += gcx * cx_ * (2.f / W); gfy += gcy * cy_ * (2.f / H);
        g[0] += c.RT[0] * g_cx + c.RT[4] * g_cy + c.RT[8] * g_cz;
        g[1] += c.RT[1] * g_cx + c.RT[5] * g_cy + c.RT[9] * g_cz;
        g[2] += c.RT[2] * g_cx + c.RT[6] * g_cy + c.RT[10] * g_cz;
      }
      const float* vpp = v_posed + (size_t)b * M + 3 * v;
      vp[0] = vpp[0]; vp[1] = vpp[1]; vp[2] = vpp[2];
      float T[12];
      for (int e = 0; e < 12; ++e) T[e] = w[0] * shA[i][e] + w[1] * shA[i][12 + e] + w[2] * shA[i][24 + e] + w[3] * shA[i][36 + e] + w[4] * shA[i][48 + e];
      gvp[0] = T[0] * g[0] + T[4] * g[1] + T[8] * g[2];
      gvp[1] = T[1] * g[0] + T[5] * g[1] + T[9] * g[2];
      gvp[2] = T[2] * g[0] + T[6] * g[1] + T[10] * g[2];
      float* o = g_vshaped + (size_t)b * Mp + 3 * v; o[0] = gvp[0]; o[1] = gvp[1]; o[2] = gvp[2];
      goff[0] += gvp[0]; goff[1] += gvp[1]; goff[2] += gvp[2];
    }
    // block reductions through shared memory: every thread deposits its values column-wise (red[value][thread], conflict
    // free), then each warp sums a quarter of the rows (4 strided reads + a warp reduction) and issues one atomic per row
    int warp = threadIdx.x >> 5;
    __syncthreads();
    for (int j = 0; j < 5; ++j)                         // g_A[j][r][ci]  (lbs.py:185: T = W A)
      for (int r = 0; r < 3; ++r)
        for (int ci = 0; ci < 4; ++ci) red[j * 12 + r * 4 + ci][threadIdx.x] = w[j] * g[r] * (ci < 3 ? vp[ci] : 1.f);
    __syncthreads();
    for (int row = warp; row < 60; row += 4) {
      float val = warp_sum(red[row][lane] + red[row][lane + 32] + red[row][lane + 64] + red[row][lane + 96]);
      if (lane == 0 && val != 0.f) atomicAdd(gA + (size_t)b * 60 + row, val);
    }
    __syncthreads();
    for (int p = 0; p < 36; ++p) {                      // g_pose_feature[p] (lbs.py:166); loads coalesced over vertices
      float val = 0.f;
      if (on) { const float* pd = posedirs + (size_t)p * M + 3 * v; val = pd[0] * gvp[0] + pd[1] * gvp[1] + pd[2] * gvp[2]; }
      red[p][threadIdx.x] = val;
    }
    for (int cc = 0; cc < 3; ++cc) red[36 + cc][threadIdx.x] = g[cc];     // translation (flame.py:624)
    __syncthreads();
    for (int row = warp; row < 39; row += 4) {
      float val = warp_sum(red[row][lane] + red[row][lane + 32] + red[row][lane + 64] + red[row][lane + 96]);
      if (lane == 0 && val != 0.f) {
        if (row < 36) atomicAdd(gpf + (size_t)b * 36 + row, val);
        else if (g_transl) atomicAdd(g_transl + (size_t)ts[b] * 3 + (row - 36), val);
      }
    }
  }
  if (g_offset && on) { atomicAdd(g_offset + 3 * v, goff[0]); atomicAdd(g_offset + 3 * v + 1, goff[1]); atomicAdd(g_offset + 3 * v + 2, goff[2]); }
  if (opt_cam) {
    float s1 = block_sum(gfx, shr), s2 = block_sum(gfy, shr);
    if (threadIdx.x == 0) { atomicAdd(acc + ACC_GFX, s1); atomicAdd(acc + ACC_GFY, s2); }
  }
}

__global__ void __launch_bounds__(128) k_pose_bwd(const float* __restrict__ poses, const PoseFwd* __restrict__ posebuf, const float* __restrict__ gA,
                                                  const float* __restrict__ gpf, const int* __restrict__ ts, const float* __restrict__ JS, int K,
                                                  float* g_rot, float* g_neck, float* g_jaw, float* g_eyes, float* __restrict__ gJ_out,
                                                  float* __restrict__ gbetas) { VH_PDL_SYNC();
  __shared__ float shJ[15];
  int b = blockIdx.x;
  if (threadIdx.x == 0) {
    float gp[15], gj[VH_NJ][3];
    PoseFwd f = posebuf[b];
    pose_backward(poses + b * 15, f, (const float(*)[12])(gA + (size_t)b * 60), gpf + (size_t)b * 36, gp, gj);
    int t = ts[b];
    for (int c = 0; c < 3; ++c) {
      if (g_rot) atomicAdd(g_rot + t * 3 + c, gp[c]);
      if (g_neck) atomicAdd(g_neck + t * 3 + c, gp[3 + c]);
      if (g_jaw) atomicAdd(g_jaw + t * 3 + c, gp[6 + c]);
    }
    if (g_eyes) for (int c = 0; c < 6; ++c) atomicAdd(g_eyes + t * 6 + c, gp[9 + c]);
    for (int e = 0; e < 15; ++e) { shJ[e] = gj[e / 3][e % 3]; gJ_out[b * 15 + e] = shJ[e]; }
  }
  __syncthreads();
  if (gbetas)
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const float* js = JS + (size_t)k * 15;
      float s = 0.f;
      for (int e = 0; e < 15; ++e) s += js[e] * shJ[e];
      atomicAdd(gbetas + (size_t)b * K + k, s);
    }
}

__global__ void k_joff_bwd(const float* __restrict__ Jreg, const float* __restrict__ gJ, int V, int B, float* __restrict__ g_offset) { VH_PDL_SYNC();
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float s[3] = {0, 0, 0};
  for (int j = 0; j < 5; ++j) {
    float w = Jreg[(size_t)j * V + v];
    if (w == 0.f) continue;
    for (int b = 0; b < B; ++b) { const float* g = gJ + b * 15 + j * 3; s[0] += w * g[0]; s[1] += w * g[1]; s[2] += w * g[2]; }
  }
  atomicAdd(g_offset + 3 * v, s[0]); atomicAdd(g_offset + 3 * v + 1, s[1]); atomicAdd(g_offset + 3 * v + 2, s[2]);
}

// g_betas[b][k] = sum_m g_vshaped[b][m] S[m][k]  -- the transposed contraction (autograd of lbs.py:238).
// S_bwd layout [3V][K]: thread <-> k, rows streamed once per chunk of frames, no cross-thread reduction.
#define BB_ROWS 64
template <int NB>
__global__ void __launch_bounds__(512) k_blend_bwd(const float* __restrict__ S, const float* __restrict__ g_vshaped, int M, int Mp, int K, int B,
                                                   float* __restrict__ gbetas) { VH_PDL_SYNC();
  __shared__ float sg[NB][BB_ROWS];
  int m0 = blockIdx.x * BB_ROWS, nm = min(BB_ROWS, M - m0);
  int b0 = blockIdx.y * NB, nb = min(NB, B - b0);
  for (int i = threadIdx.x; i < NB * BB_ROWS; i += blockDim.x) {
    int bi = i / BB_ROWS, mi = i % BB_ROWS;
    sg[bi][mi] = (bi < nb && mi < nm) ? g_vshaped[(size_t)(b0 + bi) * Mp + m0 + mi] : 0.f;
  }
  __syncthreads();
  int k = threadIdx.x;
  if (k >= K) return;
  float acc[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) acc[i] = 0.f;
  const float* s = S + (size_t)m0 * K + k;
#pragma unroll 4
  for (int mi = 0; mi < nm; ++mi) {
    float sv = s[(size_t)mi * K];
#pragma unroll
    for (int i = 0; i < NB; ++i) acc[i] += sv * sg[i][mi];
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) if (i < nb) atomicAdd(gbetas + (size_t)(b0 + i) * K + k, acc[i]);
}

__global__ void k_betas_scatter(const float* __restrict__ gbetas, const int* __restrict__ ts, int K, int n_shape, float* g_shape, float* g_expr) { VH_PDL_SYNC();
  int b = blockIdx.x, n_expr = K - n_shape;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float g = gbetas[(size_t)b * K + k];
    if (k < n_shape) { if (g_shape) atomicAdd(g_shape + k, g); }
    else if (g_expr) atomicAdd(g_expr + (size_t)ts[b] * n_expr + (k - n_shape), g);
  }
}

void launch_flame_backward(vhap_ctx* c, const vhap_params* p, const vhap_frame_batch* fb, const vhap_grads* g, int opt_cam, cudaStream_t s) {
  const bool shared = fb->geo != nullptr;
  const int* ts = shared ? fb->geo_timesteps : fb->timesteps;
  int B = shared ? fb->n_geo : fb->B, V = c->V, M = 3 * V;
  bool need_betas = g->shape || g->expr;
  if (shared) {                                   // projection adjoint of every view into its geometry's world-space gradient
    dim3 g0((V + 127) / 128, fb->B);
    LAUNCH(c, KID_SKIN_BWD, s, vh_launch(k_project_views_bwd, g0, 128, 0, s, c->verts, fb->geo, c->cam, c->g_clip, V, fb->H, fb->W, opt_cam, c->g_verts, c->acc));
  }
  dim3 g1((V + 127) / 128, (B + 1) / 2);
  LAUNCH(c, KID_SKIN_BWD, s, vh_launch(k_skin_bwd<2>, g1, 128, 0, s, c->v_posed, c->verts, c->posedirs, c->lbs_w, c->posebuf, c->cam, ts, c->g_verts, shared ? nullptr : c->g_clip, V, B, fb->H, fb->W,
                                   shared ? 0 : opt_cam, c->Mpad, c->g_vshaped, g->static_offset, g->translation, c->gA, c->gpf, c->acc));
  // the blend-shape adjoint (tensor-core contraction of g_vshaped) only needs skin_bwd's output: it runs beside the serial
  // pose_bwd -> joff_bwd pair on the second high-priority stream; both add into gbetas atomically
  bool forked = need_betas && s == c->hp[0];
  cudaStream_t s2 = forked ? c->hp[1] : s;
  if (forked) { cudaEventRecord(c->ev[EV_BLEND_FORK], s); cudaStreamWaitEvent(s2, c->ev[EV_BLEND_FORK], 0); }
  if (need_betas) {
    if (c->use_tc_blend) launch_blend_tc_bwd(c, B, s2);
    else {
      dim3 g2((M + BB_ROWS - 1) / BB_ROWS, (B + VH_MAXB_CHUNK - 1) / VH_MAXB_CHUNK);
      LAUNCH(c, KID_BLEND_BWD, s2, vh_launch(k_blend_bwd<VH_MAXB_CHUNK>, g2, 512, 0, s2, c->S_bwd, c->g_vshaped, M, c->Mpad, c->K, B, c->gbetas));
    }
    if (forked) cudaEventRecord(c->ev[EV_BLEND_DONE], s2);
  }
  LAUNCH(c, KID_POSE_BWD, s, vh_launch(k_pose_bwd, B, 128, 0, s, c->poses, c->posebuf, c->gA, c->gpf, ts, c->JS, c->K, g->rotation, g->neck_pose, g->jaw_pose, g->eyes_pose,
                               c->gJ, need_betas ? c->gbetas : nullptr));
  if (g->static_offset) LAUNCH(c, KID_JOFF_BWD, s, vh_launch(k_joff_bwd, (V + 127) / 128, 128, 0, s, c->Jreg, c->gJ, V, B, g->static_offset));
  if (need_betas) {
    if (forked) cudaStreamWaitEvent(s, c->ev[EV_BLEND_DONE], 0);
    LAUNCH(c, KID_BETAS_SCATTER, s, vh_launch(k_betas_scatter, B, 256, 0, s, c->gbetas, ts, c->K, c->n_shape, g->shape, g->expr));
  }
  (void)p;
}

// ------------------------------------------------------------------------------------------------ small regularisers
// compute_regularization_energy (tracker.py:480-605) for the parameter-space terms; one block, losses into acc[], gradients
// added to the dense grads.  Temporal terms compare with the detached previous timestep (tracker.py:616-648).
struct RegArgs {
  vhap_params p; vhap_grads g; vhap_stage_cfg cfg;
  const int* ts; int B, V, n_shape, n_expr, global_B;
  const float *w_off, *w_off_lap; const int *lap_indptr, *lap_idx; const float* lap_val; float* lap_y;
  const int *rigid_indptr, *rigid_vids; int n_rigid;
  float* acc;
};

__device__ void reg_sq_rows(const float* x, float* gx, const int* ts, int B, int D, float w_over_n, float* loss) {
  // w * mean(x[ts]^2): every (b, d) element counts (duplicated timesteps count twice, like indexing with repeats)
  for (int i = threadIdx.x; i < B * D; i += blockDim.x) {
    int t = ts[i / D], d = i % D;
    float v = x[(size_t)t * D + d];
    *loss += w_over_n * v * v;
    if (gx) atomicAdd(gx + (size_t)t * D + d, 2.f * w_over_n * v);
  }
}
__device__ void reg_smooth_rows(const float* x, float* gx, const int* ts, int B, int D, int n_t, float w_over_n, float* loss) {
  for (int i = threadIdx.x; i < B * D; i += blockDim.x) {
    int t = ts[i / D], d = i % D, tp = max(t - 1, 0);
    tp = min(tp, n_t - 1);
    float df = x[(size_t)t * D + d] - x[(size_t)tp * D + d];
    *loss += w_over_n * df * df;
    if (gx) atomicAdd(gx + (size_t)t * D + d, 2.f * w_over_n * df);
  }
}

// mean((R(0) - R(pose))^2) over [B,3,3] via rodrigues (tracker.py:662-664)
// NB the reference stacks B zero poses in front of the B poses and averages over rotmats[1:], i.e. (2B-1)*9 elements.
__device__ void reg_joint_rot(const float* x, float* gx, const int* ts, int B, int GB, int D, int off, float w, float* loss) {
  float R0[9], z[3] = {0, 0, 0};
  rodrigues(z, R0);
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int t = ts[b];
    const float* r = x + (size_t)t * D + off;
    float R[9], gR[9];
    rodrigues(r, R);
    float sc = w / (9.f * (2.f * GB - 1.f));
    for (int i = 0; i < 9; ++i) { float df = R0[i] - R[i]; *loss += sc * df * df; gR[i] = -2.f * sc * df; }
    if (gx) { float gr[3]; rodrigues_bwd(r, gR, gr); for (int c = 0; c < 3; ++c) atomicAdd(gx + (size_t)t * D + off + c, gr[c]); }
  }
}

__global__ void __launch_bounds__(1024) k_regs(RegArgs a) { VH_PDL_SYNC();
  __shared__ float sh[33];
  const vhap_stage_cfg& c = a.cfg;
  int B = a.B, GB = a.global_B;
  const int* ts = a.ts;
  const int sec = blockIdx.x;      // 0: per-frame + shape + lights, 1: offset L1, 2: offset Laplacian, 3: offset rigidity
  // ---- expression / shape (tracker.py:508-520)
  if (sec == 0 && c.opt_expr) {
    float l = 0.f;
    reg_sq_rows(a.p.expr, a.g.expr, ts, B, a.n_expr, c.w_reg_expr / ((float)GB * a.n_expr), &l);
    l = block_sum(l, sh); if (threadIdx.x == 0) a.acc[ACC_REG_EXPR] += l;
    if (c.tracking) {
      l = 0.f; reg_smooth_rows(a.p.expr, a.g.expr, ts, B, a.n_expr, a.p.n_timesteps, c.w_smooth_expr / ((float)GB * a.n_expr), &l);
      l = block_sum(l, sh); if (threadIdx.x == 0) a.acc[ACC_SMOOTH_EXPR] += l;
    }
  }
  if (sec == 0 && c.opt_shape) {
    float l = 0.f;
    for (int k = threadIdx.x; k < a.n_shape; k += blockDim.x) {
      float v = a.p.shape[k], w = c.shared_scale * c.w_reg_shape / a.n_shape;
      l += w * v * v;
      if (a.g.shape) atomicAdd(a.g.shape + k, 2.f * w * v);
    }
    l = block_sum(l, sh); if (threadIdx.x == 0) a.acc[ACC_REG_SHAPE] += l;
  }
  // ---- pose smoothness (tracker.py:616-627)
  if (sec == 0 && c.opt_pose && c.tracking) {
    float l = 0.f;
    reg_smooth_rows(a.p.translation, a.g.translation, ts, B, 3, a.p.n_timesteps, c.w_smooth_trans / (3.f * GB), &l);
    reg_smooth_rows(a.p.rotation, a.g.rotation, ts, B, 3, a.p.n_timesteps, c.w_smooth_rot / (3.f * GB), &l);
    l = block_sum(l, sh); if (threadIdx.x == 0) a.acc[ACC_SMOOTH_POSE] += l;
  }
  // ---- joints (tracker.py:496-505, 629-680)
  if (sec == 0 && c.opt_joints) {
    float l = 0.f;
    reg_joint_rot(a.p.neck_pose, a.g.neck_pose, ts, B, GB, 3, 0, c.w_reg_neck, &l);
    reg_joint_rot(a.p.jaw_pose, a.g.jaw_pose, ts, B, GB, 3, 0, c.w_reg_jaw, &l);
    reg_joint_rot(a.p.eyes_pose, a.g.eyes_pose, ts, B, GB, 6, 0, c.w_reg_eyes, &l);
    reg_joint_rot(a.p.eyes_pose, a.g.eyes_pose, ts, B, GB, 6, 3, c.w_reg_eyes, &l);
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      int t = ts[b];
      const float* j = a.p.jaw_pose + (size_t)t * 3;
      float wj = c.w_reg_jaw / GB;
      if (-j[0] > 0.f) { l += wj * 10.f * (-j[0]); if (a.g.jaw_pose) atomicAdd(a.g.jaw_pose + t * 3, -10.f * wj); }     // relu(-x)*10
      for (int d = 1; d < 3; ++d) { l += wj * 3.f * j[d] * j[d] / 2.f; if (a.g.jaw_pose) atomicAdd(a.g.jaw_pose + t * 3 + d, wj * 3.f * j[d]); }
      const float* e = a.p.eyes_pose + (size_t)t * 6;
      float we = c.w_reg_eyes / GB * 2.f;                                     // added once per eye term (two loop passes, tracker.py:675)
      for (int d = 0; d < 3; ++d) {
        float df = e[d] - e[3 + d];
        l += we * df * df / 3.f;
        if (a.g.eyes_pose) { atomicAdd(a.g.eyes_pose + t * 6 + d, 2.f * we * df / 3.f); atomicAdd(a.g.eyes_pose + t * 6 + 3 + d, -2.f * we * df / 3.f); }
      }
    }
    l = block_sum(l, sh); if (threadIdx.x == 0) a.acc[ACC_REG_JOINT] += l;
    if (c.tracking) {
      l = 0.f;
      reg_smooth_rows(a.p.neck_pose, a.g.neck_pose, ts, B, 3, a.p.n_timesteps, c.w_smooth_neck / (3.f * GB), &l);
      reg_smooth_rows(a.p.jaw_pose, a.g.jaw_pose, ts, B, 3, a.p.n_timesteps, c.w_smooth_jaw / (3.f * GB), &l);
      reg_smooth_rows(a.p.eyes_pose, a.g.eyes_pose, ts, B, 6, a.p.n_timesteps, c.w_smooth_eyes / (6.f * GB), &l);
      l = block_sum(l, sh); if (threadIdx.x == 0) a.acc[ACC_SMOOTH_JOINT] += l;
    }
  }
  // ---- lights (tracker.py:542-545)
  if (sec == 0 && c.opt_lights && c.w_reg_light >= 0.f) {
    float l = 0.f;
    for (int i = threadIdx.x; i < 27; i += blockDim.x) {
      float u = i < 3 ? 3.5449077018110318f : 0.f;          // sqrt(4 pi)
      float df = a.p.lights[i] - u, w = c.shared_scale * c.w_reg_light / 27.f;
      l += w * df * df;
      if (a.g.lights) atomicAdd(a.g.lights + i, 2.f * w * df);
    }
    l = block_sum(l, sh); if (threadIdx.x == 0) a.acc[ACC_REG_LIGHT] += l;
  }
  // ---- static offset (tracker.py:553-594).  B-independent: evaluated identically on every rank, never communicated.
  if (c.opt_static_offset && a.p.static_offset) {
    const float* off = a.p.static_offset;
    int V = a.V;
    // sections >= 4: one thread per vertex over a grid of helper blocks (no single-SM atomic bottleneck)
    if (sec >= 4) {
      int j = (sec - 4) * blockDim.x + threadIdx.x;
      float l1 = 0.f, l2 = 0.f;
      if (j < V) {
        if (c.w_reg_offset >= 0.f) {                                           // tracker.py:576-587
          float w = c.shared_scale * c.w_reg_offset / (3.f * V), wv = a.w_off ? a.w_off[j] : 1.f;
          for (int cc = 0; cc < 3; ++cc) {
            float o = off[3 * j + cc];
            l1 += w * wv * fabsf(o);
            if (a.g.static_offset && o != 0.f) atomicAdd(a.g.static_offset + 3 * j + cc, w * wv * (o > 0.f ? 1.f : -1.f));
          }
        }
        if (c.w_reg_offset_lap >= 0.f) {
          // L(base + off) - L(base) == L off (the uniform Laplacian is linear; tracker.py:682-690, flame.py:196-201).
          // gradient in gather form: g_j = sum_i L_ij * 2 w wv_i (L off)_i over the (symmetric) sparsity pattern of row j
          float w = c.shared_scale * c.w_reg_offset_lap / V;
          float gj[3] = {0, 0, 0};
          for (int qi = a.lap_indptr[j]; qi < a.lap_indptr[j + 1]; ++qi) {
            int i = a.lap_idx[qi];
            float y[3] = {0, 0, 0}, Lij = 0.f;
            for (int q = a.lap_indptr[i]; q < a.lap_indptr[i + 1]; ++q) {
              int k = a.lap_idx[q]; float lv = a.lap_val[q];
              y[0] += lv * off[3 * k]; y[1] += lv * off[3 * k + 1]; y[2] += lv * off[3 * k + 2];
              if (k == j) Lij = lv;
            }
            float wv = a.w_off_lap ? a.w_off_lap[i] : 1.f;
            if (i == j) l2 += w * wv * (y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
            float sc = 2.f * w * wv * Lij;
            gj[0] += sc * y[0]; gj[1] += sc * y[1]; gj[2] += sc * y[2];
          }
          if (a.g.static_offset) { atomicAdd(a.g.static_offset + 3 * j, gj[0]); atomicAdd(a.g.static_offset + 3 * j + 1, gj[1]); atomicAdd(a.g.static_offset + 3 * j + 2, gj[2]); }
        }
      }
      l1 = block_sum(l1, sh); l2 = block_sum(l2, sh);
      if (threadIdx.x == 0) { if (l1 != 0.f) atomicAdd(a.acc + ACC_REG_OFFSET, l1); if (l2 != 0.f) atomicAdd(a.acc + ACC_REG_OFFSET_LAP, l2); }
    }
    if (sec == 3 && c.w_reg_offset_rigid >= 0.f && a.n_rigid > 0) {
      float ltot = 0.f;
      for (int r = 0; r < a.n_rigid; ++r) {
        int s0 = a.rigid_indptr[r], n = a.rigid_indptr[r + 1] - s0;
        if (n < 2) continue;
        float mu[3];
        for (int cc = 0; cc < 3; ++cc) {
          float s = 0.f;
          for (int i = threadIdx.x; i < n; i += blockDim.x) s += off[3 * a.rigid_vids[s0 + i] + cc];
          mu[cc] = block_sum(s, sh) / n;
        }
        float w = c.shared_scale * c.w_reg_offset_rigid / (3.f * (n - 1));
        float l = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          int v = a.rigid_vids[s0 + i];
          for (int cc = 0; cc < 3; ++cc) {
            float df = off[3 * v + cc] - mu[cc];
            l += w * df * df;
            if (a.g.static_offset) atomicAdd(a.g.static_offset + 3 * v + cc, 2.f * w * df);
          }
        }
        ltot += l;
      }
      ltot = block_sum(ltot, sh); if (threadIdx.x == 0) a.acc[ACC_REG_OFFSET_RIGID] += ltot;
    }
  }
}

void launch_regs(vhap_ctx* c, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg, const vhap_grads* g, int global_B, cudaStream_t s) {
  RegArgs a;
  memset(&a, 0, sizeof(a));
  a.p = *p; if (g) a.g = *g; a.cfg = *cfg;
  a.ts = fb->timesteps; a.B = fb->B; a.V = c->V; a.n_shape = c->n_shape; a.n_expr = c->n_expr; a.global_B = global_B;
  a.w_off = c->w_off; a.w_off_lap = c->w_off_lap; a.lap_indptr = c->lap_indptr; a.lap_idx = c->lap_idx; a.lap_val = c->lap_val; a.lap_y = c->lap_y;
  a.rigid_indptr = c->rigid_indptr; a.rigid_vids = c->rigid_vids; a.n_rigid = c->n_rigid; a.acc = c->acc;
  LAUNCH(c, KID_REGS, s, vh_launch(k_regs, 4 + (c->V + 1023) / 1024, 1024, 0, s, a));
}
