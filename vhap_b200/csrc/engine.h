// Internal declarations shared by the translation units of libvhap_b200.so.
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/vhap_b200.h"
#include "common.cuh"
#include "flame_math.cuh"
#include "render_bodies.cuh"

#define VH_TILE 16                 // raster tile edge in pixels
#define VH_NPART 40                // floats per block-partial row (>= 27 + a few)
#define VH_MAXB_CHUNK 16
#define VH_DP_MAX 16                // ranks of one node
#define VH_DP_BOX_FLOATS (2 * VH_DP_MAX * 8 + 3 * VH_DP_MAX)   // mailbox: float slot[2][VH_DP_MAX][8], int flag[VH_DP_MAX] (slab), flagA / flagB[VH_DP_MAX] (texture barriers)           // frames processed per pass of the blend-shape kernels

struct CamParams { float RT[12]; float fx, fy, cx, cy; };

// fork / join events of the side chains of a step (record on one stream, wait on another; valid eagerly and inside stream capture)
enum { EV_REGS_FORK = 0, EV_REGS_DONE, EV_TEXGRAD_READY, EV_TEX_DONE, EV_VN_FORK, EV_VN_DONE, EV_GEOM_DONE, EV_BLEND_FORK, EV_BLEND_DONE,
       EV_POSE_FORK, EV_POSE_DONE, EV_LMK_DONE, EV_LIGHTS_DONE, EV_C1_DONE, EV_C2TEX_DONE, EV_COUNT };

struct vhap_ctx {
  char err[512];
  int device;
  int V, F, VT, K, n_shape, n_expr, n_lmk, n_clusters, T, max_level;
  int mip_off[VH_MAX_MIPS]; size_t mip_total;
  // ---- static model
  float *v_template, *S_fwd, *S_bwd, *posedirs, *Jreg, *lbs_w, *JS, *Jt;
  int tex_fold_reg;                               // VHAP_B200_TEXFOLD=reg: the register-pipelined fold kernel also for T >= 256 (A/B aid; default: TMA-staged)
  float* S_fwd_pad; int Mpad; int use_tc_blend;   // [K][Mpad] copy with 16-byte aligned rows for the tensor-core contraction
  i4 *faces, *faces_uv; float* verts_uv; int* lmk_faces; float* lmk_bary; int* adj_opp; uint8_t* fid2cid;
  int *vf_indptr, *vf_faces; int *lap_indptr, *lap_idx; float* lap_val; int lap_nnz;
  // ---- stage masks
  uint8_t *face_flags, *vert_flags; float *w_off, *w_off_lap; int *rigid_indptr, *rigid_vids; int n_rigid; uint8_t* uvmask_res;
  // ---- texture
  f4* mips[2]; int cur_mip; float* tex_painted; float* g_tex; float* tv_partials; int tv_nblocks;
  // ---- per-batch scratch
  int maxB, maxH, maxW, curB, curH, curW;
  float *v_shaped, *v_posed, *g_vshaped;          // [B][3V]
  float* v_shaped_part;                           // [8][B][3V] K-slice partial sums of the blend-shape contraction
  f4 *verts, *clip, *vnorm, *vnraw; i4* snap;     // [B][V]
  float* ndc;                                     // [B][V][2] clip.xy / clip.w
  float *g_clip, *g_vnorm, *g_verts;              // [B][V][4]
  PoseFwd* posebuf; float* poses;                 // [B], [B][15]
  float *gA, *gpf, *gJ, *gbetas, *betas;          // [B][60], [B][36], [B][15], [B][K], [B][K]
  CamParams* cam;                                 // [B]
  float* lap_y;                                   // [V][3]
  // pixels
  int* tri_id; f4* pre; uint8_t* signs; int* pool_list; float* final_rgba; f4 *plane_albedo, *plane_normal, *plane_diffuse;
  int want_planes;
  // raster binning
  int *tile_count, *tile_off, *tile_cursor, *tile_list; int tile_cap; int* overflow_flag;
  // pools
  int *pool_blk_count, *pool_blk_off, *pool_base, *pool_count; int pool_nblk;
  // reductions
  float* partials; int n_partials_rows;           // [rows][VH_NPART]
  unsigned long long* maxslot;                    // packed (orderable float bits << 32 | idx)
  float* scal;                                    // [16] device scalars for pass C
  float* acc;                                     // [64] misc accumulators (loss sums, focal grad...)
  const uint8_t* inj_w; const float* inj_u; const uint8_t* loss_mask;
  struct VhProf* prof;
  int *scan_aux, *scan_total;                     // [1024], [1]
  float* aa_code; int *pair_list, *pair_count;    // [2N], [2N], [1]
  f4* grgb;                                       // [N] d L / d rgb of the compacted foreground pixels
  unsigned long long* scan_state;                 // [VH_SCAN_MAX_BLOCKS + 2] look-back words + ticket + finished-block count of the single-launch scan (self-clearing)
  int* pool_tri;                                  // rasterised id per pool_list entry
  int* tex_l0_flag; int n_l0_regions;             // [regions] level-0 flags of the texel-gradient pyramid (texture.cu)
  const float* tex_apply_grad; int tex_gout_persistent;
  float* tex_loss;                                // [2] TV / residual regulariser loss of the texture the last fold (or vhap_tex_reg_loss) saw
  int tex_step_bias;                              // added to the device Adam step inside the texture update (deferred update: -1)
  cudaEvent_t render_wait_ev;                     // one-shot: the next render forward waits for it right before the shading pass   // see vhap_tex_apply_grad / vhap_set_tex_grad_persistent
  unsigned* tex_counter;                          // [1] CTA completion counter of the texture fold kernel (self-resetting)
  cudaStream_t aux[2]; cudaStream_t hp[2]; cudaEvent_t ev[EV_COUNT];   // hp: highest-priority streams for the latency-critical geometry backward
  int tex_fork_pending, no_overlap;   // fork/join of independent kernel chains
  int* dev_step; int use_dev_step;                // device counters [0] Adam step (1-based), [1] global step; used when use_dev_step
  // data-parallel peer exchange of the forward slab over NVLink (CUDA IPC mailboxes, render.cu k_forward_slab / k_finalize): replaces the
  // mid-step NCCL all-gather + its host-side glue kernels on the step's critical chain
  int dp_rank, dp_world; float* dp_box; float** dp_peers_dev; void* dp_peers_host[VH_DP_MAX]; int* dp_epoch; int* dp_err; unsigned long long* dp_wait;   // dp_wait[4]: ns spent waiting for peers (slab, barrier A, barrier B), exchanges
  // peer-memory texture update (dp_tex.cu): caller-allocated symmetric buffers g_rm / ex_rm [T][3][T], their NVSwitch multicast mappings (or NULL)
  unsigned* dp_counter; int dp_tex_forked;
  float *dp_grm, *dp_grm_mc, **dp_grm_peers_dev, *dp_exrm, *dp_exrm_mc, **dp_exrm_peers_dev, *dp_gband, *dp_exband;
  float* dev_lr_scale;                            // [1] learning-rate scale read by the Adam kernels when use_dev_step (ExponentialLR between graph replays)
};

void vh_set_error(vhap_ctx* ctx, const char* what, const char* msg);

// ---- per-kernel launch accounting: every kernel launch goes through LAUNCH(); when profiling is enabled
// (vhap_profile_enable) CUDA events are recorded around each launch on the launching stream so that bench.py can
// report per-kernel device time measured live inside its timed region.
enum { KID_CAM = 0, KID_POSE_FWD, KID_BLEND_FWD, KID_SKIN_FWD, KID_LMK, KID_VNORM, KID_VNORM_BWD, KID_SKIN_BWD, KID_POSE_BWD, KID_JOFF_BWD,
       KID_BLEND_BWD, KID_BETAS_SCATTER, KID_REGS, KID_SNAP, KID_BIN, KID_SCAN, KID_FINE, KID_RAST_OUT, KID_PASSA, KID_POOL_COUNT,
       KID_POOL_SCAN, KID_POOL_SCATTER, KID_AA_PAIRS, KID_PASSB, KID_REDUCE, KID_SLAB, KID_FINALIZE, KID_PASSC1, KID_PASSC, KID_LIGHTS_REDUCE, KID_TEX_L0, KID_MIP,
       KID_TEX_FOLD, KID_TEX_LOSS, KID_ADAM, KID_ASSEMBLE, KID_MISC, KID_COUNT };
#define VH_PROF_SLOTS 128
#define VH_SCAN_MAX_BLOCKS 1024
struct VhProf {
  int on;                                         // 1: events around launches; 2: same, also valid inside stream capture (timeline of a graph replay)
  cudaEvent_t first;
  unsigned long long launches[KID_COUNT];
  int n[KID_COUNT];
  cudaEvent_t ev[KID_COUNT][VH_PROF_SLOTS][2];
};
static inline VhProf* vh_prof(vhap_ctx* c) { return c->prof; }
static inline void vh_prof_record(VhProf* p, cudaEvent_t e, cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (p->on == 2) cudaStreamIsCapturing(s, &st);
  if (st == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(e, s, cudaEventRecordExternal);   // event-record node of the graph
  else cudaEventRecord(e, s);
}
static inline void vh_prof_begin(vhap_ctx* c, int kid, cudaStream_t s) {
  VhProf* p = vh_prof(c);
  p->launches[kid]++;
  if (p->on && p->n[kid] < VH_PROF_SLOTS) { vh_prof_record(p, p->ev[kid][p->n[kid]][0], s); if (!p->first) p->first = p->ev[kid][p->n[kid]][0]; }
}
static inline void vh_prof_end(vhap_ctx* c, int kid, cudaStream_t s) {
  VhProf* p = vh_prof(c);
  if (p->on && p->n[kid] < VH_PROF_SLOTS) { vh_prof_record(p, p->ev[kid][p->n[kid]][1], s); p->n[kid]++; }
}
// Programmatic dependent launch (opt-in, VHAP_B200_PDL=1): every kernel of the engine starts with VH_PDL_SYNC() -- "my dependents may be
// scheduled" followed by "wait until everything before me in the stream has completed and is visible" -- and can be launched with the
// programmatic-stream-serialization attribute, so that the launch latency of kernel N+1 is paid while kernel N still runs (its CTAs sit
// resident at the wait).  Nothing is read or written before the wait: the semantics are those of plain stream order, and without the
// attribute the two instructions are no-ops.  MEASURED (tools/micro/graph_gap.cu, profiles/r02_graph_gap.txt): a graph node costs 1.0-2.7 us
// plain and 0.8-1.9 us with PDL edges; on the step that is 0.7305 vs 0.7318 ms at one GPU (noise) but 0.9035 vs 0.8621 ms at two GPUs --
// the early-resident CTAs of the chain keep the collective kernels of the side streams off the SMs.  Hence OFF by default.
#if defined(__CUDA_ARCH__)
#define VH_PDL_SYNC() do { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)
#else
#define VH_PDL_SYNC() do { } while (0)
#endif
static inline int vh_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("VHAP_B200_PDL"); on = (e && e[0] == '1') ? 1 : 0; }
  return on;
}
template <typename... KArgs, typename... Args>
static inline void vh_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = vh_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
}
#define LAUNCH(c, kid, s, ...) do { vh_prof_begin((c), (kid), (s)); __VA_ARGS__; vh_prof_end((c), (kid), (s)); } while (0)

// kernel-based zero fill of up to 12 buffers in one launch (byte counts multiples of 4).  Unlike a memset node a kernel inherits the
// priority of its stream, so the step's latency-critical chain is not queued behind the CTAs of a concurrently running bulk kernel.
struct VhZeroSegs { int n; void* p[12]; size_t bytes[12]; };
void vh_zero_multi(vhap_ctx* c, const VhZeroSegs& z, cudaStream_t s);
static inline void vh_zero(vhap_ctx* c, void* p, size_t bytes, cudaStream_t s) { VhZeroSegs z; z.n = 1; z.p[0] = p; z.bytes[0] = bytes; vh_zero_multi(c, z, s); }

// flame.cu
void launch_cam_setup(vhap_ctx* c, const vhap_params* p, const vhap_frame_batch* fb, cudaStream_t s);
void launch_flame_forward(vhap_ctx* c, const vhap_params* p, const vhap_frame_batch* fb, cudaStream_t s, bool with_cam = false);   // with_cam: + camera set-up
void launch_landmarks(vhap_ctx* c, const vhap_frame_batch* fb, float w_scale, int jawline_off, float* lmks_out, float* g_lmk_in,
                      int compute_loss, int opt_cam, int global_B, cudaStream_t s);
void launch_flame_backward(vhap_ctx* c, const vhap_params* p, const vhap_frame_batch* fb, const vhap_grads* g, int opt_cam, cudaStream_t s);
void launch_vnormals(vhap_ctx* c, int B, cudaStream_t s);
void launch_vnormals_bwd(vhap_ctx* c, int B, cudaStream_t s);
void launch_tex_reg_loss(vhap_ctx* c, const float* tex_extra, const vhap_stage_cfg* cfg, cudaStream_t s);
void launch_regs(vhap_ctx* c, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg, const vhap_grads* g, int global_B, cudaStream_t s);
// blend_tc.cu
void launch_blend_tc_fwd(vhap_ctx* c, const float* offset, int B, cudaStream_t s);
void launch_blend_tc_bwd(vhap_ctx* c, int B, cudaStream_t s);
// raster.cu
void launch_raster(vhap_ctx* c, const f4* clip, i4* snap, int B, int H, int W, int* tri_id, int cull_backface, int need_snap, cudaStream_t s,
                   bool zeroed = false);      // zeroed: tile_count / tile_cursor already cleared by the caller (the step's clear kernel)
void launch_rast_out(vhap_ctx* c, const f4* clip, int B, int H, int W, const int* tri_id, float* rast, float* rast_db, cudaStream_t s);
void launch_scan(vhap_ctx* c, const int* in, int* out, int n, int* total, cudaStream_t s, int pad4 = 0);
// render.cu
void fill_render_args(vhap_ctx* c, PassArgs& P, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg, const float* lights);
void launch_render_forward(vhap_ctx* c, PassArgs& P, cudaStream_t s, bool zeroed = false);   // zeroed: maxslot / pair_count already cleared by the caller
void launch_render_finalize(vhap_ctx* c, PassArgs& P, const vhap_stage_cfg* cfg, const float* reduce_slab, int global_B, const float* lights, cudaStream_t s);
void launch_render_backward(vhap_ctx* c, PassArgs& P, const vhap_stage_cfg* cfg, const float* lights, float* g_lights, const float* ext_grad, cudaStream_t s,
                            cudaStream_t side = nullptr);
// texture.cu
void launch_tex_rebuild(vhap_ctx* c, const float* tex_extra, cudaStream_t s);
void launch_tex_fold(vhap_ctx* c, float* tex_extra, float* g_out, float* m, float* v, float lr, int step, const vhap_stage_cfg* cfg,
                     float* losses_out, cudaStream_t s);
void launch_tex_fold_grad_rm(vhap_ctx* c, float* tex_extra, float* g_rm, cudaStream_t s);
int launch_tex_band_adam(vhap_ctx* c, float* tex_extra, const float* g_band, int y_begin, int y_end, float* m, float* v, float lr, int step,
                         const vhap_stage_cfg* cfg, float* ex_band_out, cudaStream_t s);
void launch_tex_rebuild_rm(vhap_ctx* c, float* tex_extra, const float* ex_rm, cudaStream_t s);
// dp_tex.cu
void launch_dp_barrier(vhap_ctx* c, int which, cudaStream_t s);
void launch_dp_reduce_band(vhap_ctx* c, float* g_band, cudaStream_t s);
void launch_dp_bcast_band(vhap_ctx* c, const float* ex_band, cudaStream_t s);
void launch_adam(vhap_ctx* c, float* p, const float* g, float* m, float* v, int64_t n, float lr, int step, cudaStream_t s);
