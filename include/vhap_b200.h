/* libvhap_b200.so -- C ABI of the B200-native photometric head-alignment engine.
 *
 * Drop-in boundary for the hot path of ShenhanQian/VHAP.  The reference has no native code of
 * its own: every GPU kernel on this path lives in nvdiffrast / ATen and is reached through Python.  Each entry
 * point below names the reference interface it replaces (file:line relative to /root/reference).  A reference
 * maintainer binds these with ctypes; the in-repo binding is vhap_b200/_lib.py.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller (PyTorch) owns all parameter / gradient / image buffers; the ctx owns static mesh tables and
 *     scratch (allocated in vhap_ctx_create / vhap_ctx_reserve, never in the hot loop);
 *   - all work is enqueued on the cudaStream_t passed as `void* stream`; no hidden synchronisation;
 *   - return 0 on success, negative on error; message via vhap_last_error(); nothing throws across the ABI;
 *   - images inside the engine are in nvdiffrast orientation (row 0 = bottom); `target` is given in image
 *     orientation (row 0 = top) exactly like sample["rgb"] and is flipped on read.
 *   - one ctx per GPU per host thread (not thread-safe), like the reference's single-threaded tracker.
 */
#ifndef VHAP_B200_H
#define VHAP_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct vhap_ctx vhap_ctx;

/* Static FLAME model + topology, HOST pointers, copied to the device at create time.
 * Replaces the buffers FlameHead.__init__ registers (vhap/model/flame.py:70-204) and the per-call topology hash
 * nvdiffrast builds inside dr.antialias (vhap/util/render_nvdiffrast.py:465). */
typedef struct vhap_mesh_desc {
  int32_t V, F, VT, K, n_shape, n_lmk, n_clusters;
  const float*   v_template_host;   /* [V,3]                                   flame.py:99-101 */
  const float*   shapedirs_host;    /* [V,3,K]  (shape cols then expr cols)     flame.py:104-109 */
  const float*   posedirs_host;     /* [36,3V]                                  flame.py:112-114 */
  const float*   J_regressor_host;  /* [5,V]                                    flame.py:116-118 */
  const float*   lbs_weights_host;  /* [V,5]                                    flame.py:122-124 */
  const int32_t* faces_host;        /* [F,3]                                    flame.py:166 */
  const int32_t* faces_uv_host;     /* [F,3]                                    flame.py:169 */
  const float*   verts_uv_host;     /* [VT,2], v already flipped (tracker.py:315-316) */
  const int32_t* lmk_faces_host;    /* [n_lmk]                                  flame.py:131-134 */
  const float*   lmk_bary_host;     /* [n_lmk,3]                                flame.py:135-138 */
  const int32_t* adj_opp_host;      /* [F,3] opposite vertex across edge k, -1 boundary, -2 non-manifold */
  const uint8_t* fid2cid_host;      /* [F+1] cluster of face id (index 0 = background)  render_nvdiffrast.py:77-79 */
  const int32_t* vf_indptr_host;    /* [V+1] vertex -> incident faces (CSR) */
  const int32_t* vf_faces_host;     /* [3F] */
  const int32_t* lap_indptr_host;   /* [V+1] uniform Laplacian CSR incl. diagonal   flame.py:196-201 */
  const int32_t* lap_indices_host;
  const float*   lap_values_host;
} vhap_mesh_desc;

/* One batch of frames.  Parameter arrays are the tracker's full per-sequence tensors (tracker.py:1279-1341);
 * `timesteps` selects the rows used by this batch (sample["timestep_index"], tracker.py:708). */
typedef struct vhap_params {
  const float* shape;          /* [n_shape] */
  const float* expr;           /* [N_t, K-n_shape] */
  const float* rotation;       /* [N_t,3] */
  const float* neck_pose;      /* [N_t,3] */
  const float* jaw_pose;       /* [N_t,3] */
  const float* eyes_pose;      /* [N_t,6] */
  const float* translation;    /* [N_t,3] */
  const float* static_offset;  /* [V,3] or NULL */
  const float* lights;         /* [9,3] */
  const float* focal_length;   /* [1] (uncalibrated, tracker.py:1333) or NULL when K is given per frame */
  const float* tex_extra;      /* [3,T,T] residual texture (tracker.py:1296-1298) */
  int32_t n_timesteps;
} vhap_params;

/* Gradient outputs, same shapes as vhap_params (dense rows like autograd's, tracker.py:1434).  Any pointer may be
 * NULL = that parameter is not optimised in this stage (tracker.py:1465-1513). */
typedef struct vhap_grads {
  float* shape; float* expr; float* rotation; float* neck_pose; float* jaw_pose; float* eyes_pose;
  float* translation; float* static_offset; float* lights; float* focal_length;
  float* tex_grad_pyramid;     /* engine-owned layout, obtain with vhap_tex_grad_ptr(); NULL if texture frozen */
} vhap_grads;

typedef struct vhap_frame_batch {
  int32_t B, H, W;
  int32_t target_format;       /* 0: target is [B,H,W,4] fp16 RGBA; 1: [B,H,W,3] uint8 RGB exactly as the dataset decodes it -- the
                                  kernels divide by 255 in fp32 like F.to_tensor (video_dataset.py:256-260); 3 B/px instead of 8 on the wire */
  const int32_t* timesteps;    /* [B] device */
  const void*    target;       /* image orientation (row 0 top); sample["rgb"] (tracker.py:405) */
  const float*   lmk2d;        /* [B,68,3] (x_px,y_px,conf)  sample["lmk2d"] (tracker.py:358) */
  const float*   RT;           /* [B,3,4] or NULL = [I|(0,0,-1)] (tracker.py:1335-1337) */
  const float*   K;            /* [B,4]=(fx,fy,cx,cy) or NULL = from focal_length (tracker.py:141-157) */
  /* optional view sharing (calibrated multi-view batches: all cameras of one timestep carry the SAME FLAME parameters, tracker.py:213-235
   * evaluates FLAME once per frame anyway): geometry is evaluated once per distinct timestep and only projected per view.
   * geo[b] = index of frame b's timestep in geo_timesteps[n_geo]; NULL = every frame its own geometry. */
  const int32_t* geo;          /* [B] device or NULL */
  const int32_t* geo_timesteps;/* [n_geo] device */
  int32_t n_geo;
} vhap_frame_batch;

/* Loss weights and stage switches (vhap/config/base.py:126-195, :215-295).  A negative weight = term disabled (None). */
typedef struct vhap_stage_cfg {
  float w_landmark, w_photo, w_reg_shape, w_reg_expr, w_reg_neck, w_reg_jaw, w_reg_eyes;
  float w_reg_tex_tv, w_reg_tex_res, w_reg_diffuse, w_reg_light;
  float w_reg_offset, w_reg_offset_lap, w_reg_offset_rigid;
  float w_smooth_trans, w_smooth_rot, w_smooth_neck, w_smooth_jaw, w_smooth_eyes, w_smooth_expr;
  int32_t photometric;          /* PhotometricStageConfig? (tracker.py:722) */
  int32_t jawline_off;          /* use landmarks 17:68 only (tracker.py:371-373) */
  int32_t tracking;             /* 'tracking' in stage name: temporal smoothness on (tracker.py:493,502,511) */
  int32_t training;             /* 0 = evaluation mode (stage None): no disturbance, no regularisers (tracker.py:427,741) */
  int32_t opt_pose, opt_joints, opt_expr, opt_shape, opt_texture, opt_lights, opt_static_offset, opt_cam;
  int32_t bg_mode;              /* 0 = target image, 1 = constant colour bg_color (tracker.py:287-303) */
  float   bg_color[3];
  float   disturb_rate_fg, disturb_rate_bg;   /* <0 = None (render_nvdiffrast.py:428-435) */
  uint64_t rng_seed; uint64_t rng_step;        /* Philox counter base for the in-kernel disturbance */
  float   shared_scale;         /* 1/world_size: batch-independent terms (shape/light/offset/texture regularisers) are scaled by this
                                   on every rank so that the data-parallel sum-allreduce reproduces them exactly once */
} vhap_stage_cfg;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int  vhap_ctx_create(vhap_ctx** out, const vhap_mesh_desc* mesh, int32_t tex_size, int32_t device);
int  vhap_ctx_reserve(vhap_ctx* ctx, int32_t max_B, int32_t max_H, int32_t max_W);   /* scratch for a batch shape */
void vhap_ctx_destroy(vhap_ctx* ctx);
const char* vhap_last_error(const vhap_ctx* ctx);
int  vhap_abi_version(void);

/* per-stage masks: faces whose texture coordinate is detached (align_texture_except_fid, render_nvdiffrast.py:390-396)
 * and vertices whose clip position is detached inside antialias (align_boundary_except_vid, :349-352,:463-464);
 * per-vertex weights of reg_offset / reg_offset_lap (tracker.py:564-587, :607-614) and rigid-region ids (:589-594);
 * uv mask of reg_tex_res_clusters (tracker.py:536-539).  HOST pointers, copied. */
int vhap_set_stage_masks(vhap_ctx* ctx, const uint8_t* face_tex_detach_host /*[F]*/, const uint8_t* vert_aa_detach_host /*[V]*/,
                         const float* w_offset_host /*[V]*/, const float* w_offset_lap_host /*[V]*/,
                         const int32_t* rigid_indptr_host /*[n_rigid+1]*/, const int32_t* rigid_vids_host, int32_t n_rigid_regions,
                         const uint8_t* uvmask_res_host /*[T,T] 0/1*/);

/* ---- FLAME (replaces FlameHead.forward, vhap/model/flame.py:571-646 + vhap/model/lbs.py) ------------------ */
int vhap_flame_forward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb,
                       float* verts /*[B,V,3]*/, float* verts_cano /*[B,V,3] or NULL*/, float* lmks /*[B,n_lmk,3]*/,
                       void* stream);
/* backward of the above for upstream gradients g_verts [B,V,3], g_lmks [B,n_lmk,3] (either may be NULL) */
int vhap_flame_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb,
                        const float* g_verts, const float* g_lmks, const vhap_grads* g, void* stream);

/* ---- camera (replaces NVDiffRenderer.world_to_camera / camera_to_clip, render_nvdiffrast.py:162-197) -------- */
int vhap_project(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const float* verts /*[B,V,3]*/,
                 float* verts_clip /*[B,V,4]*/, void* stream);

/* adjoint of vhap_project: g_verts [B,V,3] += J^T g_clip, g_focal[0] += d/d focal_length (NULL to skip) */
int vhap_project_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const float* verts, const float* g_clip /*[B,V,4]*/,
                          float* g_verts, float* g_focal, void* stream);

/* ---- vertex normals (replaces NVDiffRenderer.compute_v_normals, render_nvdiffrast.py:297-316) ---------------- */
int vhap_vertex_normals(vhap_ctx* ctx, const float* verts /*[B,V,3]*/, int32_t B, float* vnorm /*[B,V,3]*/, void* stream);
int vhap_vertex_normals_backward(vhap_ctx* ctx, const float* verts, const float* g_vnorm /*[B,V,3]*/, int32_t B, float* g_verts /*[B,V,3], +=*/,
                                 void* stream);

/* ---- rasterise (replaces dr.rasterize at render_nvdiffrast.py:254) ---------------------------------------- */
/* tri_id [B,H,W] int32 (triangle+1, 0 empty, row 0 = bottom); rast/rast_db [B,H,W,4] float (NULL to skip). */
int vhap_rasterize(vhap_ctx* ctx, const float* verts_clip /*[B,V,4]*/, int32_t B, int32_t H, int32_t W,
                   int32_t* tri_id, float* rast, float* rast_db, int32_t cull_backface, void* stream);

/* ---- fused photometric energy (replaces FlameTracker.compute_energy's FLAME + landmark + rasterise + render_rgba
 *      + photometric + regulariser section, tracker.py:692-750, and E_total.backward(), tracker.py:1434) ------- */
/* losses_out: device float[VHAP_N_LOSS] filled per the enum below (already weighted like log_dict). */
enum { VHAP_L_TOTAL = 0, VHAP_L_LMK, VHAP_L_PHOTO, VHAP_L_REG_SHAPE, VHAP_L_REG_EXPR, VHAP_L_REG_JOINT, VHAP_L_SMOOTH_POSE,
       VHAP_L_SMOOTH_JOINT, VHAP_L_SMOOTH_EXPR, VHAP_L_REG_TEX_TV, VHAP_L_REG_TEX_RES, VHAP_L_REG_DIFFUSE, VHAP_L_REG_LIGHT,
       VHAP_L_REG_OFFSET, VHAP_L_REG_OFFSET_LAP, VHAP_L_REG_OFFSET_RIGID, VHAP_L_NFG, VHAP_L_ABSERR, VHAP_N_LOSS = 24 };
int vhap_energy_forward_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg,
                                 const vhap_grads* g /* NULL = forward only */, float* losses_out, void* stream);
/* Two-phase variant for data-parallel runs: phase 1 runs everything whose result feeds a cross-rank reduction
 * (forward + per-rank partial sums in reduce_slab: [sum|err|, n_fg, diffuse max ...]); phase 2 consumes the reduced
 * slab and runs the backward.  See DESIGN.md "Multi-GPU". */
int vhap_energy_forward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg,
                        float* reduce_slab /* device float[8] */, void* stream);
int vhap_energy_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg,
                         const float* reduce_slab /* cross-rank reduced */, const float* local_slab /* this rank's own */,
                         int32_t global_B, const vhap_grads* g, float* losses_out, void* stream);

/* modular form: rasterise + render_rgba + photometric energy (+reg_diffuse) from caller-provided clip positions and vertex
 * normals, and the analytic backward to them (NVDiffRenderer.rasterize/render_rgba, render_nvdiffrast.py:216-245,354-484;
 * compute_photometric_energy, tracker.py:391-478).  Gradient outputs are overwritten; any may be NULL. */
int vhap_render_photometric(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg,
                            const float* verts_clip /*[B,V,4]*/, const float* vnorm /*[B,V,3]*/, float* losses_out,
                            float* g_clip /*[B,V,4]*/, float* g_vnorm /*[B,V,3]*/, float* g_lights /*[27]*/, float* g_tex_pyramid, void* stream);

/* adjoint of the 'rgba' output of the last vhap_render_photometric forward for an external upstream gradient g_rgba
 * [B,H,W,4] (image orientation): lets NVDiffRenderer.render_rgba be used as a torch.autograd.Function */
int vhap_render_rgba_backward(vhap_ctx* ctx, const vhap_params* p, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg,
                              const float* g_rgba, float* g_clip, float* g_vnorm, float* g_lights, float* g_tex_pyramid, void* stream);

/* debug / logging planes of the last forward (render_out dict, render_nvdiffrast.py:476-483), image orientation.
 * which: 0 rgba (after AA), 1 rgba before AA, 2 albedo, 3 normal, 4 diffuse, 5 cid.  out [B,H,W,4] float. */
int vhap_get_plane(vhap_ctx* ctx, int32_t which, float* out, void* stream);
/* the engine's own geometry of the last forward: which = 0 world vertices, 1 clip positions, 2 vertex normals; out [B,V,4] */
int vhap_get_geometry(vhap_ctx* ctx, int32_t which, float* out, void* stream);
int vhap_set_want_planes(vhap_ctx* ctx, int32_t on);   /* make the next forward keep the logging planes */
int vhap_overflow_flag(vhap_ctx* ctx, int32_t* out_host);   /* 1 if a raster tile list overflowed its capacity (synchronises) */
/* test hook: inject the disturbance randomness (w bits: bit0 = w_fg, bit1 = w_bg; u in [0,1)); NULL = Philox. */
int vhap_set_injected_random(vhap_ctx* ctx, const uint8_t* w_bits /*[B,H,W]*/, const float* u /*[B,H,W]*/);
/* test hook: per-pixel switch of the L1 photometric term of the next forwards ([B,H,W] device bytes, image orientation like
 * `target`; 0 = pixel left out of sum|gt - pred|, the normaliser still counts it); NULL = every pixel (tracker.py:438-439). */
int vhap_set_loss_mask(vhap_ctx* ctx, const uint8_t* mask);

/* ---- texture: pyramid rebuild + regularisers + Adam (tracker.py:247-258 get_albedo, :526-539, torch.optim.Adam) -- */
float* vhap_tex_grad_ptr(vhap_ctx* ctx);                 /* device pointer of the texture-gradient pyramid */
int vhap_set_tex_painted(vhap_ctx* ctx, const float* tex_painted /*[3,T,T] device*/, void* stream);
int vhap_tex_rebuild(vhap_ctx* ctx, const float* tex_extra /*[3,T,T]*/, void* stream);   /* level 0 + mips */
/* folds the pyramid gradient to level 0, adds TV + residual-cluster gradients, writes dense grad [3,T,T] (g_out) or,
 * if adam_m/adam_v given, applies one Adam step in place to tex_extra and rebuilds the pyramid in the same pass. */
int vhap_tex_reg_fold_adam(vhap_ctx* ctx, float* tex_extra, float* g_out, float* adam_m, float* adam_v,
                           float lr, int32_t step, const vhap_stage_cfg* cfg, float photo_scale, float* losses_out, void* stream);
/* data parallel: Adam + pyramid rebuild from the dense gradient produced by vhap_tex_reg_fold_adam(g_out) and summed across ranks */
int vhap_tex_apply_grad(vhap_ctx* ctx, float* tex_extra, const float* g_dense, float* adam_m, float* adam_v, float lr, int32_t step,
                        const vhap_stage_cfg* cfg, void* stream);
/* ---- data-parallel exchange of the step's batch-global scalars over peer memory (one node, CUDA IPC; the reference has no multi-GPU path).
 * vhap_dp_init: allocate this rank's mailbox, return its 64-byte IPC handle; the caller all-gathers the handles of all ranks and passes
 * them to vhap_dp_connect.  Afterwards vhap_energy_forward pushes this rank's [sum|err|, n_fg, sum var, max diffuse] into every peer's
 * mailbox and vhap_energy_backward gathers them -- no collective between the halves of a step (their reduce_slab arguments are ignored).
 * All ranks must then call vhap_energy_forward / vhap_energy_backward the same number of times. */
int vhap_dp_init(vhap_ctx* ctx, int32_t rank, int32_t world, unsigned char* handle_out_host /*64 bytes*/);
int vhap_dp_connect(vhap_ctx* ctx, const unsigned char* handles_host /*[world][64]*/);
int vhap_dp_status(vhap_ctx* ctx, int32_t* out_host);   /* 0 ok, 1 / 2 = a peer did not answer within ~4 s (slab / texture barrier; synchronises) */
/* diagnostics: nanoseconds this rank waited for its peers since the last reset -- out[0] slab exchange, out[1] / out[2] texture barriers A / B,
 * out[3] = number of slab exchanges (synchronises) */
int vhap_dp_wait_stats(vhap_ctx* ctx, uint64_t* out_host /*[4]*/, int32_t reset);
/* Peer-memory texture update: fold -> barrier -> in-switch reduction of this rank's row band (multimem.ld_reduce through the NVSwitch
 * multicast mapping; peer loads without one) -> regularisers + Adam on the band -> multicast store of the updated rows to every rank ->
 * barrier -> pyramid rebuild; no collective library.  g_rm / ex_rm ([T][3][T] floats) are SYMMETRIC buffers allocated by the caller on
 * every rank: *_ptrs_host = every rank's device pointer (HOST array of `world`), *_multicast = the multicast address or NULL. */
int vhap_dp_tex_connect(vhap_ctx* ctx, void* const* grm_ptrs_host, void* grm_multicast, void* const* exrm_ptrs_host, void* exrm_multicast);
int vhap_dp_tex_update(vhap_ctx* ctx, float* tex_extra, float* adam_m, float* adam_v, float lr, int32_t step, const vhap_stage_cfg* cfg, void* stream);
/* the same in two parts for pipelined steps: part 1 (fold, barrier, band reduction) right after vhap_energy_backward -- it runs on the
 * library's bulk stream beside the geometry backward, vhap_dp_tex_join makes `stream` wait for it --, part 2 (band Adam, multicast store,
 * barrier, pyramid rebuild) whenever the caller wants the update applied (e.g. at the start of the next step) */
int vhap_dp_tex_part1(vhap_ctx* ctx, float* tex_extra, void* stream);
int vhap_dp_tex_join(vhap_ctx* ctx, void* stream);
int vhap_dp_tex_part2(vhap_ctx* ctx, float* tex_extra, float* adam_m, float* adam_v, float lr, int32_t step, const vhap_stage_cfg* cfg, void* stream);

/* ---- sharded texture update (data parallel; the reference has no multi-GPU path): per step
 *   vhap_tex_fold_grad_rm  photometric part of the texel gradient, dense, ROW-MAJOR g_rm[(y*3 + c)*T + x] so that a row band is contiguous
 *   (caller)               reduce-scatter of g_rm by row band over the ranks
 *   vhap_tex_band_adam     rows [y_begin, y_end) of this rank: + TV / residual gradients (rank-invariant: computed once, by the owner), Adam
 *                          on this band's rows of tex_extra / adam_m / adam_v only, updated rows also to ex_band_out[((y-y_begin)*3+c)*T+x]
 *   (caller)               all-gather of the bands into ex_rm[(y*3 + c)*T + x]
 *   vhap_tex_rebuild_rm    planar tex_extra + level 0 / 1 + mips of the new pyramid from ex_rm (get_albedo, tracker.py:247-258)
 * y_begin and the band height must be multiples of 8. */
int vhap_tex_fold_grad_rm(vhap_ctx* ctx, float* tex_extra, float* g_rm, void* stream);
int vhap_tex_band_adam(vhap_ctx* ctx, float* tex_extra, const float* g_band, int32_t y_begin, int32_t y_end, float* adam_m, float* adam_v,
                       float lr, int32_t step, const vhap_stage_cfg* cfg, float* ex_band_out, void* stream);
int vhap_tex_rebuild_rm(vhap_ctx* ctx, float* tex_extra, const float* ex_rm, void* stream);
/* on != 0: g_out of vhap_tex_reg_fold_adam is a persistent buffer only read after the call returns -> the fold may overlap the
 * geometry backward on the library's aux stream */
int vhap_set_tex_grad_persistent(vhap_ctx* ctx, int32_t on);
/* ---- deferred texture update: the update of step k executes at the start of step k+1 on a caller stream, beside the kernels of
 *      the next forward that do not read the texture, and is joined right before the shading pass --------------------------- */
int vhap_tex_defer(vhap_ctx* ctx, int32_t step_bias);           /* next texture-update call: no in-call fork; device Adam step + step_bias */
int vhap_tex_reg_loss(vhap_ctx* ctx, const float* tex_extra, const vhap_stage_cfg* cfg, void* stream);   /* regulariser loss values of the current texture */
int vhap_assemble_losses(vhap_ctx* ctx, const vhap_stage_cfg* cfg, float* losses_out, void* stream);   /* loss vector from the accumulators + vhap_tex_reg_loss values */
int vhap_set_render_wait_event(vhap_ctx* ctx, void* cuda_event);   /* one-shot wait inserted before the first texture read of the next forward */

/* ---- per-kernel accounting: every kernel launch is counted; with profiling enabled CUDA events bracket each launch on the
 *      launching stream (used by bench.py to measure the dominant kernel live inside its timed region) --------------- */
int vhap_profile_enable(vhap_ctx* ctx, int32_t on);            /* 1: on (resets counters); 2: on, and usable inside stream capture (events become
                                                                  nodes of the graph); 0: off + reset; -1: off, keep the recorded slots */
int vhap_set_overlap(vhap_ctx* ctx, int32_t on);               /* 0: no aux-stream fork/join (clean per-kernel timing); default 1 */
int vhap_profile_kernel_count(void);
const char* vhap_profile_kernel_name(int32_t kid);
int vhap_profile_read(vhap_ctx* ctx, float* avg_ms_host, uint64_t* launches_host);   /* arrays of vhap_profile_kernel_count(); synchronises */
/* start/end of every recorded launch in ms after the first recorded event (timeline of a graph replay in mode 2); returns n */
int vhap_profile_timeline(vhap_ctx* ctx, int32_t* kid_out, float* t0_ms, float* t1_ms, int32_t max_n);

/* ---- CUDA-graph support: with device-resident step counters no kernel argument varies from step to step, so a whole step
 *      (forward, backward, regularisers, Adam) can be captured once per (batch, texture ping-pong parity) and replayed ------ */
int vhap_step_counters(vhap_ctx* ctx, int32_t on, int32_t adam_step, int32_t global_step, void* stream);
int vhap_step_advance(vhap_ctx* ctx, void* stream);
/* learning-rate scale of the Adam kernels while the device step counters are on (torch's ExponentialLR stepped between epochs,
 * tracker.py:1407-1412): captured step graphs read it from device memory, so a replayed stage follows the reference's schedule */
int vhap_set_lr_scale(vhap_ctx* ctx, float scale, void* stream);
int vhap_get_cur_mip(vhap_ctx* ctx);
int vhap_set_cur_mip(vhap_ctx* ctx, int32_t v);

/* ---- fused Adam on small parameter slabs (torch.optim.Adam, tracker.py:159-211) ------------------------------ */
/* all parameter groups of one slab in a single launch: segment k = [off[k], off[k]+len[k]) with learning rate lr[k] (HOST arrays, <= 24) */
int vhap_adam_multi(vhap_ctx* ctx, float* param, const float* grad, float* m, float* v, int32_t n_seg, const int64_t* off_host,
                    const int64_t* len_host, const float* lr_host, int32_t step, void* stream);
int vhap_adam(vhap_ctx* ctx, float* param, const float* grad, float* m, float* v, int64_t n, float lr, int32_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif
