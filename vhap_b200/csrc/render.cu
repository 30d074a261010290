// Fused render passes (kernels around render_bodies.cuh).  Order per forward:
//   k_pool_count / k_pool_scan / k_pool_scatter   stable compaction of pixel indices per colour cluster (the pools of
//                     render_nvdiffrast.py:445-459); clusters >= 1 together form the compacted FOREGROUND pixel list
//   k_passA      shade every foreground pixel (walks the compacted list), write its composite colour, reg_diffuse partials
//   k_passB      disturbance + antialias + L1 partial sums + sign bytes, all pixels
//   k_forward_slab / k_finalize   global scalars (photo scale = w/(3 n_fg), reg_diffuse scale / arg-max), loss values
//   k_passC      analytic backward over the compacted foreground list -> clip positions, vertex normals, texels, lights
// All per-pixel kernels are persistent grid-stride kernels sized to the machine (148 SMs x 8 CTAs of 256 threads).
// Replaces NVDiffRenderer.render_rgba (vhap/util/render_nvdiffrast.py:354-484), compute_photometric_energy
// (vhap/model/tracker.py:391-478), reg_diffuse (tracker.py:547-550) and their autograd.
#include "engine.h"
#include "accum.h"

#define PB 256                 // threads per block in the per-pixel passes
#define NPERSIST (148 * 8)     // persistent grid: one wave of 8 resident CTAs per SM

__device__ __forceinline__ float warp_sum_r(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// reduce N per-thread values across the block and write them to row[0..N); sh must hold 8*N floats (256 threads)
template <int N, int NT = PB>
__device__ void block_reduce_store(float* vals, float* sh, float* row) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) { float v = warp_sum_r(vals[i]); if (lane == 0) sh[w * N + i] = v; }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < NT / 32; ++k) s += sh[k * N + i];
    row[i] = s;
  }
}

__device__ __forceinline__ unsigned long long pack_max(float v, int idx) {
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);          // order-preserving map float -> uint
  return ((unsigned long long)u << 32) | (unsigned)idx;
}
__device__ __forceinline__ float unpack_max_val(unsigned long long p) {
  unsigned u = (unsigned)(p >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------- cluster pools
// Each CTA covers PB * POOL_PPT consecutive pixels; warp w owns the 32 * POOL_PPT pixels [w * 32 * POOL_PPT, ...) of that range as POOL_PPT
// coalesced rows of 32, so list order = pixel order (the pools must be in torch's boolean-mask order: the reference draws a pool INDEX).
// The POOL_PPT id loads of a thread are independent and issued together (one pixel per thread was latency-bound: 0.051 ms for 50 MB).
#define POOL_PPT 4
__global__ void __launch_bounds__(PB) k_pool_count(const int* __restrict__ tri_id, const uint8_t* __restrict__ fid2cid, size_t n, int ncl, int* __restrict__ blk_count) { VH_PDL_SYNC();
  __shared__ int cnt[16];
  if (threadIdx.x < 16) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t base = ((size_t)blockIdx.x * (PB / 32) + w) * (32 * POOL_PPT) + lane;
  int id[POOL_PPT], cid[POOL_PPT];
#pragma unroll
  for (int j = 0; j < POOL_PPT; ++j) id[j] = base + 32 * j < n ? tri_id[base + 32 * j] : -1;
#pragma unroll
  for (int j = 0; j < POOL_PPT; ++j) cid[j] = id[j] >= 0 ? fid2cid[id[j]] : -1;
#pragma unroll
  for (int j = 0; j < POOL_PPT; ++j) {                           // one shared-memory add per (row, cluster present in the row)
    const unsigned grp = __match_any_sync(0xffffffffu, cid[j]);
    if (lane == __ffs(grp) - 1 && cid[j] >= 0) atomicAdd(&cnt[cid[j]], __popc(grp));
  }
  __syncthreads();
  if (threadIdx.x < 16) blk_count[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = cnt[threadIdx.x];    // [16][nblk]
}

// also emits the list of adjacent pixel pairs with different ids (pair = pixel0 * 2 + direction) for the antialias analysis
__global__ void __launch_bounds__(PB) k_pool_scatter(const int* __restrict__ tri_id, const uint8_t* __restrict__ fid2cid, size_t n, int H, int W, int ncl,
                                                     const int* __restrict__ blk_off, int* __restrict__ pool_list, int* __restrict__ pool_tri,
                                                     int* __restrict__ pair_list, int* __restrict__ pair_count,
                                                     const int* __restrict__ total, int* __restrict__ base, int* __restrict__ count) { VH_PDL_SYNC();
  __shared__ int wcnt[16][PB / 32];
  // base / count of every cluster from the scanned [16][nblk] offsets (lists laid out cluster after cluster): read by the passes that
  // follow this kernel, written here by the first CTA (was a launch of its own)
  if (blockIdx.x == 0 && threadIdx.x < 16) {
    const int c = threadIdx.x, nb = gridDim.x;
    const int b = blk_off[(size_t)c * nb], e = c < 15 ? blk_off[(size_t)(c + 1) * nb] : *total;
    base[c] = b; count[c] = e - b;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned below = (1u << lane) - 1;
  const size_t pix0 = ((size_t)blockIdx.x * (PB / 32) + w) * (32 * POOL_PPT) + lane;
  int id[POOL_PPT], cid[POOL_PPT], idr[POOL_PPT], idd[POOL_PPT];
#pragma unroll
  for (int j = 0; j < POOL_PPT; ++j) {
    const size_t pix = pix0 + 32 * j;
    id[j] = pix < n ? tri_id[pix] : -1;
    const int x = pix % W, y = (pix / W) % H;
    idr[j] = pix < n && x + 1 < W ? tri_id[pix + 1] : id[j];
    idd[j] = pix < n && y + 1 < H ? tri_id[pix + W] : id[j];
  }
#pragma unroll
  for (int j = 0; j < POOL_PPT; ++j) cid[j] = id[j] >= 0 ? fid2cid[id[j]] : -1;
  // antialias pairs: ONE reservation on the global counter per CTA (per-warp reservations were thousands of same-address atomics)
  __shared__ int wpair[PB / 32];
  __shared__ int s_pair_base;
  unsigned m0[POOL_PPT], m1[POOL_PPT];
  int tot = 0;
#pragma unroll
  for (int j = 0; j < POOL_PPT; ++j) {
    m0[j] = __ballot_sync(0xffffffffu, idr[j] != id[j]); m1[j] = __ballot_sync(0xffffffffu, idd[j] != id[j]);
    tot += __popc(m0[j]) + __popc(m1[j]);
  }
  if (lane == 0) wpair[w] = tot;
  // cluster ranks: lanes of one cluster find each other with match.any; the running per-warp count of a cluster lives in shared memory
  // (rows are processed in order, the group's lowest lane updates it), so rank = members in earlier rows + members in lower lanes
  for (int c = lane; c < 16; c += 32) wcnt[c][w] = 0;
  __syncwarp();
  int rank[POOL_PPT];
#pragma unroll
  for (int j = 0; j < POOL_PPT; ++j) {
    const unsigned grp = __match_any_sync(0xffffffffu, cid[j]);
    const int leader = __ffs(grp) - 1;
    int prev = 0;
    if (lane == leader && cid[j] >= 0) { prev = wcnt[cid[j]][w]; wcnt[cid[j]][w] = prev + __popc(grp); }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[j] = prev + __popc(grp & below);
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int sum = 0;
    for (int k = 0; k < PB / 32; ++k) sum += wpair[k];
    s_pair_base = sum ? atomicAdd(pair_count, sum) : 0;
  }
  __syncthreads();
  if (tot) {
    int at = s_pair_base;
    for (int k = 0; k < w; ++k) at += wpair[k];
#pragma unroll
    for (int j = 0; j < POOL_PPT; ++j) {
      const int pix = (int)(pix0 + 32 * j);
      if (idr[j] != id[j]) pair_list[at + __popc(m0[j] & below)] = pix * 2;
      at += __popc(m0[j]);
      if (idd[j] != id[j]) pair_list[at + __popc(m1[j] & below)] = pix * 2 + 1;
      at += __popc(m1[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < POOL_PPT; ++j)
    if (cid[j] >= 0) {
      int before = 0;
      for (int k = 0; k < w; ++k) before += wcnt[cid[j]][k];
      const int pos = blk_off[(size_t)cid[j] * gridDim.x + blockIdx.x] + before + rank[j];
      pool_list[pos] = (int)(pix0 + 32 * j);
      pool_tri[pos] = id[j];
    }
}

// ---------------------------------------------------------------------------------------------- passes
__global__ void __launch_bounds__(PB) k_aa_pairs(PassArgs P, const int* __restrict__ pair_list, const int* __restrict__ pair_count) { VH_PDL_SYNC();
  const RenderArgs& A = P.R;
  int n = *pair_count;
  for (int i = blockIdx.x * PB + threadIdx.x; i < n; i += gridDim.x * PB) {
    int pr = pair_list[i], pix = pr >> 1, d = pr & 1;
    int x, y, b; vh_unflatten(A, pix, b, y, x);
    aa_pair_body(P, b, y, x, d);
  }
}

// resident CTAs per SM the register allocation is tuned for (measured on B200, tools/ab_variants.sh): these passes are bound by
// the latency of dependent gathers, so pass B trades a few spilled registers for full occupancy (0.117 -> 0.093 ms), the colour
// adjoint runs best at 4 CTAs (64 registers), the shading adjoint at 2 (128 registers; 3 was 25% slower), pass A unconstrained.
#ifndef VH_B_MIN
#define VH_B_MIN 8
#endif
#ifndef VH_C1_MIN
#define VH_C1_MIN 4
#endif
#ifdef VH_A_MIN
__global__ void __launch_bounds__(PB, VH_A_MIN) k_passA(
#else
__global__ void __launch_bounds__(PB) k_passA(
#endif
    PassArgs P, float* __restrict__ partials, unsigned long long* __restrict__ maxslot) { VH_PDL_SYNC();
  __shared__ float sh[8 * 2];
  __shared__ unsigned long long shm[8];
  const RenderArgs& A = P.R;
  int n_fg = A.B * A.H * A.W - P.pool_count[0];
  const int* list = P.pool_list + P.pool_base[1];
  const int* tris = P.pool_tri + P.pool_base[1];
  float acc[2] = {0.f, 0.f}; float mx = -INFINITY; int mxi = 0;
  for (int i = blockIdx.x * PB + threadIdx.x; i < n_fg; i += gridDim.x * PB) {
    int pix = list[i], id = tris[i];
    int x, y, b; vh_unflatten(A, pix, b, y, x);
    passA_body(P, b, y, x, acc, mx, mxi, id);
  }
  block_reduce_store<2>(acc, sh, partials + (size_t)blockIdx.x * VH_NPART);
  unsigned long long pm = mx > -INFINITY ? pack_max(mx, mxi) : 0ull;
  for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, pm, o); pm = t > pm ? t : pm; }
  if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = pm;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < PB / 32; ++k) pm = shm[k] > pm ? shm[k] : pm;
    if (pm) atomicMax(maxslot, pm);
  }
}

__global__ void __launch_bounds__(PB, VH_B_MIN) k_passB(PassArgs P, float* __restrict__ partials) { VH_PDL_SYNC();
  __shared__ float sh[8 * 2];
  const RenderArgs& A = P.R;
  int n = A.B * A.H * A.W;
  float acc[2] = {0.f, 0.f};
  for (int pix = blockIdx.x * PB + threadIdx.x; pix < n; pix += gridDim.x * PB) {
    int x, y, b; vh_unflatten(A, pix, b, y, x);
    passB_body(P, b, y, x, acc);
  }
  block_reduce_store<2>(acc, sh, partials + (size_t)blockIdx.x * VH_NPART + 2);
}

__global__ void __launch_bounds__(PB, VH_C1_MIN) k_passC1(PassArgs P, const float* __restrict__ ext_grad, f4* __restrict__ grgb) { VH_PDL_SYNC();
  const RenderArgs& A = P.R;
  int n_fg = A.B * A.H * A.W - P.pool_count[0];
  const int* list = P.pool_list + P.pool_base[1];
  const int* tris = P.pool_tri + P.pool_base[1];
  for (int i = blockIdx.x * PB + threadIdx.x; i < n_fg; i += gridDim.x * PB) {
    int pix = list[i], id = tris[i];
    int x, y, b; vh_unflatten(A, pix, b, y, x);
    f3 g = passC1_body(P, b, y, x, ext_grad, id);
    f4 o = {g.x, g.y, g.z, 0.f};
    grgb[i] = o;                                    // indexed by list position: coalesced
  }
}

#ifndef VH_C2_MINBLOCKS
#define VH_C2_MINBLOCKS 4      // x 128 threads = 128 registers per thread; 3 x 256 (85 registers, spills) measured 0.156 vs 0.117 ms (r14)
#endif
#ifndef VH_C2_PB
#define VH_C2_PB 128           // 128-thread CTAs: 0.115 vs 0.117 ms with 256 (r14), finer tail
#endif
#ifndef VH_C2_SPLIT
#define VH_C2_SPLIT 0          // 1: texel-gradient half of pass C2 as a concurrent kernel on a second stream -- measured SLOWER (0.757 vs 0.735 ms/step, r6)
#endif
#ifndef VH_C2_AGG
#define VH_C2_AGG 1            // warp-level reduction of the per-vertex gradients over the pixels of one triangle (0: one reduction per pixel)
#endif
// texel-gradient half of the shading adjoint (see shade_pixel_texgrad): a light kernel that only issues the 8 vector reductions per pixel
__global__ void __launch_bounds__(PB, 4) k_passC2_tex(PassArgs P, const f4* __restrict__ grgb) { VH_PDL_SYNC();
  const RenderArgs& A = P.R;
  int n_fg = A.B * A.H * A.W - P.pool_count[0];
  const int* list = P.pool_list + P.pool_base[1];
  const int* tris = P.pool_tri + P.pool_base[1];
  for (int i = blockIdx.x * PB + threadIdx.x; i < n_fg; i += gridDim.x * PB) {
    int pix = list[i], id = tris[i];
    int x, y, b; vh_unflatten(A, pix, b, y, x);
    f4 g = grgb[i];
    if (id > 0) shade_pixel_texgrad(A, b, x, y, id - 1, mk3(g.x, g.y, g.z), P.g_tex);
  }
}

__global__ void __launch_bounds__(VH_C2_PB, VH_C2_MINBLOCKS) k_passC2(PassArgs P, const f4* __restrict__ grgb, float* __restrict__ partials) { VH_PDL_SYNC();
  __shared__ float sh[(VH_C2_PB / 32) * 27];
  const RenderArgs& A = P.R;
  int n_fg = A.B * A.H * A.W - P.pool_count[0];
  const int* list = P.pool_list + P.pool_base[1];
  float gl[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) gl[i] = 0.f;
  const int* tris = P.pool_tri + P.pool_base[1];
#if VH_C2_AGG
  // The foreground list is in pixel order, so the pixels of one triangle on a scan line are neighbouring lanes: their per-vertex
  // gradients are summed with a segmented warp reduction (suffix sums over the run of equal (frame, triangle) keys, 5 shuffle steps)
  // and leave as ONE vector reduction per vertex per run instead of one per pixel (the L1 reduction path was 65 % busy, ncu r02 base).
  const int lane = threadIdx.x & 31;
  const int n_round = (n_fg + 31) & ~31;
  // (a dynamic work queue -- warps drawing 32-entry chunks from a counter, grid = resident CTAs -- measured slower: 0.128 vs 0.118 ms, r13)
  for (int i = blockIdx.x * VH_C2_PB + threadIdx.x; i < n_round; i += gridDim.x * VH_C2_PB) {
    const bool on = i < n_fg;
    int id = 0, b = 0, key = -1 - lane;                 // inactive lanes: unique negative keys (runs of length 1, skipped)
    VertGrad vg;
#pragma unroll
    for (int k = 0; k < 3; ++k) { vg.gn[k] = mk3(0, 0, 0); vg.gc[k] = mk3(0, 0, 0); }
    if (on) {
      int pix = list[i]; id = tris[i];
      int x, y; vh_unflatten(A, pix, b, y, x);
      f4 g = grgb[i];
      passC2_body(P, b, y, x, mk3(g.x, g.y, g.z), gl, id, &vg);
      key = b * (A.F + 1) + id;
    }
    const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane == 0 || key != kprev;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const int run = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));       // lane of this run's head
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int run_d = __shfl_down_sync(0xffffffffu, run, d);
      const bool take = (lane + d < 32) && run_d == run;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float t;
        t = __shfl_down_sync(0xffffffffu, vg.gn[k].x, d); if (take) vg.gn[k].x += t;
        t = __shfl_down_sync(0xffffffffu, vg.gn[k].y, d); if (take) vg.gn[k].y += t;
        t = __shfl_down_sync(0xffffffffu, vg.gn[k].z, d); if (take) vg.gn[k].z += t;
        t = __shfl_down_sync(0xffffffffu, vg.gc[k].x, d); if (take) vg.gc[k].x += t;
        t = __shfl_down_sync(0xffffffffu, vg.gc[k].y, d); if (take) vg.gc[k].y += t;
        t = __shfl_down_sync(0xffffffffu, vg.gc[k].z, d); if (take) vg.gc[k].z += t;
      }
    }
    if (on && head) {
      const i4 f = A.faces[id - 1];
      const int vi[3] = {f.x, f.y, f.z};
      float* gv = P.g_vnorm + (size_t)(A.geo ? A.geo[b] : b) * A.V * 4;
      float* gc = P.g_clip + (size_t)b * A.V * 4;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (P.g_vnorm) VH_ATOMIC_ADD4(gv + (size_t)vi[k] * 4, vg.gn[k].x, vg.gn[k].y, vg.gn[k].z, 0.f);
        if (P.g_clip) VH_ATOMIC_ADD4(gc + (size_t)vi[k] * 4, vg.gc[k].x, vg.gc[k].y, 0.f, vg.gc[k].z);
      }
    }
  }
#else
  for (int i = blockIdx.x * VH_C2_PB + threadIdx.x; i < n_fg; i += gridDim.x * VH_C2_PB) {
    int pix = list[i], id = tris[i];
    int x, y, b; vh_unflatten(A, pix, b, y, x);
    f4 g = grgb[i];
    passC2_body(P, b, y, x, mk3(g.x, g.y, g.z), gl, id);
  }
#endif
  block_reduce_store<27, VH_C2_PB>(gl, sh, partials + (size_t)blockIdx.x * VH_NPART + 4);
}

// column sums of the [rows][VH_NPART] partial matrix: dst[slot ? slot[c] : c] += sum_r partials[r][col0 + c]
__global__ void __launch_bounds__(256) k_reduce_cols(const float* __restrict__ partials, int rows, int col0, int ncol, float* __restrict__ dst,
                                                     const int* __restrict__ slot) { VH_PDL_SYNC();
  __shared__ float sh[8][32];
  int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  float s = 0.f;
  if (c < ncol)
    for (int r = blockIdx.x * 8 + g; r < rows; r += gridDim.x * 8) s += partials[(size_t)r * VH_NPART + col0 + c];
  sh[g][c] = s;
  __syncthreads();
  if (g == 0 && c < ncol) {
    for (int k = 1; k < 8; ++k) s += sh[k][c];
    atomicAdd(dst + (slot ? slot[c] : c), s);
  }
}

// reduce_slab layout (floats): [0] sum |err|  [1] n(alpha_aa>0)  [2] sum var_c(diffuse) incl. background  [3] max diffuse
//                              [4] this rank's arg-max index (int bits; -1 = background)  [5] local max (to find the owner)  [6] n background px
// One CTA: column sums of the per-CTA partial rows of passes A / B (the former k_reduce_cols launch) into acc[], then the slab.
// Data parallel (dp.world > 1): the four batch-global scalars of the slab are exchanged through per-rank MAILBOXES that every peer maps over
// NVLink (CUDA IPC): this kernel stores its slab into slot [epoch & 1][rank] of every rank's mailbox and then raises flag[rank] = epoch
// there; k_finalize spins on its own mailbox's flags and reduces the slots in rank order (identical result on every rank).  Two slot
// parities make the reuse safe: a rank can only write epoch e + 2 after it passed k_finalize(e + 1), which needed every peer's flag e + 1,
// raised after that peer's k_finalize(e) had finished reading the slots of epoch e.
struct DpBox { int rank, world; float* const* peers; int* epoch; int* err; float* mine; unsigned long long* wait; };
__device__ __forceinline__ unsigned long long vh_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(256) k_forward_slab(const float* __restrict__ partials, int rows, float* __restrict__ acc,
                               const unsigned long long* __restrict__ maxslot, const float* __restrict__ lights, float n_pix_total, float* __restrict__ slab,
                               DpBox dp) { VH_PDL_SYNC();
  __shared__ float shs[64][4];
  __shared__ float tot[4];
  {
    const int col = threadIdx.x & 3, r0 = threadIdx.x >> 2;
    float s = 0.f;
    for (int r = r0; r < rows; r += 64) s += partials[(size_t)r * VH_NPART + col];
    shs[r0][col] = s;
    __syncthreads();
    if (threadIdx.x < 4) {
      float t = 0.f;
      for (int k = 0; k < 64; ++k) t += shs[k][threadIdx.x];
      const int slot = threadIdx.x == 0 ? ACC_VARSUM : (threadIdx.x == 1 ? ACC_NFGPIX : (threadIdx.x == 2 ? ACC_ABSERR : ACC_NFG));
      t += acc[slot];
      acc[slot] = t; tot[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
  }
  float nfgpix = tot[1], n_bg = n_pix_total - nfgpix;
  float dbg[3], mbg = 0.f;
  for (int c = 0; c < 3; ++c) { dbg[c] = VH_SH_C0 * lights[c] - VH_SH_C4 * lights[24 + c]; mbg += dbg[c] * (1.f / 3.f); }
  float vbg = 0.f, mxbg = fmaxf(dbg[0], fmaxf(dbg[1], dbg[2]));
  for (int c = 0; c < 3; ++c) vbg += 0.5f * (dbg[c] - mbg) * (dbg[c] - mbg);
  unsigned long long pm = *maxslot;
  float mxfg = pm ? unpack_max_val(pm) : -INFINITY;
  int idx = pm ? (int)(unsigned)(pm & 0xffffffffu) : -1;
  bool bg_is_max = n_bg > 0.f && mxbg > mxfg;
  slab[0] = tot[2]; slab[1] = tot[3]; slab[2] = tot[0] + n_bg * vbg;
  slab[3] = bg_is_max ? mxbg : mxfg;
  ((int*)slab)[4] = bg_is_max ? -1 : idx;
  slab[5] = slab[3];
  slab[6] = n_bg; slab[7] = 0.f;
  if (dp.world > 1) {
    const int e = *dp.epoch;
    for (int p = 0; p < dp.world; ++p) {
      float* box = dp.peers[p];
      float4* dst = (float4*)(box + ((size_t)(e & 1) * VH_DP_MAX + dp.rank) * 8);
      dst[0] = make_float4(slab[0], slab[1], slab[2], slab[3]);
    }
    __threadfence_system();
    for (int p = 0; p < dp.world; ++p) ((volatile int*)(dp.peers[p] + 2 * VH_DP_MAX * 8))[dp.rank] = e;
  }
}

// consumes the (possibly cross-rank reduced) slab: scal[] for pass C, loss values, background share of the light gradient
__global__ void k_finalize(const float* __restrict__ slab, const float* __restrict__ local_slab, vhap_stage_cfg cfg, const float* __restrict__ lights,
                           float n_pix_global, float* __restrict__ scal, float* __restrict__ acc, float* __restrict__ g_lights, DpBox dp) { VH_PDL_SYNC();
  float abs_sum = slab[0], nfg = slab[1], varsum = slab[2], mx = slab[3];
  if (dp.world > 1) {                                        // gather the peers' slabs from this rank's mailbox (see k_forward_slab)
    const int e = *dp.epoch;
    volatile int* flags = (volatile int*)(dp.mine + 2 * VH_DP_MAX * 8);
    const long long t0 = clock64();
    const unsigned long long w0 = vh_globaltimer();
    abs_sum = 0.f; nfg = 0.f; varsum = 0.f; mx = -INFINITY;
    for (int j = 0; j < dp.world; ++j) {
      while (flags[j] - e < 0) { if (clock64() - t0 > (1ll << 33)) { *dp.err = 1; break; } }      // ~4 s: a dead peer must not hang the GPU
      __threadfence_system();
      const volatile float* sl = dp.mine + ((size_t)(e & 1) * VH_DP_MAX + j) * 8;
      abs_sum += sl[0]; nfg += sl[1]; varsum += sl[2]; mx = fmaxf(mx, sl[3]);
    }
    *dp.epoch = e + 1;
    if (dp.wait) { dp.wait[0] += vh_globaltimer() - w0; dp.wait[3] += 1; }
  }
  float photo_scale = (cfg.w_photo >= 0.f && nfg > 0.f) ? cfg.w_photo / (3.f * nfg) : 0.f;
  scal[0] = photo_scale;
  // loss VALUES of batch-global terms: every rank holds the global value; its share is 1 / world so that the loss vectors sum to the total
  acc[ACC_PHOTO] = (cfg.w_photo >= 0.f && nfg > 0.f) ? cfg.shared_scale * cfg.w_photo * abs_sum / (3.f * nfg) : 0.f;
  bool regd = cfg.training && cfg.opt_lights && cfg.w_reg_diffuse >= 0.f;
  float g_var = regd ? cfg.w_reg_diffuse / n_pix_global : 0.f;
  float g_max = (regd && mx > 1.f) ? cfg.w_reg_diffuse : 0.f;
  bool owner = local_slab[5] == mx;                          // the rank holding the global max keeps its gradient
  int am = ((const int*)local_slab)[4];
  scal[1] = g_var;
  scal[2] = (owner && am >= 0) ? g_max : 0.f;
  ((int*)scal)[3] = am;
  acc[ACC_REG_DIFFUSE] = regd ? cfg.shared_scale * cfg.w_reg_diffuse * (fmaxf(mx - 1.f, 0.f) + varsum / n_pix_global) : 0.f;
  if (regd && g_lights) {                                    // background pixels: normal = 0 -> basis {C0, 0,..., -C4}
    float n_bg = local_slab[6];
    float dbg[3], mbg = 0.f;
    for (int c = 0; c < 3; ++c) { dbg[c] = VH_SH_C0 * lights[c] - VH_SH_C4 * lights[24 + c]; mbg += dbg[c] * (1.f / 3.f); }
    int chm = dbg[0] >= dbg[1] ? (dbg[0] >= dbg[2] ? 0 : 2) : (dbg[1] >= dbg[2] ? 1 : 2);
    for (int c = 0; c < 3; ++c) {
      float gd = n_bg * (dbg[c] - mbg) * g_var + ((owner && am < 0 && c == chm) ? g_max : 0.f);
      atomicAdd(g_lights + c, VH_SH_C0 * gd);
      atomicAdd(g_lights + 24 + c, -VH_SH_C4 * gd);
    }
  }
}

// flips a raster-orientation float4 plane into image orientation
__global__ void k_flip_plane(const float4* __restrict__ in, float4* __restrict__ out, int B, int H, int W) { VH_PDL_SYNC();
  size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)B * H * W;
  if (pix >= n) return;
  int x = pix % W, y = (pix / W) % H, b = pix / ((size_t)W * H);
  out[((size_t)b * H + (H - 1 - y)) * W + x] = in[pix];
}
void launch_flip_plane(vhap_ctx* c, const float* in, float* out, int B, int H, int W, cudaStream_t s) {
  size_t n = (size_t)B * H * W;
  LAUNCH(c, KID_MISC, s, vh_launch(k_flip_plane, (unsigned)((n + 255) / 256), 256, 0, s, (const float4*)in, (float4*)out, B, H, W));
}
__global__ void k_cid_plane(const int* __restrict__ tri_id, const uint8_t* __restrict__ fid2cid, float4* __restrict__ out, int B, int H, int W) { VH_PDL_SYNC();
  size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)B * H * W;
  if (pix >= n) return;
  int x = pix % W, y = (pix / W) % H, b = pix / ((size_t)W * H);
  int id = tri_id[pix];
  out[((size_t)b * H + (H - 1 - y)) * W + x] = make_float4((float)fid2cid[id], (float)id, 0.f, 0.f);
}
void launch_cid_plane(vhap_ctx* c, float* out, cudaStream_t s) {
  size_t n = (size_t)c->curB * c->curH * c->curW;
  LAUNCH(c, KID_MISC, s, vh_launch(k_cid_plane, (unsigned)((n + 255) / 256), 256, 0, s, c->tri_id, c->fid2cid, (float4*)out, c->curB, c->curH, c->curW));
}

void fill_render_args(vhap_ctx* c, PassArgs& P, const vhap_frame_batch* fb, const vhap_stage_cfg* cfg, const float* lights) {
  memset(&P, 0, sizeof(P));
  RenderArgs& A = P.R;
  A.B = fb->B; A.H = fb->H; A.W = fb->W; A.V = c->V; A.F = c->F; A.T = c->T; A.max_level = c->max_level;
  if ((A.W & (A.W - 1)) == 0 && (A.H & (A.H - 1)) == 0) {
    A.pow2 = 1;
    while ((1 << A.wshift) < A.W) ++A.wshift;
    while ((1 << A.hshift) < A.H) ++A.hshift;
  }
  A.faces = c->faces; A.faces_uv = c->faces_uv; A.verts_uv = c->verts_uv; A.clip = c->clip; A.vnorm = c->vnorm; A.lights = lights;
  A.mips = c->mips[c->cur_mip];
  for (int i = 0; i < VH_MAX_MIPS; ++i) A.mip_off[i] = c->mip_off[i];
  A.geo = fb->geo;
  A.tri_id = c->tri_id; A.face_flags = cfg->training ? c->face_flags : nullptr; A.vert_flags = cfg->training ? c->vert_flags : nullptr;
  A.fid2cid = c->fid2cid; A.adj_opp = c->adj_opp; A.ndc = c->ndc; A.zwbuf = (const float*)c->pre; A.tex_l0_flag = c->tex_l0_flag;
  P.target = (const uint16_t*)fb->target; P.target_u8 = fb->target_format == 1; P.pre = c->pre; P.signs = c->signs;
  P.final_rgba = c->want_planes ? c->final_rgba : nullptr;
  P.plane_albedo = c->want_planes ? c->plane_albedo : nullptr; P.plane_normal = c->want_planes ? c->plane_normal : nullptr;
  P.plane_diffuse = c->want_planes ? c->plane_diffuse : nullptr;
  P.pool_list = c->pool_list; P.pool_base = c->pool_base; P.pool_count = c->pool_count; P.pool_tri = c->pool_tri;
  P.disturb = cfg->training ? 1 : 0;                        // enable_disturbance = stage != None (tracker.py:427)
  P.rate_fg = cfg->disturb_rate_fg; P.rate_bg = cfg->disturb_rate_bg;
  P.loss_mask = c->loss_mask;
  P.inj_w = c->inj_w; P.inj_u = c->inj_u; P.seed = cfg->rng_seed; P.step = cfg->rng_step; P.step_ptr = c->use_dev_step ? c->dev_step : nullptr;
  P.bg_mode = cfg->bg_mode; P.bg_color[0] = cfg->bg_color[0]; P.bg_color[1] = cfg->bg_color[1]; P.bg_color[2] = cfg->bg_color[2];
  P.scal = c->scal; P.g_clip = c->g_clip; P.g_vnorm = c->g_vnorm; P.g_tex = nullptr; P.aa_code = c->aa_code;
}

static int* slot_table(vhap_ctx* c) {      // device copy of the acc[] slots of partial columns 0..3
  static int* d_slot[16] = {nullptr};
  int dev = c->device & 15;
  if (!d_slot[dev]) {
    static const int h_slot[4] = {ACC_VARSUM, ACC_NFGPIX, ACC_ABSERR, ACC_NFG};
    cudaMalloc(&d_slot[dev], sizeof(h_slot));
    cudaMemcpy(d_slot[dev], h_slot, sizeof(h_slot), cudaMemcpyHostToDevice);
  }
  return d_slot[dev];
}

void launch_render_forward(vhap_ctx* c, PassArgs& P, cudaStream_t s, bool zeroed) {
  const RenderArgs& A = P.R;
  size_t n = (size_t)A.B * A.H * A.W;
  int nblk = (int)((n + PB - 1) / PB);
  int* slots = slot_table(c);
  if (!zeroed) { VhZeroSegs z; z.n = 2; z.p[0] = c->maxslot; z.bytes[0] = sizeof(unsigned long long); z.p[1] = c->pair_count; z.bytes[1] = sizeof(int); vh_zero_multi(c, z, s); }
  if (c->want_planes) {
    cudaMemsetAsync(c->plane_albedo, 0, n * 16, s); cudaMemsetAsync(c->plane_normal, 0, n * 16, s); cudaMemsetAsync(c->plane_diffuse, 0, n * 16, s);
  }
  const int nblk_pool = (int)((n + (size_t)PB * POOL_PPT - 1) / ((size_t)PB * POOL_PPT));
  LAUNCH(c, KID_POOL_COUNT, s, vh_launch(k_pool_count, nblk_pool, PB, 0, s, A.tri_id, A.fid2cid, n, c->n_clusters, c->pool_blk_count));
  launch_scan(c, c->pool_blk_count, c->pool_blk_off, 16 * nblk_pool, c->scan_total, s);
  LAUNCH(c, KID_POOL_SCATTER, s, vh_launch(k_pool_scatter, nblk_pool, PB, 0, s, A.tri_id, A.fid2cid, n, A.H, A.W, c->n_clusters, c->pool_blk_off, c->pool_list, c->pool_tri, c->pair_list, c->pair_count,
                                                                   c->scan_total, c->pool_base, c->pool_count));
  int grid = nblk < NPERSIST ? nblk : NPERSIST;
  // a deferred texture update (vhap_set_render_wait_event) is joined here: everything above is independent of the texture
  if (c->render_wait_ev) { cudaStreamWaitEvent(s, c->render_wait_ev, 0); c->render_wait_ev = nullptr; }
  LAUNCH(c, KID_PASSA, s, vh_launch(k_passA, grid, PB, 0, s, P, c->partials, c->maxslot));
  LAUNCH(c, KID_AA_PAIRS, s, vh_launch(k_aa_pairs, grid, PB, 0, s, P, c->pair_list, c->pair_count));
  LAUNCH(c, KID_PASSB, s, vh_launch(k_passB, grid, PB, 0, s, P, c->partials));
  (void)slots;                              // the partial rows are summed by k_forward_slab (launch_forward_slab)
}

static DpBox dp_box_of(vhap_ctx* c) {
  DpBox d; d.rank = c->dp_rank; d.world = c->dp_peers_dev ? c->dp_world : 1; d.peers = c->dp_peers_dev; d.epoch = c->dp_epoch; d.err = c->dp_err; d.mine = c->dp_box; d.wait = c->dp_wait;
  return d;
}
void launch_forward_slab(vhap_ctx* c, const PassArgs& P, const float* lights, float* slab, cudaStream_t s) {
  const RenderArgs& A = P.R;
  size_t n = (size_t)A.B * A.H * A.W;
  int nblk = (int)((n + PB - 1) / PB), rows = nblk < NPERSIST ? nblk : NPERSIST;          // grid of passes A / B (launch_render_forward)
  LAUNCH(c, KID_SLAB, s, vh_launch(k_forward_slab, 1, 256, 0, s, c->partials, rows, c->acc, c->maxslot, lights, (float)n, slab, dp_box_of(c)));
}

void launch_finalize(vhap_ctx* c, const PassArgs& P, const vhap_stage_cfg* cfg, const float* slab_global, const float* slab_local, int global_B,
                     const float* lights, float* g_lights, cudaStream_t s) {
  const RenderArgs& A = P.R;
  LAUNCH(c, KID_FINALIZE, s, vh_launch(k_finalize, 1, 1, 0, s, slab_global, slab_local, *cfg, lights, (float)((size_t)global_B * A.H * A.W), c->scal, c->acc, g_lights, dp_box_of(c)));
}

void launch_render_backward(vhap_ctx* c, PassArgs& P, const vhap_stage_cfg* cfg, const float* lights, float* g_lights, const float* ext_grad, cudaStream_t s,
                            cudaStream_t side) {
  (void)cfg; (void)lights;
  const RenderArgs& A = P.R;
  size_t n = (size_t)A.B * A.H * A.W;
  int nblk = (int)((n + PB - 1) / PB);
  int grid = nblk < NPERSIST ? nblk : NPERSIST;
  LAUNCH(c, KID_PASSC1, s, vh_launch(k_passC1, grid, PB, 0, s, P, ext_grad, c->grgb));
  int grid2 = grid * (PB / VH_C2_PB);
#if VH_C2_SPLIT
  // the texel-gradient scatter (8 of the 10 vector reductions per pixel, no texel reads) runs as its own light kernel on a second
  // high-priority stream BESIDE the geometry half: two latency-bound kernels share the machine instead of one at 25 % occupancy
  if (side && P.g_tex) {
    cudaEventRecord(c->ev[EV_C1_DONE], s);
    cudaStreamWaitEvent(c->hp[1], c->ev[EV_C1_DONE], 0);
    LAUNCH(c, KID_PASSC1, c->hp[1], vh_launch(k_passC2_tex, grid, PB, 0, c->hp[1], P, c->grgb));
    cudaEventRecord(c->ev[EV_C2TEX_DONE], c->hp[1]);
    PassArgs Pg = P; Pg.g_tex = nullptr;
    LAUNCH(c, KID_PASSC, s, vh_launch(k_passC2, grid2, VH_C2_PB, 0, s, Pg, c->grgb, c->partials));
    cudaStreamWaitEvent(s, c->ev[EV_C2TEX_DONE], 0);
  } else
#endif
  LAUNCH(c, KID_PASSC, s, vh_launch(k_passC2, grid2, VH_C2_PB, 0, s, P, c->grgb, c->partials));
  // side != NULL: the reduction of the light-gradient partials (only the Adam step needs it) leaves the step's critical chain; the caller
  // joins EV_LIGHTS_DONE.  EV_TEXGRAD_READY doubles as "pass C complete".
  cudaStream_t ls = s;
  if (side) { cudaEventRecord(c->ev[EV_TEXGRAD_READY], s); cudaStreamWaitEvent(side, c->ev[EV_TEXGRAD_READY], 0); ls = side; }
  if (g_lights) LAUNCH(c, KID_LIGHTS_REDUCE, ls, vh_launch(k_reduce_cols, 16, 256, 0, ls, c->partials, grid2, 4, 27, g_lights, nullptr));
  if (side) cudaEventRecord(c->ev[EV_LIGHTS_DONE], side);
}
