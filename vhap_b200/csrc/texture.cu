// Texture path: albedo = tex_painted + tex_extra (get_albedo, vhap/model/tracker.py:247-258) -> float4 mip pyramid
// (the pyramid nvdiffrast builds inside dr.texture, vhap/util/render_nvdiffrast.py:399 -- here ONE copy shared by all
// frames instead of B expanded copies, tracker.py:234); fold of the texel-gradient pyramid to level 0, total-variation
// and residual-cluster regularisers (tracker.py:526-539), Adam (torch.optim.Adam, tracker.py:210) and rebuild of level 0,
// fused into one streaming pass over the 3*T*T texels.
#include <stdlib.h>
#include "engine.h"
#include "accum.h"

__global__ void k_tex_level0(const float* __restrict__ painted, const float* __restrict__ extra, int T, f4* __restrict__ out) { VH_PDL_SYNC();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)T * T;
  if (i >= n) return;
  f4 v;
  v.x = (painted ? painted[i] : 0.f) + extra[i];
  v.y = (painted ? painted[n + i] : 0.f) + extra[n + i];
  v.z = (painted ? painted[2 * n + i] : 0.f) + extra[2 * n + i];
  v.w = 0.f;
  out[i] = v;
}

// 2x2 box-filter pyramid, up to 5 levels per launch: each CTA loads a 32x32 tile of the source level into shared memory and
// emits the 16x16, 8x8, 4x4, 2x2 and 1x1 reductions of that tile to the next levels (dst level k has size ssz >> (k+1)).
__device__ __forceinline__ f4 avg4(const f4& a, const f4& b, const f4& c, const f4& d) {
  f4 o = {0.25f * (a.x + b.x + c.x + d.x), 0.25f * (a.y + b.y + c.y + d.y), 0.25f * (a.z + b.z + c.z + d.z), 0.f};
  return o;
}
__global__ void __launch_bounds__(256) k_mip_down(const f4* __restrict__ src, f4* __restrict__ pyr_base, int ssz, int nlev,
                                                  int o1, int o2, int o3, int o4, int o5) { VH_PDL_SYNC();
  __shared__ f4 A[32][33];
  __shared__ f4 Bf[16][17];
  int tile = ssz < 32 ? ssz : 32;                  // source tile edge handled by this CTA
  int tx0 = blockIdx.x * tile, ty0 = blockIdx.y * tile;
  for (int i = threadIdx.x; i < tile * tile; i += blockDim.x) {
    int x = i % tile, y = i / tile;
    A[y][x] = src[(size_t)(ty0 + y) * ssz + tx0 + x];
  }
  __syncthreads();
  const int offs[5] = {o1, o2, o3, o4, o5};
  int cur = tile;                                   // edge of the data currently in A (even levels) / Bf (odd levels)
  for (int k = 0; k < nlev; ++k) {
    int nxt = cur >> 1, dsz = ssz >> (k + 1);
    f4* dst = pyr_base + offs[k];
    int dx0 = tx0 >> (k + 1), dy0 = ty0 >> (k + 1);
    for (int i = threadIdx.x; i < nxt * nxt; i += blockDim.x) {
      int x = i % nxt, y = i / nxt;
      f4 o = (k & 1) ? avg4(Bf[2 * y][2 * x], Bf[2 * y + 1][2 * x], Bf[2 * y][2 * x + 1], Bf[2 * y + 1][2 * x + 1])
                     : avg4(A[2 * y][2 * x], A[2 * y + 1][2 * x], A[2 * y][2 * x + 1], A[2 * y + 1][2 * x + 1]);
      dst[(size_t)(dy0 + y) * dsz + dx0 + x] = o;
      if (k & 1) A[y][x] = o; else Bf[y][x] = o;
    }
    __syncthreads();
    cur = nxt;
  }
}

static void build_mips(vhap_ctx* c, f4* pyr, cudaStream_t s, int from_level = 0) {
  int l = from_level;
  while (l < c->max_level) {
    int ssz = c->T >> l, nlev = c->max_level - l < 5 ? c->max_level - l : 5;
    int tile = ssz < 32 ? ssz : 32, g = ssz / tile;
    int o[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < nlev; ++k) o[k] = c->mip_off[l + 1 + k];
    LAUNCH(c, KID_MIP, s, vh_launch(k_mip_down, dim3(g, g), 256, 0, s, pyr + c->mip_off[l], pyr, ssz, nlev, o[0], o[1], o[2], o[3], o[4]));
    l += nlev;
  }
}

void launch_tex_rebuild(vhap_ctx* c, const float* tex_extra, cudaStream_t s) {
  size_t n = (size_t)c->T * c->T;
  f4* pyr = c->mips[c->cur_mip];
  LAUNCH(c, KID_TEX_L0, s, vh_launch(k_tex_level0, (unsigned)((n + 255) / 256), 256, 0, s, c->tex_painted, tex_extra, c->T, pyr));
  build_mips(c, pyr, s);
}

struct TexFoldArgs {
  int T, max_level; int mip_off[VH_MAX_MIPS];
  const f4* tex_old; f4* tex_new; float* g_pyr;
  float* extra; float* g_out; float* m; float* v;
  const float* g_in;            // optional dense [3,T,T] gradient that REPLACES the folded one (data parallel: the all-reduced gradient)
  const uint8_t* mask;
  float w_tv, w_res;            // already divided by their mean() denominators and scaled by shared_scale
  float lr, bc1, bc2_sqrt;      // Adam: step size lr/bc1, sqrt(bias correction 2)
  int do_adam;
  int* l0_flag;                 // [regions] level-0 flags of the gradient pyramid, region = (y >> 3) * max(T / 256, 1) + (x >> 8)
  const int* step_ptr;          // device Adam step (CUDA-graph replay) or NULL
  const float* lr_scale_ptr;    // device learning-rate scale, read together with step_ptr
  int step_bias;                // added to *step_ptr (deferred update of the previous step: -1)
  // ---- sharded (data-parallel) texture update, k_tex_fold2 only
  int y_begin, y_end;           // rows handled by this launch (a row band; whole texture: 0, T)
  int rm;                       // row-major exchange layout: g_out[(y*3 + c)*T + x] over the whole texture (the reduce-scatter input, a rank's band is
                                // contiguous), g_in[((y - y_begin)*3 + c)*T + x] over the band (the reduce-scatter output)
  float* ex_band_out;           // optional: updated tex_extra of the band, row-major band-local (the all-gather input)
  int no_pyramid;               // do not write level 0 / 1 of the new pyramid (the band owner only updates tex_extra; vhap_tex_rebuild_rm rebuilds)
};

__device__ __forceinline__ float chan(const f4& t, int c) { return c == 0 ? t.x : (c == 1 ? t.y : t.z); }

// ---- The fold kernels.  In ONE pass over the texture: fold of the texel-gradient pyramid back to level 0 (box-filter adjoint, 1/4 per
// level), total-variation + residual regularisers (tracker.py:526-539), Adam (tracker.py:210), level 0 AND level 1 of the OTHER pyramid
// (the mip rebuild of render_nvdiffrast.py:399 then starts from level 1), loss partials reduced by the last CTA (deterministic order).
// A CTA walks a band of TF_ROWS rows of a 256-texel-wide strip top to bottom (wide strips: 1 KB-contiguous pieces of every planar array,
// DRAM-page friendly; square tiles measured 1.4-3x slower in round 1):
//   * rows above / below a texel (total variation) come from the previous / next iteration: the halo is 2 rows per TF_ROWS (the round-1
//     kernel re-read both neighbour rows of every 2-row strip: +16 B/texel),
//   * the coarse gradient levels are folded hierarchically per band (levels >= 3 once per 8x8 block, level 2 per 4x4, level 1 per 2x2),
//   * level 0 of the gradient pyramid (32 B/texel of read + re-zero) is only touched where the backward scattered into it: one flag per
//     (8-row, 256-column) region, raised by tex_sample_bwd, cleared here.
// k_tex_fold2: register pipeline (loads of row y+1 in flight while row y is processed), any T.  k_tex_fold3 (T >= 256): the rows are
// staged in shared memory by TMA 1-D bulk copies (cp.async.bulk + mbarrier expect-tx), TF3_NS rows deep, so that ~2 rows x 17 KB per CTA
// are in flight regardless of the register budget; the arithmetic is identical.
#ifndef TF_ROWS
#define TF_ROWS 8
#endif
#ifndef TF_MINB
#define TF_MINB 4
#endif
struct RowIn { f4 t; float ex[3], m[3], v[3], g[3]; unsigned char msk; };

__device__ __forceinline__ void row_load(const TexFoldArgs& a, int x, int y, bool l0, RowIn& r) {
  const size_t n = (size_t)a.T * a.T, i = (size_t)y * a.T + x;
  const bool need_t = a.w_tv > 0.f || a.do_adam, need_ex = a.do_adam || (a.w_res > 0.f && a.mask);
  if (need_t) r.t = a.tex_old[i];
  r.g[0] = r.g[1] = r.g[2] = 0.f;
  if (a.g_in) {
    if (a.rm) { const float* gi = a.g_in + ((size_t)(y - a.y_begin) * 3) * a.T + x; r.g[0] = gi[0]; r.g[1] = gi[a.T]; r.g[2] = gi[2 * (size_t)a.T]; }
    else { r.g[0] = a.g_in[i]; r.g[1] = a.g_in[n + i]; r.g[2] = a.g_in[2 * n + i]; }
  }
  else if (l0) { float4 g0 = *(const float4*)(a.g_pyr + i * 4); r.g[0] = g0.x; r.g[1] = g0.y; r.g[2] = g0.z; }
#pragma unroll
  for (int c = 0; c < 3; ++c) { if (need_ex) r.ex[c] = a.extra[c * n + i]; if (a.do_adam) { r.m[c] = a.m[c * n + i]; r.v[c] = a.v[c * n + i]; } }
  r.msk = (a.w_res > 0.f && a.mask) ? a.mask[i] : 0;
}

__global__ void __launch_bounds__(256, TF_MINB) k_tex_fold2(TexFoldArgs a, float* __restrict__ partials, unsigned* __restrict__ counter, float* __restrict__ acc_out) { VH_PDL_SYNC();
  __shared__ float sh[8 * 2];
  __shared__ float c1[(TF_ROWS / 2) * 128][3];        // folded gradient of levels >= 1 per level-1 texel of the band
  __shared__ float c2[((TF_ROWS + 3) / 4) * 64][3];
  __shared__ float c3[((TF_ROWS + 7) / 8) * 32][3];
  __shared__ bool is_last;
  const int T = a.T, tw = T < 256 ? T : 256, tpr = T / tw, R = T < TF_ROWS ? T : TF_ROWS;
  const int tid = threadIdx.x, lane = tid & 31;
  const int x0 = (blockIdx.x % tpr) * tw, x = x0 + tid, y0 = a.y_begin + (blockIdx.x / tpr) * R;     // (y_begin is a multiple of R: host check)
  const int n1x = tw >> 1, n1y = R >> 1, n2x = tw >> 2 ? tw >> 2 : 1, n2y = R >> 2 ? R >> 2 : 1, n3x = tw >> 3 ? tw >> 3 : 1, n3y = R >> 3 ? R >> 3 : 1;
  const bool fold = a.g_pyr != nullptr && a.g_in == nullptr;
  // ---- hierarchical fold of the coarse gradient levels (box-filter adjoint: 1/4 per level)
  for (int j = tid; j < n3x * n3y; j += 256) {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (fold) {
      const int X = (x0 >> 3) + j % n3x, Y = (y0 >> 3) + j / n3x;
      float sc = 1.f / 64.f;
      for (int l = 3; l <= a.max_level; ++l) {
        float4 p = *(const float4*)(a.g_pyr + ((size_t)a.mip_off[l] + (size_t)(Y >> (l - 3)) * (T >> l) + (X >> (l - 3))) * 4);
        g0 += p.x * sc; g1 += p.y * sc; g2 += p.z * sc;
        sc *= 0.25f;
      }
    }
    c3[j][0] = g0; c3[j][1] = g1; c3[j][2] = g2;
  }
  __syncthreads();
  for (int j = tid; j < n2x * n2y; j += 256) {
    const int jx = j % n2x, jy = j / n2x;
    const int p3 = ((jy >> 1) < n3y ? (jy >> 1) : n3y - 1) * n3x + ((jx >> 1) < n3x ? (jx >> 1) : n3x - 1);
    float g0 = c3[p3][0], g1 = c3[p3][1], g2 = c3[p3][2];
    if (fold && a.max_level >= 2) {
      const int X = (x0 >> 2) + jx, Y = (y0 >> 2) + jy;
      float4 p = *(const float4*)(a.g_pyr + ((size_t)a.mip_off[2] + (size_t)Y * (T >> 2) + X) * 4);
      g0 += p.x * (1.f / 16.f); g1 += p.y * (1.f / 16.f); g2 += p.z * (1.f / 16.f);
    }
    c2[j][0] = g0; c2[j][1] = g1; c2[j][2] = g2;
  }
  __syncthreads();
  for (int j = tid; j < n1x * n1y; j += 256) {
    const int jx = j % n1x, jy = j / n1x;
    const int p2 = ((jy >> 1) < n2y ? (jy >> 1) : n2y - 1) * n2x + ((jx >> 1) < n2x ? (jx >> 1) : n2x - 1);
    float g0 = c2[p2][0], g1 = c2[p2][1], g2 = c2[p2][2];
    if (fold && a.max_level >= 1) {
      const int X = (x0 >> 1) + jx, Y = (y0 >> 1) + jy;
      float4 p = *(const float4*)(a.g_pyr + ((size_t)a.mip_off[1] + (size_t)Y * (T >> 1) + X) * 4);
      g0 += p.x * 0.25f; g1 += p.y * 0.25f; g2 += p.z * 0.25f;
    }
    c1[j][0] = g0; c1[j][1] = g1; c1[j][2] = g2;
  }
  __syncthreads();
  float acc[2] = {0.f, 0.f};
  const bool on = tid < tw;
  int* l0p = (fold && a.l0_flag) ? a.l0_flag + (y0 >> 3) * tpr + (x0 >> 8) : nullptr;      // this CTA's region flag
  const bool l0 = fold ? (l0p ? (*l0p != 0) : true) : false;
  float bc1 = a.bc1, bc2s = a.bc2_sqrt;
  if (a.do_adam && a.step_ptr) {
    float st = (float)(a.step_ptr[0] + a.step_bias); bc2s = sqrtf(1.f - powf(0.999f, st));
    bc1 = (1.f - powf(0.9f, st)) / a.lr_scale_ptr[0];
  }
  const size_t n = (size_t)T * T;
  const bool tv = a.w_tv > 0.f;
  __syncthreads();
  if (tid == 0 && l0p) *l0p = 0;
  RowIn cur = {}, nxt;
  f4 t_up = {0, 0, 0, 0}, o_prev = {0, 0, 0, 0};
  if (on) {
    row_load(a, x, y0, l0, cur);
    t_up = (tv && y0 > 0) ? a.tex_old[(size_t)(y0 - 1) * T + x] : cur.t;
  }
  nxt = cur;
#pragma unroll 1
  for (int r = 0; r < R; ++r) {
    const int y = y0 + r;
    const size_t i = (size_t)y * T + x;
    f4 t_dn = cur.t;
    if (on) {
      if (r + 1 < R) { row_load(a, x, y + 1, l0, nxt); t_dn = nxt.t; }           // next row's loads are in flight during this row's arithmetic
      else if (tv && y + 1 < T) t_dn = a.tex_old[i + T];
    }
    f4 o = {0, 0, 0, 0};
    // left / right neighbours of this row through the warp; the strip's outside neighbours by lanes 0 / 31
    f4 tl, tr;
    tl.x = __shfl_up_sync(0xffffffffu, cur.t.x, 1); tl.y = __shfl_up_sync(0xffffffffu, cur.t.y, 1); tl.z = __shfl_up_sync(0xffffffffu, cur.t.z, 1);
    tr.x = __shfl_down_sync(0xffffffffu, cur.t.x, 1); tr.y = __shfl_down_sync(0xffffffffu, cur.t.y, 1); tr.z = __shfl_down_sync(0xffffffffu, cur.t.z, 1);
    if (on) {
      const f4 t = cur.t;
      if (tv) {
        if (lane == 0) tl = x > 0 ? a.tex_old[i - 1] : t;
        if (lane == 31 || x + 1 >= T || tid + 1 >= tw) tr = x + 1 < T ? a.tex_old[i + 1] : t;
      }
      const float* cg = c1[(r >> 1) * n1x + (tid >> 1)];
      float g[3] = {cur.g[0], cur.g[1], cur.g[2]};
      if (fold) { g[0] += cg[0]; g[1] += cg[1]; g[2] += cg[2]; }
      if (l0) *(float4*)(a.g_pyr + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);       // coarser levels: memset after the kernel
      if (tv) {                                                                      // tracker.py:526-534
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = chan(t, c), dr = v - chan(tr, c), dd = v - chan(t_dn, c), dl = chan(tl, c) - v, du = chan(t_up, c) - v;
          acc[0] += a.w_tv * (dr * dr + dd * dd);
          g[c] += 2.f * a.w_tv * (dr + dd - dl - du);
        }
      }
      if (cur.msk) {                                                                 // tracker.py:536-539
#pragma unroll
        for (int c = 0; c < 3; ++c) { acc[1] += a.w_res * cur.ex[c] * cur.ex[c]; g[c] += 2.f * a.w_res * cur.ex[c]; }
      }
      if (a.g_out) {
        if (a.rm) { float* go = a.g_out + ((size_t)y * 3) * T + x; go[0] = g[0]; go[T] = g[1]; go[2 * (size_t)T] = g[2]; }
        else { a.g_out[i] = g[0]; a.g_out[n + i] = g[1]; a.g_out[2 * n + i] = g[2]; }
      }
      o = t;
      if (a.do_adam) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const size_t k = c * n + i;
          float m = 0.9f * cur.m[c] + 0.1f * g[c];
          float v = 0.999f * cur.v[c] + 0.001f * g[c] * g[c];
          a.m[k] = m; a.v[k] = v;
          float upd = (a.lr / bc1) * m / (sqrtf(v) / bc2s + 1e-8f);
          float ne = cur.ex[c] - upd;
          a.extra[k] = ne;
          if (a.ex_band_out) a.ex_band_out[((size_t)(y - a.y_begin) * 3 + c) * T + x] = ne;
          float base = chan(t, c) - cur.ex[c];                                        // painted part
          if (c == 0) o.x = base + ne; else if (c == 1) o.y = base + ne; else o.z = base + ne;
        }
        if (!a.no_pyramid) a.tex_new[i] = o;
      }
    }
    // level 1 of the new pyramid on odd rows: avg4(A[2y][2x], A[2y+1][2x], A[2y][2x+1], A[2y+1][2x+1]), same order as k_mip_down
    if (r & 1) {
      f4 c, d;
      c.x = __shfl_down_sync(0xffffffffu, o_prev.x, 1); c.y = __shfl_down_sync(0xffffffffu, o_prev.y, 1); c.z = __shfl_down_sync(0xffffffffu, o_prev.z, 1);
      d.x = __shfl_down_sync(0xffffffffu, o.x, 1); d.y = __shfl_down_sync(0xffffffffu, o.y, 1); d.z = __shfl_down_sync(0xffffffffu, o.z, 1);
      c.w = d.w = 0.f;
      if (on && a.do_adam && !a.no_pyramid && a.max_level >= 1 && !(tid & 1))
        a.tex_new[(size_t)a.mip_off[1] + (size_t)(y >> 1) * (T >> 1) + (x >> 1)] = avg4(o_prev, o, c, d);
    }
    o_prev = o;
    t_up = cur.t;
    cur = nxt;
  }
  // block partial sums of the two loss terms; the last CTA reduces all partials in a fixed order
  const int w = tid >> 5;
  for (int q = 0; q < 2; ++q) {
    float v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[w * 2 + q] = v;
  }
  __syncthreads();
  if (tid < 2) {
    float s2 = 0.f;
    for (int k = 0; k < 8; ++k) s2 += sh[k * 2 + tid];
    partials[(size_t)blockIdx.x * 2 + tid] = s2;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int q = 0; q < 2; ++q) {
    float s2 = 0.f;
    for (int r = tid; r < (int)gridDim.x; r += 256) s2 += __ldcg(partials + (size_t)r * 2 + q);
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    __syncthreads();
    if (lane == 0) sh[w] = s2;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += sh[k];
      acc_out[q] = t;
    }
  }
  if (tid == 0) *counter = 0u;
}


// ---- TMA-staged version (T >= 256, 256-wide strips)
#ifndef TF3_NS
#define TF3_NS 3
#endif
struct __align__(16) TF3Row {
  f4 t[258];                  // texture row incl. the left / right neighbour columns of the strip
  float ex[3][256], m[3][256], v[3][256];
  f4 g[256];                  // level 0 of the gradient pyramid (fold mode) or, as float[3][256], the dense input gradient (apply / band mode)
  unsigned char msk[256];
};
__device__ __forceinline__ void tf3_wait(uint64_t* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar), done = 0;
  while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tf3_copy(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// row q of the band (q == R: the halo row below the band, texture values only) into ring slot `slot`
__device__ void tf3_issue(const TexFoldArgs& a, TF3Row* rows, uint64_t* bars, int slot, int x0, int y, bool full, bool l0, bool want_mask) {
  const int T = a.T;
  const size_t n = (size_t)T * T, i = (size_t)y * T + x0;
  TF3Row& r = rows[slot];
  uint64_t* bar = bars + slot;
  const int left = x0 > 0 ? 1 : 0, right = x0 + 256 < T ? 1 : 0;
  // what this launch needs per row: the texture values for TV / the new pyramid, tex_extra for Adam / the residual term, ... (the pure
  // gradient fold of the sharded update, vhap_tex_fold_grad_rm, needs none of them: it only streams the flagged level-0 gradient rows)
  const bool need_t = a.w_tv > 0.f || a.do_adam, need_ex = a.do_adam || want_mask;
  unsigned bytes = need_t ? (unsigned)(256 + left + right) * 16u : 0u;
  if (full) {
    if (need_ex) bytes += 3u * 1024u;
    if (a.do_adam) bytes += 6u * 1024u;
    if (a.g_in) bytes += 3u * 1024u; else if (l0) bytes += 4096u;
    if (want_mask) bytes += 256u;
  }
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
  if (need_t) tf3_copy(&r.t[1 - left], a.tex_old + i - left, (unsigned)(256 + left + right) * 16u, bar);
  if (!full) return;
  for (int c = 0; c < 3; ++c) {
    if (need_ex) tf3_copy(r.ex[c], a.extra + c * n + i, 1024u, bar);
    if (a.do_adam) { tf3_copy(r.m[c], a.m + c * n + i, 1024u, bar); tf3_copy(r.v[c], a.v + c * n + i, 1024u, bar); }
    if (a.g_in) {
      const float* gi = a.rm ? a.g_in + ((size_t)(y - a.y_begin) * 3 + c) * T + x0 : a.g_in + c * n + i;
      tf3_copy((float*)r.g + c * 256, gi, 1024u, bar);
    }
  }
  if (!a.g_in && l0) tf3_copy(r.g, a.g_pyr + i * 4, 4096u, bar);
  if (want_mask) tf3_copy(r.msk, a.mask + i, 256u, bar);
}

__global__ void __launch_bounds__(256, 3) k_tex_fold3(TexFoldArgs a, float* __restrict__ partials, unsigned* __restrict__ counter, float* __restrict__ acc_out) { VH_PDL_SYNC();
  extern __shared__ __align__(128) unsigned char tf3_smem[];
  TF3Row* rows = (TF3Row*)tf3_smem;
  __shared__ uint64_t bars[TF3_NS];
  __shared__ float sh[8 * 2];
  __shared__ float c1[(TF_ROWS / 2) * 128][3];
  __shared__ float c2[((TF_ROWS + 3) / 4) * 64][3];
  __shared__ float c3[((TF_ROWS + 7) / 8) * 32][3];
  __shared__ bool is_last;
  const int T = a.T, tpr = T >> 8, R = TF_ROWS;
  const int tid = threadIdx.x, lane = tid & 31;
  const int x0 = (blockIdx.x % tpr) << 8, x = x0 + tid, y0 = a.y_begin + (blockIdx.x / tpr) * R;
  const bool fold = a.g_pyr != nullptr && a.g_in == nullptr;
  const bool tv = a.w_tv > 0.f;
  const bool want_mask = a.w_res > 0.f && a.mask != nullptr;
  int* l0p = (fold && a.l0_flag) ? a.l0_flag + (y0 >> 3) * tpr + (x0 >> 8) : nullptr;
  const bool l0 = fold ? (l0p ? (*l0p != 0) : true) : false;
  const bool halo_dn = tv && (y0 + R < T);
  const int q_last = halo_dn ? R : R - 1;                      // last row of the pipeline (the halo row carries the texture only)
  if (tid == 0) {
    for (int i = 0; i < TF3_NS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int q = 0; q < TF3_NS && q <= q_last; ++q) tf3_issue(a, rows, bars, q, x0, y0 + q, q < R, l0, want_mask);
    if (l0p) *l0p = 0;
  }
  // ---- hierarchical fold of the coarse gradient levels while the first rows are in flight
  for (int j = tid; j < 32; j += 256) {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (fold) {
      const int X = (x0 >> 3) + j, Y = y0 >> 3;
      float sc = 1.f / 64.f;
      for (int l = 3; l <= a.max_level; ++l) {
        float4 p = *(const float4*)(a.g_pyr + ((size_t)a.mip_off[l] + (size_t)(Y >> (l - 3)) * (T >> l) + (X >> (l - 3))) * 4);
        g0 += p.x * sc; g1 += p.y * sc; g2 += p.z * sc;
        sc *= 0.25f;
      }
    }
    c3[j][0] = g0; c3[j][1] = g1; c3[j][2] = g2;
  }
  __syncthreads();
  for (int j = tid; j < 128; j += 256) {
    const int jx = j & 63, jy = j >> 6;
    const int p3 = jx >> 1;
    float g0 = c3[p3][0], g1 = c3[p3][1], g2 = c3[p3][2];
    if (fold && a.max_level >= 2) {
      float4 p = *(const float4*)(a.g_pyr + ((size_t)a.mip_off[2] + (size_t)((y0 >> 2) + jy) * (T >> 2) + (x0 >> 2) + jx) * 4);
      g0 += p.x * (1.f / 16.f); g1 += p.y * (1.f / 16.f); g2 += p.z * (1.f / 16.f);
    }
    c2[j][0] = g0; c2[j][1] = g1; c2[j][2] = g2;
  }
  __syncthreads();
  for (int j = tid; j < 512; j += 256) {
    const int jx = j & 127, jy = j >> 7;
    const int p2 = (jy >> 1) * 64 + (jx >> 1);
    float g0 = c2[p2][0], g1 = c2[p2][1], g2 = c2[p2][2];
    if (fold && a.max_level >= 1) {
      float4 p = *(const float4*)(a.g_pyr + ((size_t)a.mip_off[1] + (size_t)((y0 >> 1) + jy) * (T >> 1) + (x0 >> 1) + jx) * 4);
      g0 += p.x * 0.25f; g1 += p.y * 0.25f; g2 += p.z * 0.25f;
    }
    c1[j][0] = g0; c1[j][1] = g1; c1[j][2] = g2;
  }
  __syncthreads();
  float acc[2] = {0.f, 0.f};
  float bc1 = a.bc1, bc2s = a.bc2_sqrt;
  if (a.do_adam && a.step_ptr) {
    float st = (float)(a.step_ptr[0] + a.step_bias); bc2s = sqrtf(1.f - powf(0.999f, st));
    bc1 = (1.f - powf(0.9f, st)) / a.lr_scale_ptr[0];
  }
  const size_t n = (size_t)T * T;
  f4 t_up = {0, 0, 0, 0}, o_prev = {0, 0, 0, 0};
  const bool halo_up = tv && y0 > 0;
  if (halo_up) t_up = a.tex_old[(size_t)(y0 - 1) * T + x];
#pragma unroll 1
  for (int r = 0; r < R; ++r) {
    const int y = y0 + r, slot = r % TF3_NS;
    const size_t i = (size_t)y * T + x;
    tf3_wait(&bars[slot], (unsigned)((r / TF3_NS) & 1));
    const TF3Row& cur = rows[slot];
    const f4 t = cur.t[tid + 1];
    f4 t_dn = t;
    if (r + 1 <= q_last) {
      tf3_wait(&bars[(r + 1) % TF3_NS], (unsigned)(((r + 1) / TF3_NS) & 1));
      t_dn = rows[(r + 1) % TF3_NS].t[tid + 1];
    }
    if (r == 0 && !halo_up) t_up = t;
    const f4 tl = x > 0 ? cur.t[tid] : t, tr = x + 1 < T ? cur.t[tid + 2] : t;
    float g[3] = {0.f, 0.f, 0.f};
    if (a.g_in) { const float* gi = (const float*)cur.g; g[0] = gi[tid]; g[1] = gi[256 + tid]; g[2] = gi[512 + tid]; }
    else if (l0) { const f4 g0 = cur.g[tid]; g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; }
    if (fold) { const float* cg = c1[(r >> 1) * 128 + (tid >> 1)]; g[0] += cg[0]; g[1] += cg[1]; g[2] += cg[2]; }
    if (l0) *(float4*)(a.g_pyr + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);       // coarser levels: memset after the kernel
    float ex[3] = {cur.ex[0][tid], cur.ex[1][tid], cur.ex[2][tid]};
    if (tv) {                                                                      // tracker.py:526-534
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = chan(t, c), dr = v - chan(tr, c), dd = v - chan(t_dn, c), dl = chan(tl, c) - v, du = chan(t_up, c) - v;
        acc[0] += a.w_tv * (dr * dr + dd * dd);
        g[c] += 2.f * a.w_tv * (dr + dd - dl - du);
      }
    }
    if (want_mask && cur.msk[tid]) {                                               // tracker.py:536-539
#pragma unroll
      for (int c = 0; c < 3; ++c) { acc[1] += a.w_res * ex[c] * ex[c]; g[c] += 2.f * a.w_res * ex[c]; }
    }
    if (a.g_out) {
      if (a.rm) { float* go = a.g_out + ((size_t)y * 3) * T + x; go[0] = g[0]; go[T] = g[1]; go[2 * (size_t)T] = g[2]; }
      else { a.g_out[i] = g[0]; a.g_out[n + i] = g[1]; a.g_out[2 * n + i] = g[2]; }
    }
    f4 o = t;
    if (a.do_adam) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const size_t k = c * n + i;
        float m = 0.9f * cur.m[c][tid] + 0.1f * g[c];
        float v = 0.999f * cur.v[c][tid] + 0.001f * g[c] * g[c];
        a.m[k] = m; a.v[k] = v;
        float upd = (a.lr / bc1) * m / (sqrtf(v) / bc2s + 1e-8f);
        float ne = ex[c] - upd;
        a.extra[k] = ne;
        if (a.ex_band_out) a.ex_band_out[((size_t)(y - a.y_begin) * 3 + c) * T + x] = ne;
        float base = chan(t, c) - ex[c];                                            // painted part
        if (c == 0) o.x = base + ne; else if (c == 1) o.y = base + ne; else o.z = base + ne;
      }
      if (!a.no_pyramid) a.tex_new[i] = o;
    }
    if (r & 1) {                                       // level 1: avg4(A[2y][2x], A[2y+1][2x], A[2y][2x+1], A[2y+1][2x+1]), same order as k_mip_down
      f4 c, d;
      c.x = __shfl_down_sync(0xffffffffu, o_prev.x, 1); c.y = __shfl_down_sync(0xffffffffu, o_prev.y, 1); c.z = __shfl_down_sync(0xffffffffu, o_prev.z, 1);
      d.x = __shfl_down_sync(0xffffffffu, o.x, 1); d.y = __shfl_down_sync(0xffffffffu, o.y, 1); d.z = __shfl_down_sync(0xffffffffu, o.z, 1);
      c.w = d.w = 0.f;
      if (a.do_adam && !a.no_pyramid && a.max_level >= 1 && !(tid & 1))
        a.tex_new[(size_t)a.mip_off[1] + (size_t)(y >> 1) * (T >> 1) + (x >> 1)] = avg4(o_prev, o, c, d);
    }
    o_prev = o;
    t_up = t;
    __syncthreads();                                   // every thread is done with ring slot `slot`: refill it with row r + TF3_NS
    if (tid == 0 && r + TF3_NS <= q_last) tf3_issue(a, rows, bars, slot, x0, y0 + r + TF3_NS, r + TF3_NS < R, l0, want_mask);
  }
  // block partial sums of the two loss terms; the last CTA reduces all partials in a fixed order
  const int w = tid >> 5;
  for (int q = 0; q < 2; ++q) {
    float v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[w * 2 + q] = v;
  }
  __syncthreads();
  if (tid < 2) {
    float s2 = 0.f;
    for (int k = 0; k < 8; ++k) s2 += sh[k * 2 + tid];
    partials[(size_t)blockIdx.x * 2 + tid] = s2;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int q = 0; q < 2; ++q) {
    float s2 = 0.f;
    for (int r = tid; r < (int)gridDim.x; r += 256) s2 += __ldcg(partials + (size_t)r * 2 + q);
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    __syncthreads();
    if (lane == 0) sh[w] = s2;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += sh[k];
      acc_out[q] = t;
    }
  }
  if (tid == 0) *counter = 0u;
}

// launches the fold kernel for `rows` rows starting at a.y_begin: TMA-staged for T >= 256 (VHAP_B200_TEXFOLD=reg: register pipeline)
static void launch_fold_kernel(vhap_ctx* c, const TexFoldArgs& a, int rows, float* acc_out, cudaStream_t s) {
  const int T = c->T, tw = T < 256 ? T : 256, R = T < TF_ROWS ? T : TF_ROWS, nblk = (T / tw) * (rows / R);
  if (T >= 256 && !c->tex_fold_reg) {
    static bool attr_set = false;
    static int pad = 0;
    if (!attr_set) { const char* e = getenv("VHAP_B200_TEXFOLD_PAD_KB"); pad = e ? atoi(e) * 1024 : 0; }      // dev aid: fewer resident CTAs per SM
    const int smem = (int)(TF3_NS * sizeof(TF3Row)) + pad;
    if (!attr_set) { cudaFuncSetAttribute(k_tex_fold3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_set = true; }
    LAUNCH(c, KID_TEX_FOLD, s, vh_launch(k_tex_fold3, nblk, 256, smem, s, a, c->tv_partials, c->tex_counter, acc_out));
  } else {
    LAUNCH(c, KID_TEX_FOLD, s, vh_launch(k_tex_fold2, nblk, 256, 0, s, a, c->tv_partials, c->tex_counter, acc_out));
  }
}

// TV + residual regulariser LOSS VALUES of the current texture (tracker.py:526-539), no gradient: used when the texture update is
// deferred into the next step (the fold kernel then sees the texture one step late), so that a step's loss vector is complete.
__global__ void __launch_bounds__(256) k_tex_reg_loss(const f4* __restrict__ tex, const float* __restrict__ extra, const uint8_t* __restrict__ mask, int T,
                                                      float w_tv, float w_res, float* __restrict__ partials, unsigned* __restrict__ counter, float* __restrict__ out) { VH_PDL_SYNC();
  __shared__ float sh[8 * 2];
  __shared__ bool is_last;
  const size_t n = (size_t)T * T;
  float acc[2] = {0.f, 0.f};
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    int x = (int)(i % T), y = (int)(i / T);
    if (w_tv > 0.f) {
      f4 t = tex[i], tr = x + 1 < T ? tex[i + 1] : t, td = y + 1 < T ? tex[i + T] : t;
      for (int c = 0; c < 3; ++c) { float v = chan(t, c), dr = v - chan(tr, c), dd = v - chan(td, c); acc[0] += w_tv * (dr * dr + dd * dd); }
    }
    if (w_res > 0.f && mask && mask[i]) for (int c = 0; c < 3; ++c) { float e = extra[c * n + i]; acc[1] += w_res * e * e; }
  }
  int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int q = 0; q < 2; ++q) {
    float v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[w * 2 + q] = v;
  }
  __syncthreads();
  if (tid == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < 8; ++k) { s0 += sh[k * 2]; s1 += sh[k * 2 + 1]; }
    partials[(size_t)blockIdx.x * 2] = s0; partials[(size_t)blockIdx.x * 2 + 1] = s1;
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int q = 0; q < 2; ++q) {
    float s2 = 0.f;
    for (int r = tid; r < (int)gridDim.x; r += 256) s2 += __ldcg(partials + (size_t)r * 2 + q);
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    __syncthreads();
    if (lane == 0) sh[w] = s2;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int k = 0; k < 8; ++k) t += sh[k]; out[q] = t; }
  }
  if (tid == 0) *counter = 0u;
}
void launch_tex_reg_loss(vhap_ctx* c, const float* tex_extra, const vhap_stage_cfg* cfg, cudaStream_t s) {
  int T = c->T;
  float sh = cfg->shared_scale;
  float w_tv = (cfg->training && cfg->opt_texture && cfg->w_reg_tex_tv >= 0.f) ? sh * cfg->w_reg_tex_tv / (3.f * (float)(T - 1) * (float)T) : 0.f;
  float w_res = (cfg->training && cfg->opt_texture && cfg->w_reg_tex_res >= 0.f) ? sh * cfg->w_reg_tex_res / (3.f * (float)T * (float)T) : 0.f;
  int nblk = c->tv_nblocks < 148 * 8 ? c->tv_nblocks : 148 * 8;
  LAUNCH(c, KID_TEX_LOSS, s, vh_launch(k_tex_reg_loss, nblk, 256, 0, s, c->mips[c->cur_mip], tex_extra, c->uvmask_res, T, w_tv, w_res, c->tv_partials, c->tex_counter, c->tex_loss));
}

void launch_tex_fold(vhap_ctx* c, float* tex_extra, float* g_out, float* m, float* v, float lr, int step, const vhap_stage_cfg* cfg,
                     float* losses_out, cudaStream_t s) {
  (void)losses_out;
  TexFoldArgs a;
  memset(&a, 0, sizeof(a));
  int T = c->T;
  a.T = T; a.max_level = c->max_level;
  for (int i = 0; i < VH_MAX_MIPS; ++i) a.mip_off[i] = c->mip_off[i];
  a.tex_old = c->mips[c->cur_mip]; a.tex_new = c->mips[c->cur_mip ^ 1]; a.g_pyr = c->g_tex;
  a.extra = tex_extra; a.g_out = g_out; a.m = m; a.v = v; a.g_in = c->tex_apply_grad;
  if (a.g_in) a.g_pyr = nullptr;                       // apply mode: gradient already folded, regularised and reduced across ranks
  a.mask = c->uvmask_res; a.l0_flag = c->tex_l0_flag; a.step_ptr = c->use_dev_step ? c->dev_step : nullptr; a.lr_scale_ptr = c->dev_lr_scale; a.step_bias = c->tex_step_bias;
  float sh = cfg->shared_scale;
  // tv.mean(): (T-1)*T elements per channel, 3 channels (tracker.py:529-533); w already includes scale_factor^2 / ds^2
  a.w_tv = (cfg->training && cfg->opt_texture && cfg->w_reg_tex_tv >= 0.f) ? sh * cfg->w_reg_tex_tv / (3.f * (float)(T - 1) * (float)T) : 0.f;
  a.w_res = (cfg->training && cfg->opt_texture && cfg->w_reg_tex_res >= 0.f) ? sh * cfg->w_reg_tex_res / (3.f * (float)T * (float)T) : 0.f;
  if (a.g_in) { a.w_tv = 0.f; a.w_res = 0.f; }
  a.do_adam = (m != nullptr && v != nullptr) ? 1 : 0;
  a.lr = lr; a.bc1 = 1.f - powf(0.9f, (float)step); a.bc2_sqrt = sqrtf(1.f - powf(0.999f, (float)step));
  a.y_begin = 0; a.y_end = T;
  const int L = c->max_level >= 1 ? 1 : 0;
  launch_fold_kernel(c, a, T, a.g_in ? c->tex_loss + 2 : c->tex_loss, s);      // apply mode: scratch slots
  if (a.g_in) {                                                           // apply mode leaves the gradient pyramid alone
    if (a.do_adam) { c->cur_mip ^= 1; build_mips(c, c->mips[c->cur_mip], s, L); }
    return;
  }
  if (c->g_tex && c->max_level >= 1)                                      // coarser gradient levels (level 0: cleared in-kernel, per region)
    cudaMemsetAsync(c->g_tex + (size_t)c->mip_off[1] * 4, 0, (c->mip_total - c->mip_off[1]) * 4 * sizeof(float), s);
  if (a.do_adam) { c->cur_mip ^= 1; build_mips(c, c->mips[c->cur_mip], s, L); }    // levels 1..L were written by the fold kernel
}

// ---------------------------------------------------------------------------------------------- sharded texture update (data parallel)
// Per step and rank: (1) photometric fold of the local texel-gradient pyramid into a dense row-major gradient (vhap_tex_fold_grad_rm);
// (2) reduce-scatter by row band (NCCL, caller); (3) the band owner adds the total-variation / residual gradients -- rank-invariant, so
// computed ONCE, by the owner, at full weight -- and runs Adam on its T/N rows only (vhap_tex_band_adam: Adam state traffic / N);
// (4) all-gather of the updated bands (NCCL, caller); (5) every rank rebuilds level 0 / 1 + mips from the gathered texture
// (vhap_tex_rebuild_rm).  Replaces fold(+regularisers) -> 50 MB all-reduce -> full-texture Adam on every rank.
static void fill_fold_args(vhap_ctx* c, TexFoldArgs& a, float* tex_extra) {
  memset(&a, 0, sizeof(a));
  a.T = c->T; a.max_level = c->max_level;
  for (int i = 0; i < VH_MAX_MIPS; ++i) a.mip_off[i] = c->mip_off[i];
  a.tex_old = c->mips[c->cur_mip]; a.tex_new = c->mips[c->cur_mip ^ 1]; a.g_pyr = c->g_tex;
  a.extra = tex_extra; a.mask = c->uvmask_res; a.l0_flag = c->tex_l0_flag;
  a.step_ptr = c->use_dev_step ? c->dev_step : nullptr; a.lr_scale_ptr = c->dev_lr_scale; a.step_bias = c->tex_step_bias;
  a.y_begin = 0; a.y_end = c->T;
}
void launch_tex_fold_grad_rm(vhap_ctx* c, float* tex_extra, float* g_rm, cudaStream_t s) {
  TexFoldArgs a; fill_fold_args(c, a, tex_extra);
  a.g_out = g_rm; a.rm = 1; a.do_adam = 0; a.w_tv = 0.f; a.w_res = 0.f; a.mask = nullptr;
  launch_fold_kernel(c, a, c->T, c->tex_loss + 2, s);
  if (c->g_tex && c->max_level >= 1)
    cudaMemsetAsync(c->g_tex + (size_t)c->mip_off[1] * 4, 0, (c->mip_total - c->mip_off[1]) * 4 * sizeof(float), s);
}

int launch_tex_band_adam(vhap_ctx* c, float* tex_extra, const float* g_band, int y_begin, int y_end, float* m, float* v, float lr, int step,
                         const vhap_stage_cfg* cfg, float* ex_band_out, cudaStream_t s) {
  int T = c->T, R = T < TF_ROWS ? T : TF_ROWS;
  if (y_begin < 0 || y_end > T || y_begin >= y_end || (y_begin % R) || ((y_end - y_begin) % R)) return -1;
  TexFoldArgs a; fill_fold_args(c, a, tex_extra);
  a.g_pyr = nullptr; a.g_in = g_band; a.rm = 1; a.y_begin = y_begin; a.y_end = y_end; a.m = m; a.v = v; a.do_adam = 1; a.ex_band_out = ex_band_out; a.no_pyramid = 1;
  // the regularisers are rank-invariant: the band owner computes them, at full weight (no 1/world share)
  a.w_tv = (cfg->training && cfg->opt_texture && cfg->w_reg_tex_tv >= 0.f) ? cfg->w_reg_tex_tv / (3.f * (float)(T - 1) * (float)T) : 0.f;
  a.w_res = (cfg->training && cfg->opt_texture && cfg->w_reg_tex_res >= 0.f) ? cfg->w_reg_tex_res / (3.f * (float)T * (float)T) : 0.f;
  a.lr = lr; a.bc1 = 1.f - powf(0.9f, (float)step); a.bc2_sqrt = sqrtf(1.f - powf(0.999f, (float)step));
  launch_fold_kernel(c, a, y_end - y_begin, c->tex_loss + 2, s);
  return 0;
}

// level 0 (painted + extra) and level 1 of the OTHER pyramid from the gathered row-major texture, planar tex_extra refreshed on the way
__global__ void __launch_bounds__(256) k_tex_rebuild_rm(const float* __restrict__ ex_rm, const float* __restrict__ painted, int T, float* __restrict__ extra,
                                                        f4* __restrict__ lvl0, f4* __restrict__ lvl1, int has_l1) { VH_PDL_SYNC();
  const int tw = T < 256 ? T : 256, tpr = T / tw, tid = threadIdx.x;
  const int x = (blockIdx.x % tpr) * tw + tid, y = (blockIdx.x / tpr) * 2;
  const size_t n = (size_t)T * T;
  f4 o[2] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  if (tid < tw) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const size_t i = (size_t)(y + r) * T + x;
      const float* e = ex_rm + ((size_t)(y + r) * 3) * T + x;
      float e0 = e[0], e1 = e[T], e2 = e[2 * (size_t)T];
      extra[i] = e0; extra[n + i] = e1; extra[2 * n + i] = e2;
      o[r].x = (painted ? painted[i] : 0.f) + e0; o[r].y = (painted ? painted[n + i] : 0.f) + e1; o[r].z = (painted ? painted[2 * n + i] : 0.f) + e2;
      lvl0[i] = o[r];
    }
  }
  f4 c, d;
  c.x = __shfl_down_sync(0xffffffffu, o[0].x, 1); c.y = __shfl_down_sync(0xffffffffu, o[0].y, 1); c.z = __shfl_down_sync(0xffffffffu, o[0].z, 1);
  d.x = __shfl_down_sync(0xffffffffu, o[1].x, 1); d.y = __shfl_down_sync(0xffffffffu, o[1].y, 1); d.z = __shfl_down_sync(0xffffffffu, o[1].z, 1);
  c.w = d.w = 0.f;
  if (tid < tw && has_l1 && !(tid & 1)) lvl1[(size_t)(y >> 1) * (T >> 1) + (x >> 1)] = avg4(o[0], o[1], c, d);
}
void launch_tex_rebuild_rm(vhap_ctx* c, float* tex_extra, const float* ex_rm, cudaStream_t s) {
  int T = c->T, tw = T < 256 ? T : 256;
  f4* pyr = c->mips[c->cur_mip ^ 1];
  int has_l1 = c->max_level >= 1;
  LAUNCH(c, KID_TEX_L0, s, vh_launch(k_tex_rebuild_rm, (T / tw) * (T / 2), 256, 0, s, ex_rm, c->tex_painted, T, tex_extra, pyr, pyr + c->mip_off[has_l1 ? 1 : 0], has_l1));
  c->cur_mip ^= 1;
  build_mips(c, c->mips[c->cur_mip], s, has_l1);
}

__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                       float step_size, float bc2_sqrt) { VH_PDL_SYNC();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  float mi = 0.9f * m[i] + 0.1f * gi, vi = 0.999f * v[i] + 0.001f * gi * gi;
  m[i] = mi; v[i] = vi;
  p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + 1e-8f);
}

struct AdamSegs { int n_seg; int off[24]; int len[24]; float lr[24]; };
__global__ void k_adam_multi(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, AdamSegs sg, int total,
                             float inv_bc1, float bc2_sqrt, const int* __restrict__ step_ptr, const float* __restrict__ lr_scale_ptr) { VH_PDL_SYNC();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (step_ptr) { float st = (float)step_ptr[0]; inv_bc1 = lr_scale_ptr[0] / (1.f - powf(0.9f, st)); bc2_sqrt = sqrtf(1.f - powf(0.999f, st)); }
  int k = 0, base = 0;
  while (k < sg.n_seg - 1 && i >= base + sg.len[k]) { base += sg.len[k]; ++k; }
  int j = sg.off[k] + (i - base);
  float gi = g[j];
  float mi = 0.9f * m[j] + 0.1f * gi, vi = 0.999f * v[j] + 0.001f * gi * gi;
  m[j] = mi; v[j] = vi;
  p[j] -= sg.lr[k] * inv_bc1 * mi / (sqrtf(vi) / bc2_sqrt + 1e-8f);
}
void launch_adam_multi(vhap_ctx* c, float* p, const float* g, float* m, float* v, int n_seg, const int64_t* off, const int64_t* len, const float* lr,
                       int step, cudaStream_t s) {
  AdamSegs sg; memset(&sg, 0, sizeof(sg));
  sg.n_seg = n_seg > 24 ? 24 : n_seg;
  int total = 0;
  for (int k = 0; k < sg.n_seg; ++k) { sg.off[k] = (int)off[k]; sg.len[k] = (int)len[k]; sg.lr[k] = lr[k]; total += (int)len[k]; }
  float bc1 = 1.f - powf(0.9f, (float)step), bc2s = sqrtf(1.f - powf(0.999f, (float)step));
  if (total > 0) LAUNCH(c, KID_ADAM, s, vh_launch(k_adam_multi, (total + 255) / 256, 256, 0, s, p, g, m, v, sg, total, 1.f / bc1, bc2s, c->use_dev_step ? c->dev_step : nullptr, c->dev_lr_scale));
}

void launch_adam(vhap_ctx* c, float* p, const float* g, float* m, float* v, int64_t n, float lr, int step, cudaStream_t s) {
  float bc1 = 1.f - powf(0.9f, (float)step), bc2s = sqrtf(1.f - powf(0.999f, (float)step));
  LAUNCH(c, KID_ADAM, s, vh_launch(k_adam, (unsigned)((n + 255) / 256), 256, 0, s, p, g, m, v, n, lr / bc1, bc2s));
}
