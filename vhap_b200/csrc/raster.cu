// Tile-binned triangle rasteriser (replaces dr.rasterize, vhap/util/render_nvdiffrast.py:254).
// Specification: oracle/raster.py rasterize_ids (exact integer coverage, fp32 depth plane, lower-id tie break); the
// triangle ids produced here are bit-identical to it.  Pipeline per call:
//   k_snap (only for caller-provided clip positions; the fused path snaps inside k_skin_fwd)
//   k_bin<COUNT> -> k_scan -> k_bin<FILL> : per-frame 16x16-pixel tile lists of triangle ids
//   k_fine : one CTA (4 warps) per tile; lanes set up triangles into shared memory, each warp owns an 8x8 sub-tile and
//            uses __ballot_sync to skip triangles whose bounding box misses it; every lane tests 2 pixels with int32 edge
//            functions relative to the tile origin (saturated so the sign is exact) and keeps the nearest hit in registers.
#include "engine.h"

#define SNAP_GUARD 131072.f
#define FINE_CHUNK 128

__global__ void k_snap(const f4* __restrict__ clip, i4* __restrict__ snap, float* __restrict__ ndc, int n, int H, int W) { VH_PDL_SYNC();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  f4 cl = clip[i];
  bool valid = isfinite(cl.x) && isfinite(cl.y) && isfinite(cl.z) && isfinite(cl.w) && cl.w > 0.f;
  i4 sn = {0, 0, 0, 0};
  if (valid) {
    float sx = __fmul_rn(__fdiv_rn(cl.x, cl.w), (float)(W * 8));
    float sy = __fmul_rn(__fdiv_rn(cl.y, cl.w), (float)(H * 8));
    sx = fminf(fmaxf(sx, -SNAP_GUARD), SNAP_GUARD); sy = fminf(fmaxf(sy, -SNAP_GUARD), SNAP_GUARD);
    sn.x = __float2int_rn(sx); sn.y = __float2int_rn(sy);
    sn.z = __float_as_int(__fdiv_rn(cl.z, cl.w)); sn.w = 1;
  }
  snap[i] = sn;
  if (ndc) { ndc[2 * i] = cl.x / cl.w; ndc[2 * i + 1] = cl.y / cl.w; }
}

__device__ __forceinline__ int floordiv16(int a) { return a >> 4; }       // arithmetic shift = floor for negatives
__device__ __forceinline__ int ceildiv16(int a) { return -((-a) >> 4); }

// pixel bounding box of the samples a triangle can cover; false = culled
__device__ __forceinline__ bool tri_bbox(const i4& s0, const i4& s1, const i4& s2, int H, int W, int cull_backface,
                                         int& px0, int& px1, int& py0, int& py1, long long& area2) {
  if (!(s0.w && s1.w && s2.w)) return false;
  long long d1x = s1.x - s0.x, d1y = s1.y - s0.y, d2x = s2.x - s0.x, d2y = s2.y - s0.y;
  area2 = d1x * d2y - d2x * d1y;
  if (area2 == 0) return false;
  if (cull_backface && area2 < 0) return false;
  int mnx = min(s0.x, min(s1.x, s2.x)), mxx = max(s0.x, max(s1.x, s2.x));
  int mny = min(s0.y, min(s1.y, s2.y)), mxy = max(s0.y, max(s1.y, s2.y));
  px0 = max(ceildiv16(mnx + W * 8 - 8), 0); px1 = min(floordiv16(mxx + W * 8 - 8), W - 1);
  py0 = max(ceildiv16(mny + H * 8 - 8), 0); py1 = min(floordiv16(mxy + H * 8 - 8), H - 1);
  return px0 <= px1 && py0 <= py1;
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_bin(const i4* __restrict__ snap, const i4* __restrict__ faces, int V, int F, int B, int H, int W,
                                             int tiles_x, int tiles_y, int cull_backface, int* __restrict__ tile_count,
                                             const int* __restrict__ tile_off, int* __restrict__ tile_cursor, int* __restrict__ tile_list,
                                             int tile_cap, int* __restrict__ overflow) { VH_PDL_SYNC();
  int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (t >= F) return;
  i4 f = faces[t];
  const i4* sb = snap + (size_t)b * V;
  i4 s0 = sb[f.x], s1 = sb[f.y], s2 = sb[f.z];
  int px0, px1, py0, py1; long long area2;
  if (!tri_bbox(s0, s1, s2, H, W, cull_backface, px0, px1, py0, py1, area2)) return;
  int tx0 = px0 / VH_TILE, tx1 = px1 / VH_TILE, ty0 = py0 / VH_TILE, ty1 = py1 / VH_TILE;
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) {
      int tile = (b * tiles_y + ty) * tiles_x + tx;
      if (!FILL) atomicAdd(tile_count + tile, 1);
      else {
        int pos = tile_off[tile] + atomicAdd(tile_cursor + tile, 1);
        if (pos < tile_cap) tile_list[pos] = t; else *overflow = 1;
      }
    }
}

// exclusive scan of n ints in three small launches: 1024-thread blocks scan 4096 coalesced elements each and emit
// their totals, one block scans the totals, the totals are added back (n up to 4 Mi elements)
__global__ void __launch_bounds__(1024) k_scan_local(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ aux, int n) { VH_PDL_SYNC();
  __shared__ int sh[1024];
  int i0 = (blockIdx.x * 1024 + threadIdx.x) * 4;
  int v[4], s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[k] = (i0 + k < n) ? in[i0 + k] : 0; s += v[k]; }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  int excl = sh[threadIdx.x] - s;
#pragma unroll
  for (int k = 0; k < 4; ++k) { if (i0 + k < n) out[i0 + k] = excl; excl += v[k]; }
  if (threadIdx.x == 1023) aux[blockIdx.x] = sh[1023];
}
__global__ void __launch_bounds__(1024) k_scan_aux(int* __restrict__ aux, int naux, int* __restrict__ total) { VH_PDL_SYNC();
  __shared__ int sh[1024];
  int s = threadIdx.x < naux ? aux[threadIdx.x] : 0;
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  if (threadIdx.x < naux) aux[threadIdx.x] = sh[threadIdx.x] - s;
  if (threadIdx.x == 1023 && total) *total = sh[1023];
}
__global__ void __launch_bounds__(1024) k_scan_add(int* __restrict__ out, const int* __restrict__ aux, int n) { VH_PDL_SYNC();
  int i0 = (blockIdx.x * 1024 + threadIdx.x) * 4, a = aux[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 4; ++k) if (i0 + k < n) out[i0 + k] += a;
}
// Single-launch exclusive scan (decoupled look-back): blocks take a ticket (dynamic block id = forward progress), scan their
// 4096 elements, publish (flag | aggregate) as ONE 64-bit word and the first warp looks back over the predecessors' words, 32 at
// a time, until it meets an inclusive prefix.  state[0..nb) (words), state[nb] (ticket) and state[nb + 1] (finished blocks) are zero on
// entry and are cleared again by the block that finishes last, so consecutive scans need no clear launch.  pad4: the input is rounded
// up to multiples of 4 on load (tile list starts are padded to 16 bytes for the TMA copies of the fine rasteriser).
#define SCAN_FLAG_A 1ull
#define SCAN_FLAG_P 2ull
__global__ void __launch_bounds__(1024) k_scan_lb(const int* __restrict__ in, int* __restrict__ out, int n, unsigned long long* state, int nb, int* __restrict__ total,
                                                   int pad4) { VH_PDL_SYNC();
  __shared__ int sh[32];
  __shared__ int s_bid, s_prefix;
  if (threadIdx.x == 0) s_bid = (int)atomicAdd((unsigned*)(state + nb), 1u);
  __syncthreads();
  const int bid = s_bid, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int i0 = (bid * 1024 + threadIdx.x) * 4;
  int v[4], s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[k] = (i0 + k < n) ? in[i0 + k] : 0; if (pad4) v[k] = (v[k] + 3) & ~3; s += v[k]; }
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) sh[w] = incl;
  __syncthreads();
  if (w == 0) {
    int x = sh[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
    sh[lane] = x;
  }
  __syncthreads();
  incl += w > 0 ? sh[w - 1] : 0;
  const int block_total = sh[31];
  if (w == 0) {
    if (lane == 0) atomicExch(state + bid, ((bid == 0 ? SCAN_FLAG_P : SCAN_FLAG_A) << 32) | (unsigned)block_total);
    int prefix = 0;
    int j = bid - 1;                                   // nearest predecessor of this window
    while (j >= 0) {
      int idx = j - lane;
      unsigned long long st = SCAN_FLAG_P << 32;       // lanes before block 0 act as an empty inclusive prefix
      if (idx >= 0) { do { st = *(volatile unsigned long long*)(state + idx); } while ((st >> 32) == 0ull); }
      unsigned pm = __ballot_sync(0xffffffffu, (st >> 32) == SCAN_FLAG_P);
      int first = pm ? __ffs(pm) - 1 : 32;             // nearest predecessor holding an inclusive prefix
      int val = (lane <= first && idx >= 0) ? (int)(unsigned)st : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
      prefix += val;
      if (pm) break;
      j -= 32;
    }
    if (lane == 0) {
      if (bid > 0) atomicExch(state + bid, (SCAN_FLAG_P << 32) | (unsigned)(prefix + block_total));
      s_prefix = prefix;
      if (bid == nb - 1 && total) *total = prefix + block_total;
    }
  }
  __syncthreads();
  int excl = s_prefix + incl - s;
#pragma unroll
  for (int k = 0; k < 4; ++k) { if (i0 + k < n) out[i0 + k] = excl; excl += v[k]; }
  // self-reset: every block has finished its look-back when the last one arrives here
  __shared__ int s_last;
  if (threadIdx.x == 0) { __threadfence(); s_last = atomicAdd((unsigned*)(state + nb + 1), 1u) == (unsigned)(nb - 1); }
  __syncthreads();
  if (s_last) for (int j = threadIdx.x; j < nb + 2; j += 1024) state[j] = 0ull;
}

// tile list starts are padded to 4 ints (16 bytes) so that the fine rasteriser can stage them with TMA bulk copies
__global__ void k_pad_counts(const int* __restrict__ cnt, int* __restrict__ padded, int n) { VH_PDL_SYNC();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) padded[i] = (cnt[i] + 3) & ~3;
}

void launch_scan(vhap_ctx* c, const int* in, int* out, int n, int* total, cudaStream_t s, int pad4) {
  int nb = (n + 4095) / 4096;
  static const bool three_pass = getenv("VHAP_B200_SCAN3") != nullptr;       // previous three-launch scan, kept for A/B timing
  if (three_pass || nb > VH_SCAN_MAX_BLOCKS) {
    if (pad4) { LAUNCH(c, KID_SCAN, s, vh_launch(k_pad_counts, (n + 255) / 256, 256, 0, s, in, out, n)); in = out; }     // (in place: local scan reads before it writes)
    LAUNCH(c, KID_SCAN, s, vh_launch(k_scan_local, nb, 1024, 0, s, in, out, c->scan_aux, n));
    LAUNCH(c, KID_SCAN, s, vh_launch(k_scan_aux, 1, 1024, 0, s, c->scan_aux, nb, total));
    LAUNCH(c, KID_SCAN, s, vh_launch(k_scan_add, nb, 1024, 0, s, out, c->scan_aux, n));
    return;
  }
  LAUNCH(c, KID_SCAN, s, vh_launch(k_scan_lb, nb, 1024, 0, s, in, out, n, c->scan_state, nb, total, pad4));
}

struct FineTri {        // shared-memory record, struct of arrays
  int e[3][FINE_CHUNK], ea[3][FINE_CHUNK], eb[3][FINE_CHUNK];
  float zx[FINE_CHUNK], zy[FINE_CHUNK], zc[FINE_CHUNK];
  int id[FINE_CHUNK];
  unsigned bbox[FINE_CHUNK];    // tile-relative pixel bbox packed x0 | x1<<8 | y0<<16 | y1<<24 ; 0xffffffff = skip
};

__device__ __forceinline__ int sat30(long long v) {
  const long long L = 1ll << 30;
  return (int)(v > L ? L : (v < -L ? -L : v));
}

__global__ void __launch_bounds__(128) k_fine(const i4* __restrict__ snap, const i4* __restrict__ faces, const int* __restrict__ tile_count,
                                              const int* __restrict__ tile_off, const int* __restrict__ tile_list, int tile_cap,
                                              int V, int H, int W, int tiles_x, int tiles_y, int cull_backface, int* __restrict__ tri_id) { VH_PDL_SYNC();
  __shared__ FineTri T;
  __shared__ __align__(16) int ids_s[2][FINE_CHUNK];     // triangle-id chunks staged by TMA (cp.async.bulk), double buffered
  __shared__ uint64_t bar[2];
  int tile = blockIdx.x;
  int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  int n = tile_count[tile], off = tile_off[tile];
  if (off + n > tile_cap) n = max(tile_cap - off, 0);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int sx0 = (warp & 1) * 8, sy0 = (warp >> 1) * 8;              // sub-tile origin inside the tile
  int lx = sx0 + (lane & 7), ly0 = sy0 + (lane >> 3), ly1 = ly0 + 4;
  int X0 = tx * VH_TILE, Y0 = ty * VH_TILE;
  int gx = X0 + lx, gy0 = Y0 + ly0, gy1 = Y0 + ly1;
  float Cxf = (float)((2 * gx + 1 - W) * 8), Cy0f = (float)((2 * gy0 + 1 - H) * 8), Cy1f = (float)((2 * gy1 + 1 - H) * 8);
  float bz0 = INFINITY, bz1 = INFINITY; int bi0 = 0, bi1 = 0;
  const i4* sb = snap + (size_t)b * V;
  long long Cx0 = (long long)(2 * X0 + 1 - W) * 8, Cy0 = (long long)(2 * Y0 + 1 - H) * 8;   // sample of the tile's first pixel
  // TMA: one elected thread arms an mbarrier with the byte count and issues a 1-D bulk copy global -> shared of the next chunk
  // of the tile's triangle list; the copy of chunk k+1 overlaps the set-up and coverage tests of chunk k
  auto issue = [&](int base, int buf) {
    int cnt = min(FINE_CHUNK, n - base);
    uint32_t bytes = (uint32_t)(((cnt + 3) >> 2) << 4);
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[buf]), dst = (uint32_t)__cvta_generic_to_shared(&ids_s[buf][0]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(tile_list + off + base), "r"(bytes), "r"(bar_a) : "memory");
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (n > 0) issue(0, 0);
  }
  __syncthreads();
  int chunk = 0;
  for (int base = 0; base < n; base += FINE_CHUNK, ++chunk) {
    int j = threadIdx.x, cnt = min(FINE_CHUNK, n - base), buf = chunk & 1;
    __syncthreads();                                   // previous chunk fully consumed (its buffer and T may be overwritten)
    if (threadIdx.x == 0 && base + FINE_CHUNK < n) issue(base + FINE_CHUNK, buf ^ 1);
    {
      uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[buf]), parity = (chunk >> 1) & 1, done = 0;
      while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_a), "r"(parity) : "memory");
    }
    if (j < cnt) {
      int t = ids_s[buf][j];
      i4 f = faces[t];
      i4 s0 = sb[f.x], s1 = sb[f.y], s2 = sb[f.z];
      int px0, px1, py0, py1; long long area2;
      unsigned bb = 0xffffffffu;
      if (tri_bbox(s0, s1, s2, H, W, cull_backface, px0, px1, py0, py1, area2)) {
        int rx0 = max(px0 - X0, 0), rx1 = min(px1 - X0, VH_TILE - 1), ry0 = max(py0 - Y0, 0), ry1 = min(py1 - Y0, VH_TILE - 1);
        if (rx0 <= rx1 && ry0 <= ry1) {
          bb = (unsigned)rx0 | ((unsigned)rx1 << 8) | ((unsigned)ry0 << 16) | ((unsigned)ry1 << 24);
          int sgn = area2 > 0 ? 1 : -1;
          const i4 sv[3] = {s0, s1, s2};
          for (int k = 0; k < 3; ++k) {                       // edge k: v_{k+1} -> v_{k+2}
            const i4& A = sv[(k + 1) % 3]; const i4& Bv = sv[(k + 2) % 3];
            long long dx = (long long)sgn * (Bv.x - A.x), dy = (long long)sgn * (Bv.y - A.y);
            long long E = dx * (Cy0 - A.y) - dy * (Cx0 - A.x);
            bool tl = (dy < 0) || (dy == 0 && dx < 0);
            if (!tl) E -= 1;                                   // E > 0  <=>  E - 1 >= 0
            T.e[k][j] = sat30(E);
            T.ea[k][j] = (int)(-16 * dy);                      // step per pixel in x
            T.eb[k][j] = (int)(16 * dx);                       // step per pixel in y
          }
          // fp32 depth plane, every operation individually rounded (no FMA), see oracle/raster.py
          float z0 = __int_as_float(s0.z), z1 = __int_as_float(s1.z), z2 = __int_as_float(s2.z);
          float f1x = (float)(s1.x - s0.x), f1y = (float)(s1.y - s0.y), f2x = (float)(s2.x - s0.x), f2y = (float)(s2.y - s0.y);
          float dz1 = __fsub_rn(z1, z0), dz2 = __fsub_rn(z2, z0);
          float af = __ll2float_rn(area2);
          float zx = __fdiv_rn(__fsub_rn(__fmul_rn(dz1, f2y), __fmul_rn(dz2, f1y)), af);
          float zy = __fdiv_rn(__fsub_rn(__fmul_rn(dz2, f1x), __fmul_rn(dz1, f2x)), af);
          float zc = __fsub_rn(__fsub_rn(z0, __fmul_rn(zx, (float)s0.x)), __fmul_rn(zy, (float)s0.y));
          T.zx[j] = zx; T.zy[j] = zy; T.zc[j] = zc; T.id[j] = t + 1;
        }
      }
      T.bbox[j] = bb;
    }
    __syncthreads();
    for (int g = 0; g < cnt; g += 32) {
      int q = g + lane;
      bool hit = false;
      if (q < cnt) {
        unsigned bb = T.bbox[q];
        if (bb != 0xffffffffu) {
          int rx0 = bb & 255, rx1 = (bb >> 8) & 255, ry0 = (bb >> 16) & 255, ry1 = bb >> 24;
          hit = rx0 <= sx0 + 7 && rx1 >= sx0 && ry0 <= sy0 + 7 && ry1 >= sy0;
        }
      }
      unsigned mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        int k = g + __ffs(mask) - 1;
        mask &= mask - 1;
        int e0 = T.e[0][k] + T.ea[0][k] * lx, e1 = T.e[1][k] + T.ea[1][k] * lx, e2 = T.e[2][k] + T.ea[2][k] * lx;
        int b0 = T.eb[0][k], b1 = T.eb[1][k], b2 = T.eb[2][k];
        float zx = T.zx[k], zy = T.zy[k], zc = T.zc[k];
        int id = T.id[k];
        float zxx = __fmul_rn(zx, Cxf);
        if ((e0 + b0 * ly0 | e1 + b1 * ly0 | e2 + b2 * ly0) >= 0) {
          float z = __fadd_rn(__fadd_rn(zxx, __fmul_rn(zy, Cy0f)), zc);
          if (z >= -1.f && z <= 1.f && (z < bz0 || (z == bz0 && id < bi0))) { bz0 = z; bi0 = id; }
        }
        if ((e0 + b0 * ly1 | e1 + b1 * ly1 | e2 + b2 * ly1) >= 0) {
          float z = __fadd_rn(__fadd_rn(zxx, __fmul_rn(zy, Cy1f)), zc);
          if (z >= -1.f && z <= 1.f && (z < bz1 || (z == bz1 && id < bi1))) { bz1 = z; bi1 = id; }
        }
      }
    }
  }
  if (gx < W) {
    if (gy0 < H) tri_id[((size_t)b * H + gy0) * W + gx] = bi0;
    if (gy1 < H) tri_id[((size_t)b * H + gy1) * W + gx] = bi1;
  }
}

void launch_raster(vhap_ctx* c, const f4* clip, i4* snap, int B, int H, int W, int* tri_id, int cull_backface, int need_snap, cudaStream_t s, bool zeroed) {
  int V = c->V, F = c->F;
  if (need_snap) LAUNCH(c, KID_SNAP, s, vh_launch(k_snap, (B * V + 255) / 256, 256, 0, s, clip, snap, c->ndc, B * V, H, W));
  int tiles_x = (W + VH_TILE - 1) / VH_TILE, tiles_y = (H + VH_TILE - 1) / VH_TILE, ntiles = B * tiles_x * tiles_y;
  if (!zeroed) { VhZeroSegs z; z.n = 2; z.p[0] = c->tile_count; z.p[1] = c->tile_cursor; z.bytes[0] = z.bytes[1] = sizeof(int) * ntiles; vh_zero_multi(c, z, s); }
  dim3 g((F + 255) / 256, B);
  LAUNCH(c, KID_BIN, s, vh_launch(k_bin<false>, g, 256, 0, s, snap, c->faces, V, F, B, H, W, tiles_x, tiles_y, cull_backface, c->tile_count, nullptr, nullptr, nullptr, 0, c->overflow_flag));
  launch_scan(c, c->tile_count, c->tile_off, ntiles, nullptr, s, 1);          // starts of the tile lists, padded to 4 entries
  LAUNCH(c, KID_BIN, s, vh_launch(k_bin<true>, g, 256, 0, s, snap, c->faces, V, F, B, H, W, tiles_x, tiles_y, cull_backface, c->tile_count, c->tile_off, c->tile_cursor, c->tile_list,
                                c->tile_cap, c->overflow_flag));
  LAUNCH(c, KID_FINE, s, vh_launch(k_fine, ntiles, 128, 0, s, snap, c->faces, c->tile_count, c->tile_off, c->tile_list, c->tile_cap, V, H, W, tiles_x, tiles_y, cull_backface, tri_id));
}

// dr.rasterize's float outputs for the modular API: rast = (u, v, z/w, id), rast_db = (du/dx, du/dy, dv/dx, dv/dy)
__global__ void k_rast_out(RenderArgs A, float* __restrict__ rast, float* __restrict__ rast_db) { VH_PDL_SYNC();
  size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)A.B * A.H * A.W;
  if (pix >= n) return;
  int x = pix % A.W, y = (pix / A.W) % A.H, b = pix / ((size_t)A.W * A.H);
  int id = A.tri_id[pix];
  float4 r = {0, 0, 0, 0}, d = {0, 0, 0, 0};
  if (id > 0) {
    TriSetup s;
    tri_setup(A, b, x, y, id - 1, s);
    r = make_float4(s.b0, s.b1, s.zw, (float)id);
    d = make_float4(s.dudx, s.dudy, s.dvdx, s.dvdy);
  }
  if (rast) ((float4*)rast)[pix] = r;
  if (rast_db) ((float4*)rast_db)[pix] = d;
}

void launch_rast_out(vhap_ctx* c, const f4* clip, int B, int H, int W, const int* tri_id, float* rast, float* rast_db, cudaStream_t s) {
  RenderArgs A;
  memset(&A, 0, sizeof(A));
  A.B = B; A.H = H; A.W = W; A.V = c->V; A.F = c->F; A.faces = c->faces; A.clip = clip; A.tri_id = tri_id;
  size_t n = (size_t)B * H * W;
  LAUNCH(c, KID_RAST_OUT, s, vh_launch(k_rast_out, (unsigned)((n + 255) / 256), 256, 0, s, A, rast, rast_db));
}
