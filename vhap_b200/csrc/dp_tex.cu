// Data-parallel texture update over peer memory (NVLink / NVSwitch), no collective library on the path.
//
// Per step and rank (row-major exchange layout [T][3][T], rank r owns the rows [r T/N, (r+1) T/N)):
//   fold          local photometric fold of the texel-gradient pyramid -> g_rm                       (texture.cu, k_tex_fold3)
//   barrier A     every rank's g_rm is complete                                                       (mailbox flags, one-warp kernel k_dp_barrier)
//   reduce        g_band = sum over ranks of g_rm[band]: ONE multimem.ld_reduce.add.v4.f32 per 16 bytes through the NVSwitch multicast
//                 mapping of g_rm (in-switch reduction, NVLS) -- or, without multicast support, a loop of peer loads in rank order
//   band Adam     + TV / residual gradients, Adam on the band's rows of tex_extra / m / v             (texture.cu, vhap_tex_band_adam)
//   broadcast     the updated rows go to every rank's ex_rm with multimem.st (one store, the switch replicates) -- or peer stores
//   barrier B     every band has landed everywhere                                                    (k_dp_barrier)
//   rebuild       level 0 / 1 + mips of the new pyramid from ex_rm, planar tex_extra refreshed       (texture.cu, vhap_tex_rebuild_rm)
// The symmetric buffers (g_rm, ex_rm) and their multicast mappings are allocated by the caller (torch symmetric memory) and handed over
// as raw pointers (vhap_dp_tex_connect); the barriers use the CUDA-IPC mailboxes of vhap_dp_init / vhap_dp_connect.
// Replaces NCCL reduce-scatter + all-gather of parallel.TexShardComm; the reference has no multi-GPU path.
#include "engine.h"

// Barriers: one-warp kernels on the mailbox flags (a waiting rank must not occupy the machine: a version with the wait folded into the
// many-CTA reduce kernel starved the forward chain that runs beside the texture update -- vertex normals 0.026 -> 0.087 ms, r7 timeline).
//   A  "every rank's g_rm is complete", B  "every band has landed everywhere".
// Reuse is safe: a rank overwrites its g_rm (next fold) only after it passed barrier B, i.e. after every peer finished reducing; a rank
// writes into a peer's ex_rm (next broadcast) only after that peer passed barrier A again, i.e. after its rebuild of this step.
#define DP_FLAG_OFF(which) ((size_t)2 * VH_DP_MAX * 8 + (size_t)(1 + (which)) * VH_DP_MAX)      // after the slab slots and the slab flags
__global__ void k_dp_barrier(float* mine, float* const* peers, int rank, int world, int which, int* epoch, int* err, unsigned long long* wait) { VH_PDL_SYNC();
  if (threadIdx.x != 0) return;
  const int e = epoch[which] + 1;
  epoch[which] = e;
  __threadfence_system();
  for (int p = 0; p < world; ++p) ((volatile int*)(peers[p] + DP_FLAG_OFF(which)))[rank] = e;
  volatile const int* flags = (volatile const int*)(mine + DP_FLAG_OFF(which));
  const long long t0 = clock64();
  unsigned long long w0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(w0));
  for (int j = 0; j < world; ++j)
    while (flags[j] - e < 0) { if (clock64() - t0 > (1ll << 33)) { *err = 2; break; } }      // ~4 s: a dead peer must not hang the GPU
  __threadfence_system();
  if (wait) { unsigned long long w1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(w1)); wait[1 + which] += w1 - w0; }
}

__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// g_band[i] = sum_r g_rm_r[band_off + i]   (float4 granularity)
__global__ void __launch_bounds__(256) k_dp_reduce_band(const float* mc, float* const* peers, int world, size_t band_off4, size_t n4, float4* __restrict__ out) { VH_PDL_SYNC();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (mc) {
    // four independent in-switch reductions in flight per thread: the load-reduce round trip through the NVSwitch is several microseconds
    // (one per thread at 148 CTAs measured 0.21 TB/s, r10 timeline)
    for (; i + 3 * stride < n4; i += 4 * stride) {
      float4 a = mc_ld_reduce(mc + (band_off4 + i) * 4), b = mc_ld_reduce(mc + (band_off4 + i + stride) * 4);
      float4 c = mc_ld_reduce(mc + (band_off4 + i + 2 * stride) * 4), d = mc_ld_reduce(mc + (band_off4 + i + 3 * stride) * 4);
      out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
    }
    for (; i < n4; i += stride) out[i] = mc_ld_reduce(mc + (band_off4 + i) * 4);
    return;
  }
  for (; i < n4; i += stride) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      float4 v = __ldcv((const float4*)peers[r] + band_off4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[i] = s;
  }
}
// every rank's ex_rm[band_off + i] = ex_band[i]
__global__ void __launch_bounds__(256) k_dp_bcast_band(float* mc, float* const* peers, int world, size_t band_off4, size_t n4, const float4* __restrict__ in) { VH_PDL_SYNC();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = in[i];
    if (mc) mc_st(mc + (band_off4 + i) * 4, v);
    else for (int r = 0; r < world; ++r) ((float4*)peers[r])[band_off4 + i] = v;
  }
}

void launch_dp_barrier(vhap_ctx* c, int which, cudaStream_t s) {
  LAUNCH(c, KID_MISC, s, vh_launch(k_dp_barrier, 1, 32, 0, s, c->dp_box, c->dp_peers_dev, c->dp_rank, c->dp_world, which, c->dp_epoch + 2, c->dp_err, c->dp_wait));
}
void launch_dp_reduce_band(vhap_ctx* c, float* g_band, cudaStream_t s) {
  const size_t nb4 = (size_t)3 * c->T * c->T / c->dp_world / 4;
  const int grid = (int)((nb4 + 1023) / 1024 < 148 * 4 ? (nb4 + 1023) / 1024 : 148 * 4);
  LAUNCH(c, KID_MISC, s, vh_launch(k_dp_reduce_band, grid, 256, 0, s, c->dp_grm_mc, c->dp_grm_peers_dev, c->dp_world, nb4 * c->dp_rank, nb4, (float4*)g_band));
}
void launch_dp_bcast_band(vhap_ctx* c, const float* ex_band, cudaStream_t s) {
  const size_t nb4 = (size_t)3 * c->T * c->T / c->dp_world / 4;
  const int grid = (int)((nb4 + 255) / 256 < 148 * 4 ? (nb4 + 255) / 256 : 148 * 4);
  LAUNCH(c, KID_MISC, s, vh_launch(k_dp_bcast_band, grid, 256, 0, s, c->dp_exrm_mc, c->dp_exrm_peers_dev, c->dp_world, nb4 * c->dp_rank, nb4, (const float4*)ex_band));
}
