// Blend-shape contraction on the 5th-generation tensor cores (tcgen05 + TMEM), used in both directions:
//   forward   v_shaped[b][m] = template[m] + offset[m] + sum_k S[m][k] beta[b][k]          (lbs.py:218-239, flame.py:602-608)
//   backward  g_beta[b][k]   = sum_m S[m][k] g_vshaped[b][m]                                 (autograd of the einsum)
// Both are D[rows x 16] = A[rows x red] . B[16 x red]^T with the frame batch as the N = 16 dimension of one UMMA
// (M = 128, N = 16, K = 8 per instruction, kind::tf32, fp32 accumulation in tensor memory).
//
// Precision: a single TF32 product carries ~5e-4 relative error, above the 1e-4 parity bar, so every product is done as
// 3xTF32: x = x_hi + x_lo with x_hi = the tf32 truncation the tensor core applies to the raw fp32 bits and x_lo = x - x_hi
// (exact in fp32, fits tf32 up to 2^-21); D += A_hi B_hi + A_lo B_hi + A_hi B_lo.  The contraction is HBM-bound
// (~25 MB of bases per launch, arithmetic intensity ~B/2 flop/byte), the two extra MMA passes are free.
//
// Data movement: the A tile (128 rows x 32 fp32 = one 128-byte swizzle span per row) is brought in with 16-byte cp.async
// (LDGSTS) copies written directly in the SWIZZLE_128B K-major layout the UMMA shared-memory descriptors expect
// (chunk c of row r goes to chunk c ^ (r & 7) of its 1024-byte 8-row atom); a 3-stage ring is recycled through mbarriers
// that tcgen05.commit arrives on.  One elected thread issues the MMAs; four warps read the accumulator back with tcgen05.ld.
#include "engine.h"

#define TC_ROWS 128          // UMMA M
#define TC_N 16              // UMMA N = frames per launch slice
#define TC_KB 32             // fp32 elements per stage along the reduction (= 128 bytes = one swizzle span)
#define TC_STAGES 3
#define TC_A_BYTES (TC_ROWS * TC_KB * 4)     // 16 KB
#define TC_B_BYTES (TC_N * TC_KB * 4)        // 2 KB
#define TC_STAGE_BYTES (2 * TC_A_BYTES + 2 * TC_B_BYTES)
#define TC_SMEM_BYTES (TC_STAGES * TC_STAGE_BYTES + 1024)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, addr = smem_u32(bar);
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {   // src_bytes < 16 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// UMMA shared-memory descriptor: K-major operand, SWIZZLE_128B, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address        bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major) bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset: 8-row atom pitch            bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::tf32: D = f32, A = B = tf32, both K-major, M = 128, N = 16
__device__ __forceinline__ uint32_t umma_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format  = F32
  d |= 2u << 7;                       // a_format  = TF32
  d |= 2u << 10;                      // b_format  = TF32
  d |= (uint32_t)(TC_N >> 3) << 17;   // n_dim
  d |= (uint32_t)(TC_ROWS >> 4) << 24;// m_dim
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[row0 .. row0+128) x 16 = sum over red in [r0, r1) of A[row][red] * Bm[n][red]
//   A:  [rows][lda] fp32, lda * 4 bytes a multiple of 16;  Bm: [nb][ldb] fp32 (nb <= 16 valid rows), ldb * 4 multiple of 16
//   mode 0 (forward): out[n * ldo + row] = add0[row] + add1[row] + D   (plain store, one CTA owns the whole reduction)
//   mode 1 (backward): atomicAdd(out[n * ldo + row], D)                (reduction split over blockIdx.y)
__global__ void __launch_bounds__(128, 1) k_blend_tc(const float* __restrict__ A, int lda, int rows, const float* __restrict__ Bm, int ldb, int nb,
                                                     int nred, int red_per_cta, int mode, const float* __restrict__ add0,
                                                     const float* __restrict__ add1, float* __restrict__ out, int ldo) { VH_PDL_SYNC();
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_free[TC_STAGES];     // stage may be overwritten (its MMAs completed)
  __shared__ uint64_t bar_done;                // accumulator complete
  __shared__ uint32_t tmem_base_sh;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B needs 1024-byte alignment
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TC_ROWS;
  const int r0 = blockIdx.y * red_per_cta, r1 = min(r0 + red_per_cta, nred);
  const int nstage = (r1 - r0 + TC_KB - 1) / TC_KB;

  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) mbar_init(&bar_free[i], 1);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {     // allocate 32 TMEM columns (the accumulator uses 16)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_sh;
  const uint32_t idesc = umma_idesc();

  auto stage_ptr = [&](int s) { return smem + (size_t)s * TC_STAGE_BYTES; };
  // issue the cp.async copies of reduction block kb into stage s (A raw + B raw); out-of-range rows / columns are zero-filled
  auto load_stage = [&](int kb, int s) {
    uint8_t* sp = stage_ptr(s);
    uint32_t a_s = smem_u32(sp), b_s = smem_u32(sp + 2 * TC_A_BYTES);
    int red0 = r0 + kb * TC_KB;
    for (int i = tid; i < TC_ROWS * 8; i += 128) {            // 128 rows x 8 chunks of 16 bytes
      int r = i >> 3, c = i & 7;
      int grow = row0 + r, gred = red0 + c * 4;
      int valid = (grow < rows && gred < r1) ? min(16, (r1 - gred) * 4) : 0;
      const float* src = A + (size_t)(grow < rows ? grow : 0) * lda + (gred < nred ? gred : 0);
      cp_async16(a_s + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4), src, valid);
    }
    for (int i = tid; i < TC_N * 8; i += 128) {
      int r = i >> 3, c = i & 7;
      int gred = red0 + c * 4;
      int valid = (r < nb && gred < r1) ? min(16, (r1 - gred) * 4) : 0;
      const float* src = Bm + (size_t)(r < nb ? r : 0) * ldb + (gred < nred ? gred : 0);
      cp_async16(b_s + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4), src, valid);
    }
  };

  // prologue
  for (int s = 0; s < TC_STAGES - 1; ++s) { if (s < nstage) load_stage(s, s); cp_async_commit(); }
  uint32_t free_parity[TC_STAGES] = {0, 0, 0};
  for (int kb = 0; kb < nstage; ++kb) {
    int s = kb % TC_STAGES;
    // prefetch block kb + STAGES-1 into the stage used by block kb-1 (wait until its MMAs have drained)
    int kn = kb + TC_STAGES - 1;
    if (kn < nstage) {
      int sn = kn % TC_STAGES;
      if (kn >= TC_STAGES) { mbar_wait(&bar_free[sn], free_parity[sn]); free_parity[sn] ^= 1; }
      load_stage(kn, sn);
    }
    cp_async_commit();
    cp_async_wait<TC_STAGES - 1>();            // block kb has landed (for this thread's copies)
    __syncthreads();
    // split: x_lo = x - tf32_trunc(x) into the second buffer (same swizzled positions); raw buffer serves as x_hi
    uint8_t* sp = stage_ptr(s);
    {
      float4* a_raw = (float4*)sp; float4* a_lo = (float4*)(sp + TC_A_BYTES);
      for (int i = tid; i < TC_A_BYTES / 16; i += 128) {
        float4 x = a_raw[i], l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        a_lo[i] = l;
      }
      float4* b_raw = (float4*)(sp + 2 * TC_A_BYTES); float4* b_lo = (float4*)(sp + 2 * TC_A_BYTES + TC_B_BYTES);
      for (int i = tid; i < TC_B_BYTES / 16; i += 128) {
        float4 x = b_raw[i], l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        b_lo[i] = l;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t a_s = smem_u32(sp), al_s = a_s + TC_A_BYTES, b_s = a_s + 2 * TC_A_BYTES, bl_s = b_s + TC_B_BYTES;
#pragma unroll
      for (int k = 0; k < TC_KB / 8; ++k) {                          // 4 UMMA K-steps of 8 tf32 (32 bytes) inside the swizzle span
        uint32_t ko = k * 32;
        umma_tf32(tmem_d, umma_desc(a_s + ko), umma_desc(b_s + ko), idesc, (kb | k) != 0);
        umma_tf32(tmem_d, umma_desc(al_s + ko), umma_desc(b_s + ko), idesc, 1);
        umma_tf32(tmem_d, umma_desc(a_s + ko), umma_desc(bl_s + ko), idesc, 1);
      }
      umma_commit(&bar_free[s]);                                      // arrives when these MMAs (and all earlier ones) completed
      if (kb == nstage - 1) umma_commit(&bar_done);
    }
  }
  // epilogue: accumulator TMEM -> registers (lane = row within the warp's 32-row slab, 16 columns = frames)
  mbar_wait(&bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                 "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  int row = row0 + warp * 32 + lane;
  if (row < rows) {
    float base = 0.f;
    if (mode == 0) base = (add0 ? add0[row] : 0.f) + (add1 ? add1[row] : 0.f);
#pragma unroll
    for (int n = 0; n < TC_N; ++n) {
      if (n >= nb) break;
      float d = __uint_as_float(v[n]);
      if (mode == 0) out[(size_t)n * ldo + row] = base + d;
      else atomicAdd(out + (size_t)n * ldo + row, d);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32) : "memory");
}

static bool tc_attr_set = false;

// forward: v_shaped[b][m] for b in [0,B): A = S_bwd [3V][K], Bm = betas [B][K]
void launch_blend_tc_fwd(vhap_ctx* c, const float* offset, int B, cudaStream_t s) {
  if (!tc_attr_set) { cudaFuncSetAttribute(k_blend_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES); tc_attr_set = true; }
  int M = 3 * c->V, K = c->K;
  for (int b0 = 0; b0 < B; b0 += TC_N) {
    int nb = B - b0 < TC_N ? B - b0 : TC_N;
    dim3 g((M + TC_ROWS - 1) / TC_ROWS, 1);
    LAUNCH(c, KID_BLEND_FWD, s, vh_launch(k_blend_tc, g, 128, TC_SMEM_BYTES, s, c->S_bwd, K, M, c->betas + (size_t)b0 * K, K, nb, K, K, 0, c->v_template, offset,
                                                                         c->v_shaped + (size_t)b0 * M, M));
  }
}

// backward: gbetas[b][k] += sum_m S[m][k] g[b][m]: A = S_fwd_pad [K][Mpad] (rows k, reduction m contiguous), Bm = g_vshaped_pad [B][Mpad]
void launch_blend_tc_bwd(vhap_ctx* c, int B, cudaStream_t s) {
  if (!tc_attr_set) { cudaFuncSetAttribute(k_blend_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES); tc_attr_set = true; }
  int M = 3 * c->V, K = c->K, Mp = c->Mpad;
  const int red_per_cta = 512;                       // 16 stages per CTA
  for (int b0 = 0; b0 < B; b0 += TC_N) {
    int nb = B - b0 < TC_N ? B - b0 : TC_N;
    dim3 g((K + TC_ROWS - 1) / TC_ROWS, (M + red_per_cta - 1) / red_per_cta);
    LAUNCH(c, KID_BLEND_BWD, s, vh_launch(k_blend_tc, g, 128, TC_SMEM_BYTES, s, c->S_fwd_pad, Mp, K, c->g_vshaped + (size_t)b0 * Mp, Mp, nb, M, red_per_cta, 1, nullptr,
                                                                         nullptr, c->gbetas + (size_t)b0 * K, K));
  }
}
