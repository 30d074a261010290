// Shared helpers.  Functions marked VH_HD compile for the device (nvcc) and, for the math unit checks under
// tests/hostcheck (g++, -DVH_HOST_CHECK), for the host.  The host build is a developer aid for verifying the
// analytic derivatives against the oracle without a GPU; it is never linked into libvhap_b200.so.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define VH_HD __host__ __device__ __forceinline__
#define VH_D __device__ __forceinline__
#define VH_UNROLL _Pragma("unroll")
#else
#define VH_HD inline
#define VH_D inline
#define VH_UNROLL
#endif

#include <string.h>
VH_HD float vh_i2f(int i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float f; memcpy(&f, &i, 4); return f;
#endif
}
VH_HD int vh_f2i(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int i; memcpy(&i, &f, 4); return i;
#endif
}

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
struct i4 { int x, y, z, w; };

VH_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
VH_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
VH_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
VH_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
VH_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
VH_HD float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
VH_HD f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// Atomic accumulation to distinct-address arrays (vertex / texel gradients).
#if defined(__CUDA_ARCH__)
#define VH_ATOMIC_ADD(ptr, val) atomicAdd((ptr), (val))
// one 16-byte vector reduction (red.global.add.v4.f32, sm_90+) instead of 3-4 scalar ones; ptr must be 16-byte aligned
#define VH_ATOMIC_ADD4(ptr, x, y, z, w) atomicAdd(reinterpret_cast<float4*>(ptr), make_float4((x), (y), (z), (w)))
#else
#define VH_ATOMIC_ADD(ptr, val) (*(ptr) += (val))
#define VH_ATOMIC_ADD4(ptr, x, y, z, w) do { (ptr)[0] += (x); (ptr)[1] += (y); (ptr)[2] += (z); (ptr)[3] += (w); } while (0)
#endif

#ifndef VH_HOST_CHECK
#include <cuda_runtime.h>
#define VH_CUDA_OK(expr)                                                          \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) { vh_set_error(ctx, #expr, cudaGetErrorString(_e)); return -2; } \
  } while (0)
#endif

// SH constants (render_nvdiffrast.py:83-96)
#define VH_SH_C0 0.28209479177387814f      /* 1/sqrt(4pi) */
#define VH_SH_C1 1.0233267079464885f       /* (2pi/3) sqrt(3/(4pi)) */
#define VH_SH_C2 0.8580855308097834f       /* (pi/4) 3 sqrt(5/(12pi)) */
#define VH_SH_C3 0.4290427654048917f       /* (pi/4)(3/2) sqrt(5/(12pi)) */
#define VH_SH_C4 0.24770795610037571f      /* (pi/4)(1/2) sqrt(5/(4pi)) */
