// Shared device/host helpers for libvml_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vml_b200.h"

namespace vml {

using bf16 = __nv_bfloat16;

void set_error(const char* fmt, ...);
void register_kernel(const char* name);
void count_launches(int n);
// Raise (never lower) a kernel's opt-in dynamic shared-memory limit.  A captured CUDA-graph node keeps
// the launch's own size but tools that re-launch graph nodes (ncu) use the CURRENT function attribute,
// so a later, smaller launch of the same kernel must not shrink it.
cudaError_t ensure_dyn_smem(const void* func, size_t bytes);

// Device-side bounds checks of the debug build (python -m vml_b200.build --debug -> libvml_b200_dbg.so, -DVML_DEBUG_BOUNDS):
// compute-sanitizer is not available on the GPU pool, so the indices a kernel derives from device data (cell codes, live
// counts, ring slots, tensor-memory columns, shared-memory boxes) are asserted in place; a violation prints its location and
// traps, which surfaces as a launch failure (VmlError) in the GPU test that runs the forward on this build.
#ifdef VML_DEBUG_BOUNDS
#define VML_DBG_ASSERT(cond)                                                                              \
  do {                                                                                                    \
    if (!(cond)) {                                                                                        \
      printf("VML_DBG_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
             (int)threadIdx.x);                                                                           \
      asm volatile("trap;");                                                                              \
    }                                                                                                     \
  } while (0)
#else
#define VML_DBG_ASSERT(cond) do { } while (0)
#endif

#define VML_CHECK_ARG(cond)                                                        \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      ::vml::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);      \
      return VML_ERR_ARG;                                                          \
    }                                                                              \
  } while (0)

#define VML_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      ::vml::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return VML_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// n kernels were just enqueued by this launcher
#define VML_LAUNCHED(n)          \
  do {                           \
    ::vml::count_launches(n);    \
    VML_CUDA(cudaGetLastError()); \
  } while (0)

// ---- activation element access (float or bf16 storage, fp32 math) -----------------------
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T>
__device__ __forceinline__ T from_f(float x);
template <>
__device__ __forceinline__ float from_f<float>(float x) { return x; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

// load 4 consecutive elements as floats (16B / 8B aligned)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
// 8 consecutive elements
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld8(const float* p) {
  f8 r; float4 a = ld4(p), b = ld4(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ f8 ld8(const bf16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  f8 r; const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
  return r;
}
__device__ __forceinline__ f8 unpack8(const uint4& u) {   // 8 bf16 held in registers -> fp32
  f8 r; const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
  return r;
}
__device__ __forceinline__ void st8(float* p, const f8& r) {
  st4(p, make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
  st4(p + 4, make_float4(r.v[4], r.v[5], r.v[6], r.v[7]));
}
__device__ __forceinline__ void st8(bf16* p, const f8& r) {
  uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ void decode_cell(int code, int& b, int& i, int& j) {
  b = code >> 16; i = (code >> 8) & 0xff; j = code & 0xff;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

}  // namespace vml
