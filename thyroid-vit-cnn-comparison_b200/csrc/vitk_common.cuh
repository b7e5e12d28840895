// vitk_common.cuh -- shared device/host helpers for libvitk (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vitk.h"

namespace vitk {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define VITK_CHECK_ARG(cond, ...)              \
  do {                                         \
    if (!(cond)) {                             \
      ::vitk::set_error(__VA_ARGS__);          \
      return VITK_ERR_INVALID;                 \
    }                                          \
  } while (0)

#define VITK_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ::vitk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                         \
      return VITK_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define VITK_LAUNCH_CHECK()                                                                 \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) {                                                               \
      ::vitk::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),        \
                        __FILE__, __LINE__);                                                \
      return VITK_ERR_CUDA;                                                                 \
    }                                                                                       \
    ::vitk::count_launch();                                                                 \
  } while (0)

inline int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

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

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 16-bit pair in the runtime-selected element type (fp16 or bf16)
__device__ __forceinline__ uint32_t pack16(float lo, float hi, bool fp16) { return fp16 ? pack_f16(lo, hi) : pack_bf16(lo, hi); }
__device__ __forceinline__ float2 unpack16(uint32_t u, bool fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return unpack_bf16(u);
}

// erf with |abs error| <= ~3e-7 (Abramowitz & Stegun 7.1.26 in fp32; exp and reciprocal run on the SFU pipe).
// The CUDA libm erff costs ~3x more instructions; at 4 x D GELU evaluations per token the exact-erf epilogue would
// otherwise be ISSUE-bound rather than HBM-bound.  Returns erf(x) and e = exp(-x*x) (reused by the derivative).
__device__ __forceinline__ float erf_fast(float x, float& e) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  e = __expf(-ax * ax);
  return copysignf(fmaf(-poly * t, e, 1.0f), x);
}
// exact (erf) GELU, as nn.GELU() default (vision_transformer_base.py:212-219)
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752f, e));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;  // = exp(-x*x/2)
  const float cdf = 0.5f * (1.0f + erf_fast(x * 0.70710678118654752f, e));
  return fmaf(x * 0.3989422804014327f, e, cdf);
}

// streaming 128-bit global accesses
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ uint4 ldg_u4(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ uint2 ldg_u2(const void* p) {
  return __ldg(reinterpret_cast<const uint2*>(p));
}

#endif  // __CUDACC__

}  // namespace vitk
