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

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Kernels launched through launch_pdl() may begin (prologue: barrier init, TMEM allocation, descriptor prefetch) while the
// previous kernel of the stream is still draining; they call pdl_wait() before touching any global memory, which blocks
// until the previous grid has completed and its writes are visible.  pdl_trigger() lets the NEXT kernel start launching.
// Inside a captured CUDA graph these become programmatic dependency edges.  VITK_PDL=0 falls back to plain launches.
bool pdl_enabled();

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_pdl_cluster(kernel, grid, block, smem, st, 1, static_cast<Args&&>(args)...);
}
#endif

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

// Exact-erf GELU (nn.GELU() default, vision_transformer_base.py:212-219) and its derivative in one pass:
//   gelu(x) = x * Phi(x),  gelu'(x) = Phi(x) + x * phi(x),  Phi(x) = 0.5 * erfc(-x / sqrt2).
// erfc(|u|) = (a1 t + ... + a5 t^5) exp(-u^2), t = 1 / (1 + p |u|)  (Abramowitz & Stegun 7.1.26, |abs error| <= 1.5e-7);
// the reciprocal and the exponential are single MUFU instructions (approx.ftz: no denormal slow paths), the 0.5 is folded
// into the coefficients.  ~17 FP32 instructions + 2 MUFU per element: at 4 x D evaluations per token the libm erff
// (3x the instructions) made the fc1 epilogue issue-bound instead of HBM-bound.
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_erf_both(float x, float& g, float& dg) {
  const float t = rcp_ftz(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f));
  float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  const float e = ex2_ftz(x * x * (-0.5f * 1.4426950408889634f));  // exp(-x^2 / 2)
  const float h = poly * t * e;                                     // 0.5 * erfc(|x| / sqrt2) = Phi(-|x|)
  const float cdf = x >= 0.f ? 1.0f - h : h;
  g = x * cdf;
  dg = fmaf(x * 0.3989422804014327f, e, cdf);
}
// Two elements at a time on the packed fp32 pipe (FFMA2 / FMUL2 / FADD2, sm_100): half the issue slots of the scalar
// version for the polynomial, the products and the final combine; the two MUFUs per element stay scalar.
__device__ __forceinline__ void gelu_erf_both2(float2 x, float2& g, float2& dg) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = __ffma2_rn(make_float2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f), ax,
                                make_float2(1.0f, 1.0f));
  const float2 t = make_float2(rcp_ftz(den.x), rcp_ftz(den.y));
  float2 poly = __ffma2_rn(t, make_float2(0.5f * 1.061405429f, 0.5f * 1.061405429f), make_float2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
  poly = __ffma2_rn(poly, t, make_float2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
  poly = __ffma2_rn(poly, t, make_float2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
  poly = __ffma2_rn(poly, t, make_float2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
  const float2 q = __fmul2_rn(__fmul2_rn(x, x), make_float2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f));
  const float2 e = make_float2(ex2_ftz(q.x), ex2_ftz(q.y));           // exp(-x^2 / 2)
  const float2 h = __fmul2_rn(__fmul2_rn(poly, t), e);                  // Phi(-|x|)
  // Phi(x) = 0.5 + copysign(0.5 - h, x)
  const float2 w = __fadd2_rn(make_float2(0.5f, 0.5f), make_float2(-h.x, -h.y));
  const float2 cdf = __fadd2_rn(make_float2(0.5f, 0.5f), make_float2(copysignf(w.x, x.x), copysignf(w.y, x.y)));
  g = __fmul2_rn(x, cdf);
  dg = __ffma2_rn(__fmul2_rn(x, make_float2(0.3989422804014327f, 0.3989422804014327f)), e, cdf);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float g, dg;
  gelu_erf_both(x, g, dg);
  return g;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float g, dg;
  gelu_erf_both(x, g, dg);
  return dg;
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
