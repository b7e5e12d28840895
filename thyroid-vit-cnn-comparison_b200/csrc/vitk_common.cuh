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

// ------------------------------------------------------------------ dropout masks (nn.Dropout of the reference, drop_rate > 0)
// Counter-based: the keep / drop decision of element (row, col) of dropout site `site` is a pure function of
// (seed, site, row * ld + col), so no mask is ever stored -- forward epilogues and the backward kernels that need the
// same mask simply recompute it.  One Philox4x32-7 call yields 8 x 16-bit uniforms for 8 consecutive elements
// (ld % 8 == 0); an element is dropped when its uniform < thresh = round(p * 65536), kept values are scaled by 1/(1-p).
struct DropSpec {
  const unsigned long long* seed;  // device scalar drawn per step (nullptr = dropout off)
  uint32_t thresh;
  float inv_keep;
  uint32_t site;
};
__host__ inline DropSpec make_drop_spec(const void* seed, float p, int site) {
  DropSpec d;
  d.seed = (seed != nullptr && p > 0.f) ? static_cast<const unsigned long long*>(seed) : nullptr;
  d.thresh = uint32_t(p * 65536.f + 0.5f);
  d.inv_keep = 1.f / (1.f - p);
  d.site = uint32_t(site);
  return d;
}
__device__ __forceinline__ uint4 philox4x32_7(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// uniforms of elements [8 * blk, 8 * blk + 8): element j sits in 16-bit half (j & 1) of word (j >> 1)
__device__ __forceinline__ uint4 drop_bits8(unsigned long long seed, uint32_t site, unsigned long long blk) {
  return philox4x32_7(make_uint4(uint32_t(blk), uint32_t(blk >> 32), site, 0x5eedu), make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
}
__device__ __forceinline__ float2 drop_pair(uint32_t word, uint32_t thresh, float inv_keep) {
  return make_float2((word & 0xffffu) >= thresh ? inv_keep : 0.f, (word >> 16) >= thresh ? inv_keep : 0.f);
}
// factors of the 4 elements [4 * q4, 4 * q4 + 4) (q4 = element index / 4) -- for kernels that own one float4 per thread
__device__ __forceinline__ float4 drop_factors4(const DropSpec& d, unsigned long long seed, unsigned long long q4) {
  const uint4 b = drop_bits8(seed, d.site, q4 >> 1);
  const float2 lo = drop_pair((q4 & 1) ? b.z : b.x, d.thresh, d.inv_keep), hi = drop_pair((q4 & 1) ? b.w : b.y, d.thresh, d.inv_keep);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
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
