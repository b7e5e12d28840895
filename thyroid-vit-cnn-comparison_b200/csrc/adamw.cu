// adamw.cu -- multi-tensor AdamW over one flat parameter buffer, fused with Lightning's global
// L2-norm gradient clipping, with the fp32 -> 16-bit weight shadows the tensor-core GEMMs read, and with
// the dynamic loss-scale bookkeeping of the fp16 gradient path.
//
// Restates torch.optim.AdamW(param_groups, lr, betas, eps=1e-8, weight_decay) as configured at
// src/training/lightning_modules.py:599-604 (ViT) and :1108-1113 (distillation: one group per
// parameter with lr = base_lr*lr_scale, :1101-1103) and clip_grad_norm_(max_norm) from
// configs/trainer/default.yaml:21,54:
//     coef = min(1, max_norm / (||g||_2 + 1e-6));  g <- g*coef
//     p <- p*(1 - lr*wd);  m <- b1*m + (1-b1)*g;  v <- b2*v + (1-b2)*g^2
//     p <- p - (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
//
// Loss scaling (torch.cuda.amp.GradScaler semantics, all on the device so the step replays from a CUDA
// graph): activation gradients travel as fp16 multiplied by S; parameter gradients were already divided by
// S by the kernels that produced them.  If ||g||^2 is not finite the step is skipped and S halves; after
// `growth_interval` consecutive clean steps S doubles.
//
// HBM roofline: 28 B/param (read p,g,m,v = 16 B; write p,m,v = 12 B) + 2 B per 16-bit shadow; the norm
// pass re-reads g (4 B/param).  state = {step, lr, grad_sqnorm, clip_coef}.
#include "vitk_common.cuh"

namespace vitk {
namespace {

// ||g||^2 with a FIXED summation order: every block writes its partial sum, the last block to finish adds the partials in
// index order.  (A float atomicAdd per block is shorter, but its order differs from launch to launch; the clip coefficient
// -- and with it every updated parameter -- would then differ in the last bits between data-parallel ranks that hold
// bit-identical all-reduced gradients, and the replicas would drift apart.)  scratch: uint counter at [0] (its place must not
// depend on the grid size: the buffer is shared by launches of different sizes), float partials from [1].
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ state,
                                                     float* __restrict__ scratch) {
  __shared__ float red[8];
  __shared__ bool is_last;
  float acc = 0.f;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ldg_f4(g + 4 * i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += g[i] * g[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch);
  float* partial = scratch + 1;
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w];
    partial[blockIdx.x] = v;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) t += __ldcg(partial + i);   // thread-strided, then a fixed tree
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w];
    state[2] += v;          // the accumulator is cleared by the optimizer's tick kernel
    *counter = 0u;          // ready for the next launch
  }
}

__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) <= 3.402823466e+38f; }  // false for inf and nan

__global__ void __launch_bounds__(256)
    adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 __nv_bfloat16* __restrict__ p16, __half* __restrict__ ph16, const long long* __restrict__ chunk_off,
                 const int* __restrict__ chunk_len, const float* __restrict__ chunk_lr_scale,
                 const float* __restrict__ chunk_wd, const float* __restrict__ state, float beta1, float beta2, float eps,
                 float max_norm) {
  if (!finite_f(state[2])) return;  // overflowed step: skip (every CTA sees the same value)
  const int c = blockIdx.x;
  const long long off = chunk_off[c];
  const int len = chunk_len[c];
  const float t = state[0] + 1.f;
  const float lr = state[1] * chunk_lr_scale[c];
  const float wd = chunk_wd[c];
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(state[2]) + 1e-6f));
  const float bc1 = 1.f - powf(beta1, t);
  const float bc2 = 1.f - powf(beta2, t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.f - lr * wd;
  const int len4 = len >> 2;
  for (int i = threadIdx.x; i < len4; i += blockDim.x) {
    const long long e = off + 4 * (long long)i;
    float4 pv = *reinterpret_cast<const float4*>(p + e);
    float4 gv = ldg_f4(g + e);
    float4 mv = *reinterpret_cast<const float4*>(m + e);
    float4 vv = *reinterpret_cast<const float4*>(v + e);
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k] * coef;
      mp[k] = beta1 * mp[k] + (1.f - beta1) * gk;
      vp[k] = beta2 * vp[k] + (1.f - beta2) * gk * gk;
      const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
      pp[k] = pp[k] * decay - step_size * (mp[k] / denom);
    }
    *reinterpret_cast<float4*>(p + e) = pv;
    *reinterpret_cast<float4*>(m + e) = mv;
    *reinterpret_cast<float4*>(v + e) = vv;
    if (p16 != nullptr) *reinterpret_cast<uint2*>(p16 + e) = make_uint2(pack_bf16(pv.x, pv.y), pack_bf16(pv.z, pv.w));
    if (ph16 != nullptr) *reinterpret_cast<uint2*>(ph16 + e) = make_uint2(pack_f16(pv.x, pv.y), pack_f16(pv.z, pv.w));
  }
  for (int i = (len4 << 2) + threadIdx.x; i < len; i += blockDim.x) {
    const long long e = off + i;
    const float gk = g[e] * coef;
    const float mk = beta1 * m[e] + (1.f - beta1) * gk;
    const float vk = beta2 * v[e] + (1.f - beta2) * gk * gk;
    const float denom = sqrtf(vk) * inv_sqrt_bc2 + eps;
    const float pk = p[e] * decay - step_size * (mk / denom);
    p[e] = pk; m[e] = mk; v[e] = vk;
    if (p16 != nullptr) p16[e] = __float2bfloat16(pk);
    if (ph16 != nullptr) ph16[e] = __float2half_rn(pk);
  }
}

__device__ __forceinline__ void amp_bookkeeping(float* amp, bool overflow, int growth_interval) {
  if (amp == nullptr) return;
  if (overflow) {
    amp[0] = fmaxf(amp[0] * 0.5f, 1.f);
    amp[2] = 0.f;
    amp[3] += 1.f;
    amp[4] = 1.f;
  } else {
    amp[2] += 1.f;
    amp[4] = 0.f;
    if (growth_interval > 0 && amp[2] >= float(growth_interval)) {
      amp[0] = fminf(amp[0] * 2.f, 16777216.f);
      amp[2] = 0.f;
    }
  }
  amp[1] = 1.f / amp[0];
}

// after the update: step += 1 (clean steps only), publish the clip coefficient, clear the norm accumulator
__global__ void adamw_tick_kernel(float* state, float* amp, float max_norm, int growth_interval) {
  const bool ok = finite_f(state[2]);
  state[3] = (ok && max_norm > 0.f) ? fminf(1.f, max_norm / (sqrtf(state[2]) + 1e-6f)) : 1.f;
  if (ok) state[0] += 1.f;
  state[2] = 0.f;
  amp_bookkeeping(amp, !ok, growth_interval);
}

// compat path (external torch optimizer): zero the gradients of an overflowed step
__global__ void amp_zero_if_overflow_kernel(float* __restrict__ g, long long n, const float* __restrict__ scratch) {
  if (finite_f(scratch[2])) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] = 0.f;
}
__global__ void amp_tick_kernel(float* scratch, float* amp, int growth_interval) {
  amp_bookkeeping(amp, !finite_f(scratch[2]), growth_interval);
  scratch[2] = 0.f;
}

int sq_grid(long long n) {
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_sqnorm_scratch_floats(void) { return num_sms() * 8 + 4; }

extern "C" int vitk_grad_sqnorm(const float* grads, int64_t n, float* state, float* scratch, void* stream) {
  VITK_CHECK_ARG(grads && state && scratch && n > 0, "vitk_grad_sqnorm: bad args");
  VITK_CHECK_ARG((reinterpret_cast<uintptr_t>(grads) & 15) == 0, "vitk_grad_sqnorm: grads must be 16-byte aligned");
  sqnorm_kernel<<<sq_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(grads, n, state, scratch);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16,
                               void* params_fp16, const int64_t* chunk_off, const int32_t* chunk_len,
                               const float* chunk_lr_scale, const float* chunk_wd, int32_t n_chunks, float* state,
                               float* amp_state, float beta1, float beta2, float eps, float max_grad_norm,
                               int32_t growth_interval, void* stream) {
  VITK_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && chunk_off && chunk_len && chunk_lr_scale && chunk_wd && state,
                 "vitk_adamw_step: null pointer");
  VITK_CHECK_ARG(n_chunks > 0, "vitk_adamw_step: n_chunks must be > 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  adamw_kernel<<<n_chunks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, reinterpret_cast<__nv_bfloat16*>(params_bf16),
                                         reinterpret_cast<__half*>(params_fp16), reinterpret_cast<const long long*>(chunk_off),
                                         chunk_len, chunk_lr_scale, chunk_wd, state, beta1, beta2, eps, max_grad_norm);
  VITK_LAUNCH_CHECK();
  adamw_tick_kernel<<<1, 1, 0, st>>>(state, amp_state, max_grad_norm, growth_interval);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_amp_update(float* grads, int64_t n, float* amp_state, float* scratch4, float* sq_scratch,
                               int32_t growth_interval, void* stream) {
  VITK_CHECK_ARG(grads && amp_state && scratch4 && sq_scratch && n > 0, "vitk_amp_update: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sqnorm_kernel<<<sq_grid(n), 256, 0, st>>>(grads, n, scratch4, sq_scratch);
  VITK_LAUNCH_CHECK();
  amp_zero_if_overflow_kernel<<<sq_grid(n), 256, 0, st>>>(grads, n, scratch4);
  VITK_LAUNCH_CHECK();
  amp_tick_kernel<<<1, 1, 0, st>>>(scratch4, amp_state, growth_interval);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
