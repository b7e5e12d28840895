// pool_head.cu -- the general classification tail of VisionTransformerBase.forward_features / forward
// (vision_transformer_base.py:468-486): final LayerNorm, pooling over a token range (pool_type 'gap': mean of x[:, 1:], or of
// every token when there is no class token; 'cls' with a representation layer: the single row 0), the optional
// `pre_logits` = Linear + Tanh (:380-386) and the head.  The default configuration (cls pooling, no pre_logits, DeiT's two
// heads) keeps its fused kernels in elementwise.cu; this file serves the remaining constructor options.
// All of it is a few hundred KB of fp32 work per step: one CTA per image, fp32 throughout, no tensor cores.
#include "vitk_common.cuh"

namespace vitk {
namespace {

constexpr int POOL_WARPS = 8;

// pooled[b, :] = mean over rows t in [t0, t1) of LayerNorm(x[b, t, :]); mean / rstd of those rows are saved for backward.
__global__ void __launch_bounds__(POOL_WARPS * 32)
    pool_norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ pooled, float* __restrict__ mean, float* __restrict__ rstd, int T, int dim, int t0,
                         int t1, float eps) {
  extern __shared__ float acc[];   // [POOL_WARPS][dim]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my = acc + warp * dim;
  for (int i = lane; i < dim; i += 32) my[i] = 0.f;
  for (int t = t0 + warp; t < t1; t += POOL_WARPS) {
    const float* xr = x + ((long long)b * T + t) * dim;
    float s = 0.f;
    for (int i = lane; i < dim; i += 32) s += xr[i];
    const float mu = warp_sum(s) / float(dim);
    float q = 0.f;
    for (int i = lane; i < dim; i += 32) {
      const float d = xr[i] - mu;
      q += d * d;
    }
    const float rs = rsqrtf(warp_sum(q) / float(dim) + eps);
    if (lane == 0) {
      mean[(long long)b * T + t] = mu;
      rstd[(long long)b * T + t] = rs;
    }
    for (int i = lane; i < dim; i += 32) my[i] += (xr[i] - mu) * rs * gamma[i] + beta[i];
  }
  __syncthreads();
  const float inv_n = 1.f / float(t1 - t0);
  for (int i = threadIdx.x; i < dim; i += POOL_WARPS * 32) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < POOL_WARPS; ++w) v += acc[w * dim + i];
    pooled[(long long)b * dim + i] = v * inv_n;
  }
}

// backward of the above for a TRUE dpooled [B, dim]: dx (fp32, x S) and dx16 (x S x branch factor x dropout mask) for every
// row of the image (zero outside [t0, t1)), dgamma / dbeta / dcolsum accumulated with true values.
__global__ void __launch_bounds__(POOL_WARPS * 32)
    pool_norm_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ x, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma, float* __restrict__ dx,
                         __nv_bfloat16* __restrict__ dx16, int fp16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         float* __restrict__ dcolsum, const float* __restrict__ loss_scale,
                         const float* __restrict__ branch_scale, DropSpec drop, int T, int dim, int t0, int t1) {
  extern __shared__ float acc[];   // [2][POOL_WARPS][dim]: dgamma partials, dcolsum partials
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float S = loss_scale != nullptr ? __ldg(loss_scale) : 1.f;
  const unsigned long long dseed = drop.seed != nullptr ? __ldg(drop.seed) : 0ull;
  float* my_g = acc + warp * dim;
  float* my_c = acc + (POOL_WARPS + warp) * dim;
  for (int i = lane; i < dim; i += 32) {
    my_g[i] = 0.f;
    my_c[i] = 0.f;
  }
  const float inv_n = 1.f / float(t1 - t0);
  const float* dp = dpooled + (long long)b * dim;
  for (int t = warp; t < T; t += POOL_WARPS) {
    const long long row = (long long)b * T + t;
    float* dxr = dx + row * dim;
    if (t < t0 || t >= t1) {   // rows that do not feed the pooled feature
      for (int i = lane; i < dim; i += 32) {
        dxr[i] = 0.f;
        if (dx16 != nullptr) dx16[row * dim + i] = __float2bfloat16(0.f);   // +0 has the same bits in fp16 and bf16
      }
      continue;
    }
    const float* xr = x + row * dim;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < dim; i += 32) {
      const float g = dp[i] * inv_n * gamma[i];
      const float xh = (xr[i] - mu) * rs;
      s1 += g;
      s2 += g * xh;
      my_g[i] += dp[i] * inv_n * xh;
    }
    const float m1 = warp_sum(s1) / float(dim), m2 = warp_sum(s2) / float(dim);
    const float bs = branch_scale != nullptr ? __ldg(branch_scale + row) : 1.f;
    for (int i = lane; i < dim; i += 32) {
      const float g = dp[i] * inv_n * gamma[i];
      const float xh = (xr[i] - mu) * rs;
      const float o = rs * (g - m1 - xh * m2);
      dxr[i] = o * S;
      float ob = o * bs;
      if (drop.seed != nullptr) {
        const unsigned long long e = (unsigned long long)row * dim + i;
        const uint4 bits = drop_bits8(dseed, drop.site, e >> 3);
        const int j = int(e & 7);
        const uint32_t w = j < 2 ? bits.x : j < 4 ? bits.y : j < 6 ? bits.z : bits.w;
        ob *= (((j & 1) ? (w >> 16) : (w & 0xffffu)) >= drop.thresh) ? drop.inv_keep : 0.f;
      }
      if (dx16 != nullptr) {
        if (fp16) reinterpret_cast<__half*>(dx16)[row * dim + i] = __float2half_rn(ob * S);
        else dx16[row * dim + i] = __float2bfloat16(ob * S);
      }
      my_c[i] += ob;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < dim; i += POOL_WARPS * 32) {
    float g = 0.f, c = 0.f;
#pragma unroll
    for (int w = 0; w < POOL_WARPS; ++w) {
      g += acc[w * dim + i];
      c += acc[(POOL_WARPS + w) * dim + i];
    }
    atomicAdd(dgamma + i, g);
    atomicAdd(dbeta + i, dp[i]);   // sum over the pooled rows of dpooled / n
    if (dcolsum != nullptr) atomicAdd(dcolsum + i, c);
  }
}

// y[b, c] = act(sum_i x[b, i] * W[c, i] + bias[c]); one warp per output element.  act: 0 identity, 1 tanh
__global__ void dense_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                 float* __restrict__ y, int B, int in_dim, int out_dim, int act) {
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * out_dim) return;
  const int b = int(w / out_dim), c = int(w % out_dim);
  float s = 0.f;
  for (int i = lane; i < in_dim; i += 32) s += x[(long long)b * in_dim + i] * W[(long long)c * in_dim + i];
  s = warp_sum(s);
  if (lane == 0) {
    s += bias != nullptr ? bias[c] : 0.f;
    y[(long long)b * out_dim + c] = act == 1 ? tanhf(s) : s;
  }
}
// dz = dy * act'(y)
__global__ void dense_dz_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz, long long n,
                                int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dz[i] = act == 1 ? dy[i] * (1.f - y[i] * y[i]) : dy[i];
}
// dx[b, i] = sum_c dz[b, c] * W[c, i]
__global__ void dense_dx_kernel(const float* __restrict__ dz, const float* __restrict__ W, float* __restrict__ dx, int B,
                                int in_dim, int out_dim) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)B * in_dim;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = int(idx / in_dim), i = int(idx % in_dim);
    float s = 0.f;
    for (int c = 0; c < out_dim; ++c) s += dz[(long long)b * out_dim + c] * W[(long long)c * in_dim + i];
    dx[idx] = s;
  }
}
// dW[c, i] += sum_b dz[b, c] * x[b, i];  db[c] += sum_b dz[b, c]  (one thread owns each output: no atomics)
__global__ void dense_dw_kernel(const float* __restrict__ dz, const float* __restrict__ x, float* __restrict__ dW,
                                float* __restrict__ db, int B, int in_dim, int out_dim) {
  const long long nW = (long long)out_dim * in_dim;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nW + out_dim;
       idx += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    if (idx < nW) {
      const int c = int(idx / in_dim), i = int(idx % in_dim);
      for (int b = 0; b < B; ++b) s += dz[(long long)b * out_dim + c] * x[(long long)b * in_dim + i];
      dW[idx] += s;
    } else if (db != nullptr) {
      const int c = int(idx - nW);
      for (int b = 0; b < B; ++b) s += dz[(long long)b * out_dim + c];
      db[c] += s;
    }
  }
}

inline int grid_for(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_pool_norm_fwd(const float* x, const float* gamma, const float* beta, float* pooled, float* mean, float* rstd,
                                  int32_t B, int32_t T, int32_t dim, int32_t t0, int32_t t1, float eps, void* stream) {
  VITK_CHECK_ARG(x && gamma && beta && pooled && mean && rstd, "vitk_pool_norm_fwd: null pointer");
  VITK_CHECK_ARG(B > 0 && dim > 0 && dim <= 2048 && 0 <= t0 && t0 < t1 && t1 <= T, "vitk_pool_norm_fwd: bad shape / token range");
  const size_t smem = (size_t)POOL_WARPS * dim * sizeof(float);
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(pool_norm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POOL_WARPS * 2048 * 4));
    configured = true;
  }
  pool_norm_fwd_kernel<<<B, POOL_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, gamma, beta, pooled, mean, rstd, T, dim,
                                                                                           t0, t1, eps);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_pool_norm_bwd(const float* dpooled, const float* x, const float* mean, const float* rstd, const float* gamma,
                                  float* dx, void* dx16, int32_t dx16_dtype, float* dgamma, float* dbeta, float* dcolsum,
                                  const float* loss_scale, const float* branch_scale, const vitk_dropout* branch_drop, int32_t B,
                                  int32_t T, int32_t dim, int32_t t0, int32_t t1, void* stream) {
  VITK_CHECK_ARG(dpooled && x && mean && rstd && gamma && dx && dgamma && dbeta, "vitk_pool_norm_bwd: null pointer");
  VITK_CHECK_ARG(B > 0 && dim > 0 && dim <= 2048 && 0 <= t0 && t0 < t1 && t1 <= T, "vitk_pool_norm_bwd: bad shape / token range");
  const DropSpec ds = branch_drop != nullptr ? make_drop_spec(branch_drop->seed, branch_drop->p, branch_drop->site)
                                             : make_drop_spec(nullptr, 0.f, 0);
  VITK_CHECK_ARG(ds.seed == nullptr || (branch_drop->p < 1.f && dim % 8 == 0), "vitk_pool_norm_bwd: dropout needs p < 1, dim %% 8 == 0");
  const size_t smem = (size_t)2 * POOL_WARPS * dim * sizeof(float);
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(pool_norm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * POOL_WARPS * 2048 * 4));
    configured = true;
  }
  pool_norm_bwd_kernel<<<B, POOL_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      dpooled, x, mean, rstd, gamma, dx, reinterpret_cast<__nv_bfloat16*>(dx16), int(dx16_dtype == VITK_FP16), dgamma, dbeta, dcolsum,
      loss_scale, branch_scale, ds, T, dim, t0, t1);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_dense_fwd(const float* x, const float* W, const float* bias, float* y, int32_t B, int32_t in_dim,
                              int32_t out_dim, int32_t act, void* stream) {
  VITK_CHECK_ARG(x && W && y && B > 0 && in_dim > 0 && out_dim > 0 && (act == 0 || act == 1), "vitk_dense_fwd: bad args");
  const long long warps = (long long)B * out_dim;
  dense_fwd_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, W, bias, y, B, in_dim, out_dim,
                                                                                                  act);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_dense_bwd(const float* dy, const float* y, const float* x, const float* W, float* dz, float* dx, float* dW,
                              float* db, int32_t B, int32_t in_dim, int32_t out_dim, int32_t act, void* stream) {
  VITK_CHECK_ARG(dy && x && W && dz && dW && B > 0 && in_dim > 0 && out_dim > 0 && (act == 0 || (act == 1 && y)),
                 "vitk_dense_bwd: bad args (dz scratch [B, out_dim] is required; y is required for tanh)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)B * out_dim;
  dense_dz_kernel<<<grid_for(n, 256), 256, 0, st>>>(dy, y, dz, n, act);
  VITK_LAUNCH_CHECK();
  if (dx != nullptr) {
    dense_dx_kernel<<<grid_for((long long)B * in_dim, 256), 256, 0, st>>>(dz, W, dx, B, in_dim, out_dim);
    VITK_LAUNCH_CHECK();
  }
  dense_dw_kernel<<<grid_for((long long)out_dim * in_dim + out_dim, 128), 128, 0, st>>>(dz, x, dW, db, B, in_dim, out_dim);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
