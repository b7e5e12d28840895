// layernorm.cu -- LayerNorm forward / backward, one warp per token row, 128-bit accesses,
// shuffle reductions.  Replaces nn.LayerNorm(dim) (eps 1e-5, affine) at
// vision_transformer_base.py:263,273 (norm1/norm2 of every Block).
//
// The residual stream stays fp32 (x in, dx out); the normalised activations feeding the
// tensor-core GEMMs are emitted as bf16.  Backward also folds in the residual gradient
// (dx = dres + LN'(dy)), emits the bf16 copy the next dgrad/wgrad GEMM consumes and the column
// sum of dx (= bias gradient of the Linear that wrote the residual branch), so the residual
// gradient is read once and written once per LayerNorm.
//
// HBM roofline (algorithmic bytes per row, D = dim):
//   fwd: 4D (x) + 2D (y) + 8 (stats)            bwd: 2D (dy) + 4D (x) + 4D (dres) + 4D (dx) + 2D (dx bf16)
#include "vitk_common.cuh"

namespace vitk {
namespace {

constexpr int LN_WARPS = 8;

template <int CHUNKS>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               __nv_bfloat16* __restrict__ y, int y_fp16,
                                                               float* __restrict__ mean,
                                                               float* __restrict__ rstd, long long rows, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const int nvec = dim >> 2;
  const float inv_dim = 1.f / float(dim);
  float4 gm[CHUNKS], bt[CHUNKS];
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int i = lane + 32 * c;
    gm[c] = i < nvec ? ldg_f4(gamma + 4 * i) : make_float4(0, 0, 0, 0);
    bt[c] = i < nvec ? ldg_f4(beta + 4 * i) : make_float4(0, 0, 0, 0);
  }
  for (long long row = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * LN_WARPS) {
    const float* xr = x + row * dim;
    float4 v[CHUNKS];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = lane + 32 * c;
      v[c] = i < nvec ? ldg_f4(xr + 4 * i) : make_float4(0, 0, 0, 0);
      s += v[c].x + v[c].y + v[c].z + v[c].w;
    }
    const float mu = warp_sum(s) * inv_dim;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = lane + 32 * c;
      if (i < nvec) {
        const float a = v[c].x - mu, b = v[c].y - mu, cc = v[c].z - mu, d = v[c].w - mu;
        q += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rs = rsqrtf(warp_sum(q) * inv_dim + eps);
    __nv_bfloat16* yr = y + row * dim;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = lane + 32 * c;
      if (i < nvec) {
        const float a = (v[c].x - mu) * rs * gm[c].x + bt[c].x;
        const float b = (v[c].y - mu) * rs * gm[c].y + bt[c].y;
        const float cc = (v[c].z - mu) * rs * gm[c].z + bt[c].z;
        const float d = (v[c].w - mu) * rs * gm[c].w + bt[c].w;
        *reinterpret_cast<uint2*>(yr + 4 * i) = make_uint2(pack16(a, b, y_fp16), pack16(cc, d, y_fp16));
      }
    }
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

template <int CHUNKS>
__global__ void __launch_bounds__(LN_WARPS * 32)
    ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_fp16, const float* __restrict__ x, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dres,
                  float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16, int dx_fp16, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, float* __restrict__ dcolsum, const float* __restrict__ unscale, long long rows,
                  int dim) {
  extern __shared__ float red[];  // [3][dim]
  const int lane = threadIdx.x & 31;
  const int nvec = dim >> 2;
  const float inv_dim = 1.f / float(dim);
  for (int i = threadIdx.x; i < 3 * dim; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  float4 gm[CHUNKS], dg[CHUNKS], db[CHUNKS], dc[CHUNKS];
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int i = lane + 32 * c;
    gm[c] = i < nvec ? ldg_f4(gamma + 4 * i) : make_float4(0, 0, 0, 0);
    dg[c] = db[c] = dc[c] = make_float4(0, 0, 0, 0);
  }
  for (long long row = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * LN_WARPS) {
    const float mu = mean[row], rs = rstd[row];
    const float* xr = x + row * dim;
    const __nv_bfloat16* dyr = dy + row * dim;
    float4 xh[CHUNKS], g[CHUNKS];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = lane + 32 * c;
      if (i < nvec) {
        const float4 xv = ldg_f4(xr + 4 * i);
        const uint2 dyu = ldg_u2(dyr + 4 * i);
        const float2 d01 = unpack16(dyu.x, dy_fp16), d23 = unpack16(dyu.y, dy_fp16);
        xh[c] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[c] = make_float4(d01.x * gm[c].x, d01.y * gm[c].y, d23.x * gm[c].z, d23.y * gm[c].w);
        dg[c].x += d01.x * xh[c].x; dg[c].y += d01.y * xh[c].y; dg[c].z += d23.x * xh[c].z; dg[c].w += d23.y * xh[c].w;
        db[c].x += d01.x; db[c].y += d01.y; db[c].z += d23.x; db[c].w += d23.y;
        s1 += g[c].x + g[c].y + g[c].z + g[c].w;
        s2 += g[c].x * xh[c].x + g[c].y * xh[c].y + g[c].z * xh[c].z + g[c].w * xh[c].w;
      } else {
        xh[c] = g[c] = make_float4(0, 0, 0, 0);
      }
    }
    const float m1 = warp_sum(s1) * inv_dim;
    const float m2 = warp_sum(s2) * inv_dim;
    float* dxr = dx + row * dim;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = lane + 32 * c;
      if (i < nvec) {
        float4 o = make_float4(rs * (g[c].x - m1 - xh[c].x * m2), rs * (g[c].y - m1 - xh[c].y * m2),
                               rs * (g[c].z - m1 - xh[c].z * m2), rs * (g[c].w - m1 - xh[c].w * m2));
        if (dres != nullptr) {
          const float4 r = ldg_f4(dres + row * dim + 4 * i);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(dxr + 4 * i) = o;
        if (dx_bf16 != nullptr)
          *reinterpret_cast<uint2*>(dx_bf16 + row * dim + 4 * i) = make_uint2(pack16(o.x, o.y, dx_fp16), pack16(o.z, o.w, dx_fp16));
        dc[c].x += o.x; dc[c].y += o.y; dc[c].z += o.z; dc[c].w += o.w;
      }
    }
  }
  // block reduction of the three column sums, then one atomic per column per block
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int i = lane + 32 * c;
    if (i < nvec) {
      float* r0 = red + 4 * i;
      atomicAdd(r0 + 0, dg[c].x); atomicAdd(r0 + 1, dg[c].y); atomicAdd(r0 + 2, dg[c].z); atomicAdd(r0 + 3, dg[c].w);
      float* r1 = red + dim + 4 * i;
      atomicAdd(r1 + 0, db[c].x); atomicAdd(r1 + 1, db[c].y); atomicAdd(r1 + 2, db[c].z); atomicAdd(r1 + 3, db[c].w);
      if (dcolsum != nullptr) {
        float* r2 = red + 2 * dim + 4 * i;
        atomicAdd(r2 + 0, dc[c].x); atomicAdd(r2 + 1, dc[c].y); atomicAdd(r2 + 2, dc[c].z); atomicAdd(r2 + 3, dc[c].w);
      }
    }
  }
  __syncthreads();
  const float u = unscale != nullptr ? __ldg(unscale) : 1.f;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i] * u);
    atomicAdd(dbeta + i, red[dim + i] * u);
    if (dcolsum != nullptr) atomicAdd(dcolsum + i, red[2 * dim + i] * u);
  }
}

int ln_grid(long long rows) {
  const long long blocks = (rows + LN_WARPS - 1) / LN_WARPS;
  const long long cap = (long long)num_sms() * 8;  // multiple of the SM count, 8 resident CTAs / SM
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace
}  // namespace vitk

using namespace vitk;

#define LN_DISPATCH(CH, ...)            \
  switch (CH) {                         \
    case 1: { constexpr int C_ = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int C_ = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int C_ = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int C_ = 4; __VA_ARGS__; } break; \
    case 6: { constexpr int C_ = 6; __VA_ARGS__; } break; \
    case 8: { constexpr int C_ = 8; __VA_ARGS__; } break; \
    default: set_error("layernorm: dim too large"); return VITK_ERR_UNSUPPORTED; \
  }

static int ln_chunks(int dim) {
  const int c = (dim / 4 + 31) / 32;
  if (c <= 4) return c;
  if (c <= 6) return 6;
  if (c <= 8) return 8;
  return 99;
}

extern "C" int vitk_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int32_t y_dtype,
                                  float* mean, float* rstd, int64_t rows, int32_t dim, float eps, void* stream) {
  VITK_CHECK_ARG(x && gamma && beta && y && mean && rstd, "vitk_layernorm_fwd: null pointer");
  VITK_CHECK_ARG(y_dtype == VITK_BF16 || y_dtype == VITK_FP16, "vitk_layernorm_fwd: y must be bf16 or fp16");
  VITK_CHECK_ARG(rows > 0 && dim > 0 && dim % 4 == 0 && dim <= 1024, "vitk_layernorm_fwd: dim=%d must be a multiple of 4, <= 1024", dim);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ch = ln_chunks(dim);
  LN_DISPATCH(ch, ln_fwd_kernel<C_><<<ln_grid(rows), LN_WARPS * 32, 0, st>>>(
                      x, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), int(y_dtype == VITK_FP16), mean, rstd, rows, dim, eps));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* mean, const float* rstd,
                                  const float* gamma, const float* dres, float* dx, void* dx16, int32_t dx16_dtype,
                                  float* dgamma, float* dbeta, float* dcolsum, const float* grad_unscale, int64_t rows,
                                  int32_t dim, void* stream) {
  VITK_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta, "vitk_layernorm_bwd: null pointer");
  VITK_CHECK_ARG((dy_dtype == VITK_BF16 || dy_dtype == VITK_FP16) && (dx16_dtype == VITK_BF16 || dx16_dtype == VITK_FP16),
                 "vitk_layernorm_bwd: dy / dx16 must be bf16 or fp16");
  VITK_CHECK_ARG(rows > 0 && dim > 0 && dim % 4 == 0 && dim <= 1024, "vitk_layernorm_bwd: dim=%d must be a multiple of 4, <= 1024", dim);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ch = ln_chunks(dim);
  const size_t smem = 3 * (size_t)dim * sizeof(float);
  // fewer, fatter blocks than forward: each block ends with 3*dim global atomics
  long long blocks = (rows + LN_WARPS * 4 - 1) / (LN_WARPS * 4);
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  LN_DISPATCH(ch, ln_bwd_kernel<C_><<<(int)blocks, LN_WARPS * 32, smem, st>>>(
                      reinterpret_cast<const __nv_bfloat16*>(dy), int(dy_dtype == VITK_FP16), x, mean, rstd, gamma, dres, dx,
                      reinterpret_cast<__nv_bfloat16*>(dx16), int(dx16_dtype == VITK_FP16), dgamma, dbeta, dcolsum,
                      grad_unscale, rows, dim));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
