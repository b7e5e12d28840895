// elementwise.cu -- the memory-bound glue of the ViT/DeiT step: casts, patch gather, token
// assembly and its backward, pooled final-norm + heads, bias-gradient column sums, ensemble
// probabilities and attention rollout.  All kernels are coalesced and 128-bit vectorised; grids
// are capped at a multiple of the SM count and grid-stride over the rest.
#include "vitk_common.cuh"

namespace vitk {
namespace {

typedef __nv_bfloat16 bf16;

inline int capped_grid(long long work_items, int threads, int ctas_per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------ fp32 -> bf16
__global__ void cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, __half* __restrict__ dsth, long long n) {
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = ldg_f4(src + 8 * i), b = ldg_f4(src + 8 * i + 4);
    if (dst != nullptr)
      *reinterpret_cast<uint4*>(dst + 8 * i) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    if (dsth != nullptr)
      *reinterpret_cast<uint4*>(dsth + 8 * i) =
          make_uint4(pack_f16(a.x, a.y), pack_f16(a.z, a.w), pack_f16(b.x, b.y), pack_f16(b.z, b.w));
  }
  for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (dst != nullptr) dst[i] = __float2bfloat16(src[i]);
    if (dsth != nullptr) dsth[i] = __float2half_rn(src[i]);
  }
}

// ------------------------------------------------------------------ column sums of a bf16 matrix
// block = 32 column-vectors (8 bf16 each) x 8 row lanes; grid.x = column chunk, grid.y = row slab
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, int fp16, float* __restrict__ out,
                                                     const float* __restrict__ unscale, long long rows, int dim,
                                                     int rows_per_block) {
  __shared__ float red[8][32][8];
  const int cv = blockIdx.x * 32 + threadIdx.x;  // column vector index
  const int col = cv * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  if (col < dim) {
    long long r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {  // four independent 16-byte loads in flight per thread
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = ldg_u4(x + (r + 8 * k) * dim + col);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = unpack16(u[k].x, fp16), b = unpack16(u[k].y, fp16), c = unpack16(u[k].z, fp16), d = unpack16(u[k].w, fp16);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
      }
    }
    for (; r < r1; r += 8) {
      const uint4 u = ldg_u4(x + r * dim + col);
      const float2 a = unpack16(u.x, fp16), b = unpack16(u.y, fp16), c = unpack16(u.z, fp16), d = unpack16(u.w, fp16);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
      acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && col < dim) {
    const float u = unscale != nullptr ? __ldg(unscale) : 1.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x][j];
      atomicAdd(out + col + j, s * u);
    }
  }
}

// ------------------------------------------------------------------ patch gather (fp32 NCHW -> bf16 [B*gh*gw, C*P*P])
__global__ void patchify_kernel(const float* __restrict__ img, bf16* __restrict__ patches, int fp16, int B, int C, int H, int W,
                                int P) {
  const int wv = W >> 3;
  const long long total = (long long)B * C * H * wv;
  const int gw = W / P, gh = H / P;
  const int Kdim = C * P * P;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int xv = int(i % wv);
    long long r = i / wv;
    const int y = int(r % H);
    r /= H;
    const int c = int(r % C);
    const int b = int(r / C);
    const float* src = img + (((long long)b * C + c) * H + y) * W + xv * 8;
    const float4 a0 = ldg_f4(src), a1 = ldg_f4(src + 4);
    const int x = xv * 8;
    const int py = y / P, ky = y - py * P, px = x / P, kx = x - px * P;
    const long long doff = ((long long)(b * gh + py) * gw + px) * Kdim + c * P * P + ky * P + kx;
    *reinterpret_cast<uint4*>(patches + doff) = make_uint4(pack16(a0.x, a0.y, fp16), pack16(a0.z, a0.w, fp16),
                                                           pack16(a1.x, a1.y, fp16), pack16(a1.z, a1.w, fp16));
  }
}

// channel-last patch vectors (projection_type='linear'): one thread writes two adjacent 16-bit outputs k, k+1 of
// k = (ky*P + kx)*C + c -- coalesced 4-byte stores; the loads walk C image planes (L1/L2 absorb the reuse)
__global__ void patchify_hwc_kernel(const float* __restrict__ img, bf16* __restrict__ patches, int fp16, int B, int C, int H, int W,
                                    int P) {
  const int gw = W / P, gh = H / P;
  const int Kdim = C * P * P;          // even: P % 8 == 0
  const int kpairs = Kdim >> 1;
  const long long total = (long long)B * gh * gw * kpairs;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int kp = int(i % kpairs);
    long long r = i / kpairs;
    const int px = int(r % gw);
    r /= gw;
    const int py = int(r % gh);
    const int b = int(r / gh);
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = 2 * kp + e;
      const int c = k % C, pix = k / C;
      const int ky = pix / P, kx = pix - ky * P;
      v[e] = __ldg(img + (((long long)b * C + c) * H + py * P + ky) * W + px * P + kx);
    }
    reinterpret_cast<uint32_t*>(patches)[i] = pack16(v[0], v[1], fp16);
  }
}

// ------------------------------------------------------------------ cls / dist token rows
__global__ void prefix_tokens_kernel(float* __restrict__ x, const float* __restrict__ cls_tok, const float* __restrict__ dist_tok,
                                     const float* __restrict__ pos, int B, int T, int dim, int n_prefix, DropSpec drop) {
  const int dv = dim >> 2;
  const long long total = (long long)B * n_prefix * dv;
  const unsigned long long dseed = drop.seed != nullptr ? __ldg(drop.seed) : 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = int(i % dv);
    const int t = int((i / dv) % n_prefix);
    const int b = int(i / ((long long)dv * n_prefix));
    const float4 tk = ldg_f4((t == 0 ? cls_tok : dist_tok) + 4 * v);
    const float4 ps = ldg_f4(pos + (long long)t * dim + 4 * v);
    float4 o = make_float4(tk.x + ps.x, tk.y + ps.y, tk.z + ps.z, tk.w + ps.w);
    if (drop.seed != nullptr) {   // pos_drop (vision_transformer_base.py:452): same site / indexing as the patch-token epilogue
      const float4 m = drop_factors4(drop, dseed, (((unsigned long long)b * T + t) * dim + 4 * v) >> 2);
      o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
    }
    *reinterpret_cast<float4*>(x + ((long long)b * T + t) * dim + 4 * v) = o;
  }
}

// ------------------------------------------------------------------ backward of token assembly
// grid (ceil(T / TOK_TOKS), ceil(B / TOK_IMGS)); block (dim/4 columns [<= 256], TOK_TOKS tokens): a thread owns one float4
// column of one token and walks its slice of the batch.  The patch-bias gradient (a sum over every patch token and image) is
// first reduced over the block's tokens in shared memory and all accumulations leave as 16-byte vector REDs: the first version
// (one block per token, four scalar atomics per column) sent ~1 500 same-address atomics per bias column to the L2 and ran at
// 0.2 of the HBM rate.
constexpr int TOK_IMGS = 32;
constexpr int TOK_TOKS = 4;
__device__ __forceinline__ void red_add4(float* p, float4 v) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  } else {
    atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
  }
}
__global__ void tokens_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dpos, float* __restrict__ dcls,
                                  float* __restrict__ ddist, bf16* __restrict__ dpatch, int fp16, float* __restrict__ dbias,
                                  const float* __restrict__ unscale, int B, int T, int dim, int n_prefix, DropSpec drop) {
  extern __shared__ float4 s_tok[];   // [TOK_TOKS][blockDim.x]
  const float u = unscale != nullptr ? __ldg(unscale) : 1.f;
  const unsigned long long dseed = drop.seed != nullptr ? __ldg(drop.seed) : 0ull;
  const int t = blockIdx.x * TOK_TOKS + threadIdx.y;
  const int b0 = blockIdx.y * TOK_IMGS;
  const int b1 = min(B, b0 + TOK_IMGS);
  const int rows_per_img = T - n_prefix;
  const int nv = dim >> 2;
  for (int v0 = 0; v0 < nv; v0 += blockDim.x) {   // block-uniform trip count (the loop body synchronises)
    const int v = v0 + threadIdx.x;
    const bool active = v < nv && t < T;
    float4 acc = make_float4(0, 0, 0, 0);
    if (active) {
      for (int bb = b0; bb < b1; bb += 8) {   // 8 independent 16-byte loads in flight per thread
        float4 gs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          gs[k] = (bb + k < b1) ? ldg_f4(dx + ((long long)(bb + k) * T + t) * dim + 4 * v) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int b = bb + k;
          if (b >= b1) break;
          float4 g = gs[k];
          if (drop.seed != nullptr) {   // gradient through pos_drop: the mask the forward applied to this token element
            const float4 m = drop_factors4(drop, dseed, (((unsigned long long)b * T + t) * dim + 4 * v) >> 2);
            g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
          }
          acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
          if (t >= n_prefix && dpatch != nullptr)
            *reinterpret_cast<uint2*>(dpatch + ((long long)b * rows_per_img + (t - n_prefix)) * dim + 4 * v) =
                make_uint2(pack16(g.x, g.y, fp16), pack16(g.z, g.w, fp16));
        }
      }
      acc.x *= u; acc.y *= u; acc.z *= u; acc.w *= u;
      if (dpos != nullptr) red_add4(dpos + (long long)t * dim + 4 * v, acc);
      if (t == 0 && n_prefix >= 1 && dcls != nullptr) red_add4(dcls + 4 * v, acc);
      else if (t == 1 && n_prefix >= 2 && ddist != nullptr) red_add4(ddist + 4 * v, acc);
    }
    // patch-bias gradient: the block's patch tokens are summed here, one vector RED per column and block
    s_tok[threadIdx.y * blockDim.x + threadIdx.x] = (active && t >= n_prefix) ? acc : make_float4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.y == 0 && v < nv && dbias != nullptr) {
      float4 sum = s_tok[threadIdx.x];
#pragma unroll
      for (int k = 1; k < TOK_TOKS; ++k) {
        const float4 o = s_tok[k * blockDim.x + threadIdx.x];
        sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
      }
      if (blockIdx.x * TOK_TOKS + TOK_TOKS > n_prefix) red_add4(dbias + 4 * v, sum);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ leading rows of every image: compact <-> dense
// The head reads only tokens 0..n-1 of the last block's output (cls / dist), so that block's proj / MLP run on B*n rows:
// gather picks those rows out of a [B, T, row] tensor, expand writes them back into a dense tensor whose other rows are zero
// (the gradient the earlier, dense part of the backward continues from).  16-byte vectors; row_v = vectors per row.
__global__ void gather_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long total, int n_row_v,
                                   long long img_v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / n_row_v;
    const int r = (int)(i - b * n_row_v);
    dst[i] = __ldg(src + b * img_v + r);
  }
}
__global__ void expand_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long total, int n_row_v,
                                   long long img_v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / img_v;
    const long long r = i - b * img_v;
    dst[i] = r < n_row_v ? __ldg(src + b * n_row_v + r) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------ pooled final norm + heads
// one warp per (image, head)
__global__ void head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ W0, const float* __restrict__ b0, const float* __restrict__ W1,
                                const float* __restrict__ b1, float* __restrict__ logits0, float* __restrict__ logits1,
                                float* __restrict__ xhat, float* __restrict__ rstd, float* __restrict__ pooled, int B, int T,
                                int dim, int C, int n_heads, float eps) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= B * n_heads) return;
  const int hd = w / B, b = w - hd * B;
  const float* xr = x + ((long long)b * T + hd) * dim;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s += xr[i];
  const float mu = warp_sum(s) / float(dim);
  float q = 0.f;
  for (int i = lane; i < dim; i += 32) {
    const float d = xr[i] - mu;
    q += d * d;
  }
  const float rs = rsqrtf(warp_sum(q) / float(dim) + eps);
  float* xh = xhat + ((long long)hd * B + b) * dim;
  for (int i = lane; i < dim; i += 32) {
    const float v = (xr[i] - mu) * rs;
    xh[i] = v;
    if (pooled != nullptr) pooled[((long long)hd * B + b) * dim + i] = v * gamma[i] + beta[i];   // norm(x)[:, hd] (forward_features)
  }
  if (lane == 0) rstd[hd * B + b] = rs;
  __syncwarp();
  const float* W = hd == 0 ? W0 : W1;
  const float* bias = hd == 0 ? b0 : b1;
  float* lg = (hd == 0 ? logits0 : logits1) + (long long)b * C;
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (int i = lane; i < dim; i += 32) acc += (xh[i] * gamma[i] + beta[i]) * W[(long long)c * dim + i];
    acc = warp_sum(acc);
    if (lane == 0) lg[c] = acc + (bias != nullptr ? bias[c] : 0.f);
  }
}

__global__ void head_bwd_kernel(const float* __restrict__ dl0, const float* __restrict__ dl1, const float* __restrict__ xhat,
                                const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ W0,
                                const float* __restrict__ W1, float* __restrict__ dx, bf16* __restrict__ dx16, int fp16,
                                float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ db0,
                                float* __restrict__ db1, float* __restrict__ dcolsum, const float* __restrict__ loss_scale,
                                const float* __restrict__ branch_scale, DropSpec drop, int B, int T, int dim, int C, int n_heads) {
  const float S = loss_scale != nullptr ? __ldg(loss_scale) : 1.f;
  const unsigned long long dseed = drop.seed != nullptr ? __ldg(drop.seed) : 0ull;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= B * n_heads) return;
  const int hd = w / B, b = w - hd * B;
  const float* dl = (hd == 0 ? dl0 : dl1) + (long long)b * C;
  const float* W = hd == 0 ? W0 : W1;
  float* dbh = hd == 0 ? db0 : db1;
  const float* xh = xhat + ((long long)hd * B + b) * dim;
  const float rs = rstd[hd * B + b];
  // dxn = dl . W ; g = dxn * gamma ; LN backward on the single row
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < dim; i += 32) {
    float dxn = 0.f;
    for (int c = 0; c < C; ++c) dxn += dl[c] * W[(long long)c * dim + i];
    const float g = dxn * gamma[i];
    s1 += g;
    s2 += g * xh[i];
    atomicAdd(dgamma + i, dxn * xh[i]);
    atomicAdd(dbeta + i, dxn);
  }
  const float m1 = warp_sum(s1) / float(dim), m2 = warp_sum(s2) / float(dim);
  float* dxr = dx + ((long long)b * T + hd) * dim;
  const float bs = branch_scale != nullptr ? __ldg(branch_scale + (long long)b * T + hd) : 1.f;
  for (int i = lane; i < dim; i += 32) {
    float dxn = 0.f;
    for (int c = 0; c < C; ++c) dxn += dl[c] * W[(long long)c * dim + i];
    const float g = dxn * gamma[i];
    const float o = rs * (g - m1 - xh[i] * m2);
    dxr[i] = o * S;
    float ob = o * bs;   // gradient entering the last MLP branch (stochastic depth, then that branch's dropout mask)
    if (drop.seed != nullptr) {
      const unsigned long long e = ((unsigned long long)b * T + hd) * dim + i;
      const uint4 bits = drop_bits8(dseed, drop.site, e >> 3);
      const int j = int(e & 7);
      const uint32_t w = j < 2 ? bits.x : j < 4 ? bits.y : j < 6 ? bits.z : bits.w;
      ob *= (((j & 1) ? (w >> 16) : (w & 0xffffu)) >= drop.thresh) ? drop.inv_keep : 0.f;
    }
    if (dx16 != nullptr) {
      if (fp16) reinterpret_cast<__half*>(dx16)[((long long)b * T + hd) * dim + i] = __float2half_rn(ob * S);
      else dx16[((long long)b * T + hd) * dim + i] = __float2bfloat16(ob * S);
    }
    if (dcolsum != nullptr) atomicAdd(dcolsum + i, ob);
  }
  if (dbh != nullptr)
    for (int c = lane; c < C; c += 32) atomicAdd(dbh + c, dl[c]);
}

// dW_h[c, i] += sum_b dl_h[b, c] * (xhat_h[b, i] * gamma[i] + beta[i]); one thread per (h, c, i) and batch slab
// (blockIdx.y): the batch loop is split over 16-sample slabs so the kernel is not one long dependent-load chain
__global__ void head_wgrad_kernel(const float* __restrict__ dl0, const float* __restrict__ dl1, const float* __restrict__ xhat,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ dW0,
                                  float* __restrict__ dW1, int B, int dim, int C, int n_heads) {
  const long long total = (long long)n_heads * C * dim;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = int(idx % dim);
    const int c = int((idx / dim) % C);
    const int hd = int(idx / ((long long)dim * C));
    const float* dl = hd == 0 ? dl0 : dl1;
    const float* xh = xhat + (long long)hd * B * dim;
    const float gm = gamma[i], bt = beta[i];
    float acc = 0.f;
    const int b0 = blockIdx.y * 16, b1 = min(B, b0 + 16);
#pragma unroll 4
    for (int b = b0; b < b1; ++b) acc += __ldg(dl + (long long)b * C + c) * (__ldg(xh + (long long)b * dim + i) * gm + bt);
    float* dW = hd == 0 ? dW0 : dW1;
    atomicAdd(dW + (long long)c * dim + i, acc);
  }
}

// ------------------------------------------------------------------ stochastic depth (DropPath, vision_transformer_base.py:56-64)
// scale[br, b*T + t] = floor(keep + u[br, b]) / keep, keep = 1 - drop_prob[br]: one Bernoulli draw per (branch, sample),
// expanded to the token rows so that GEMM epilogues / LayerNorm backward read one float per row
// factors[r, c] = 0 or 1/(1-p): the mask every fused kernel derives for (seed, site) -- exported for tests / debugging
__global__ void dropout_mask_kernel(float* __restrict__ factors, long long n4, DropSpec drop) {
  const unsigned long long dseed = __ldg(drop.seed);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    *reinterpret_cast<float4*>(factors + 4 * i) = drop_factors4(drop, dseed, (unsigned long long)i);
}

__global__ void droppath_scale_kernel(const float* __restrict__ u, const float* __restrict__ drop_prob, float* __restrict__ scale,
                                      int branches, int B, int T) {
  const long long total = (long long)branches * B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long sb = i / T;                 // branch * B + sample
    const int br = int(sb / B);
    const float keep = 1.f - drop_prob[br];
    scale[i] = floorf(keep + u[sb]) / keep;
  }
}

// ------------------------------------------------------------------ ensemble: sum_f w_f softmax(logits_f) -> argmax
__global__ void ensemble_kernel(const float* __restrict__ logits, const float* __restrict__ weights, float* __restrict__ probs,
                                long long* __restrict__ pred, int F, int B, int C) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* pr = probs + (long long)b * C;
  for (int c = 0; c < C; ++c) pr[c] = 0.f;
  for (int f = 0; f < F; ++f) {
    const float* lg = logits + ((long long)f * B + b) * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, lg[c]);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(lg[c] - mx);
    const float wf = weights[f] / sum;
    for (int c = 0; c < C; ++c) pr[c] += wf * expf(lg[c] - mx);
  }
  int best = 0;
  float bv = pr[0];
  for (int c = 1; c < C; ++c)
    if (pr[c] > bv) {
      bv = pr[c];
      best = c;
    }
  pred[b] = best;
}

// ------------------------------------------------------------------ attention rollout
// fused[b] = fuse_h(probs[l,b,h]); a = (fused + I)/2, rows renormalised; R <- a @ R  (R starts at I)
__global__ void rollout_fuse_kernel(const float* __restrict__ probs, float* __restrict__ a, int l, int B, int H, int N, int fusion) {
  const long long total = (long long)B * N;  // one thread-row per (b, i); threads of a warp stride over j
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= total) return;
  const int i = int(row % N);
  const int b = int(row / N);
  float sum = 0.f;
  float* ar = a + row * N;
  for (int j = lane; j < N; j += 32) {
    float f = fusion == 0 ? 0.f : (fusion == 1 ? -INFINITY : INFINITY);
    for (int h = 0; h < H; ++h) {
      const float p = probs[((((long long)l * B + b) * H + h) * N + i) * N + j];
      f = fusion == 0 ? f + p : (fusion == 1 ? fmaxf(f, p) : fminf(f, p));
    }
    if (fusion == 0) f /= float(H);
    f = 0.5f * (f + (i == j ? 1.f : 0.f));
    ar[j] = f;
    sum += f;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < N; j += 32) ar[j] *= inv;
}
// Rout[b] = a[b] @ Rin[b]   (N x N fp32, N ~ 200: plain tiled loop, eval-only path)
__global__ void rollout_matmul_kernel(const float* __restrict__ a, const float* __restrict__ rin, float* __restrict__ rout, int N) {
  const int b = blockIdx.z;
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const float* ar = a + ((long long)b * N + i) * N;
  const float* r = rin + (long long)b * N * N;
  float acc = 0.f;
  for (int k = 0; k < N; ++k) acc += ar[k] * r[(long long)k * N + j];
  rout[((long long)b * N + i) * N + j] = acc;
}
// One ROW of the rollout without ever forming an N x N product: R = A_L ... A_1, so row r of R is
//   v_L = e_r^T A_L,  v_{l} = v_{l+1} A_l   (l = L-1 .. 1),   A_l[i,:] = 0.5 (fuse_h P_l[i,:] + e_i) / rowsum_i
// i.e. L vector-matrix products per image: the attention maps (the only large operand, L*H*N*N floats per image) are
// read exactly once, 198x fewer flops than the matrix chain.  One CTA per image; warp w owns rows i = w, w+W, ...; lanes
// stride over the columns (coalesced 128-byte row segments); per-warp partial accumulators live in shared memory.
constexpr int ROLL_WARPS = 16;
// NJ = ceil(N / 32) for N <= 256 (row segments held in registers), 0 = any N (two passes over the row); HT = number of heads
// when it is one of the reference's (3 / 6 / 12: the head loop unrolls and all HT * NJ loads of a row are in flight at once --
// with a run-time head count each head's 7 loads waited for the previous head's, ~3 us per row), 0 = run-time H
template <int NJ, int HT>
__global__ void __launch_bounds__(ROLL_WARPS * 32)
    rollout_row_kernel(const float* __restrict__ probs, float* __restrict__ out, long long layer_stride, long long batch_stride, int L,
                       int B, int H, int N, int row, int fusion) {
  extern __shared__ float roll_smem[];
  float* v = roll_smem;                 // [N] current row vector
  float* part = roll_smem + N;          // [ROLL_WARPS][N] per-warp partial sums of the next vector
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < N; j += blockDim.x) v[j] = j == row ? 1.f : 0.f;
  __syncthreads();
  for (int l = L - 1; l >= 0; --l) {
    float* mine = part + warp * N;
    for (int j = lane; j < N; j += 32) mine[j] = 0.f;
    const float* base = probs + (long long)l * layer_stride + (long long)b * batch_stride;   // [H][N][N] maps of (layer l, image b)
    for (int i = warp; i < N; i += ROLL_WARPS) {
      const float vi = v[i];
      if (vi == 0.f) continue;          // warp-uniform: the first product only touches row `row`
      if (NJ > 0) {
        // row i of every head in registers (NJ * 32 >= N columns per lane-stripe): all H * NJ loads of the row are in flight
        // together, the fused row and its sum are formed once
        float f[NJ > 0 ? NJ : 1];
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) f[jj] = fusion == 0 ? 0.f : (fusion == 1 ? -INFINITY : INFINITY);
        if (HT > 0) {
          float pv[HT > 0 ? HT : 1][NJ > 0 ? NJ : 1];
#pragma unroll
          for (int h = 0; h < HT; ++h) {
            const float* prow = base + ((long long)h * N + i) * N;
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
              const int j = lane + 32 * jj;
              pv[h][jj] = j < N ? __ldg(prow + j) : 0.f;
            }
          }
#pragma unroll
          for (int h = 0; h < HT; ++h) {
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj)
              f[jj] = fusion == 0 ? f[jj] + pv[h][jj] : (fusion == 1 ? fmaxf(f[jj], pv[h][jj]) : fminf(f[jj], pv[h][jj]));
          }
        } else {
          for (int h = 0; h < H; ++h) {
            const float* prow = base + ((long long)h * N + i) * N;
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
              const int j = lane + 32 * jj;
              const float p = j < N ? __ldg(prow + j) : 0.f;
              f[jj] = fusion == 0 ? f[jj] + p : (fusion == 1 ? fmaxf(f[jj], p) : fminf(f[jj], p));
            }
          }
        }
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          const int j = lane + 32 * jj;
          if (fusion == 0) f[jj] /= float(H);
          f[jj] = j < N ? 0.5f * (f[jj] + (i == j ? 1.f : 0.f)) : 0.f;
          sum += f[jj];
        }
        sum = warp_sum(sum);
        const float w = vi / sum;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          const int j = lane + 32 * jj;
          if (j < N) mine[j] += w * f[jj];
        }
        continue;
      }
      float sum = 0.f;
      for (int j = lane; j < N; j += 32) {
        float f = fusion == 0 ? 0.f : (fusion == 1 ? -INFINITY : INFINITY);
        for (int h = 0; h < H; ++h) {
          const float p = __ldg(base + ((long long)h * N + i) * N + j);
          f = fusion == 0 ? f + p : (fusion == 1 ? fmaxf(f, p) : fminf(f, p));
        }
        if (fusion == 0) f /= float(H);
        sum += 0.5f * (f + (i == j ? 1.f : 0.f));
      }
      sum = warp_sum(sum);
      const float w = vi / sum;
      for (int j = lane; j < N; j += 32) {     // second visit of the same row: served by L1/L2
        float f = fusion == 0 ? 0.f : (fusion == 1 ? -INFINITY : INFINITY);
        for (int h = 0; h < H; ++h) {
          const float p = __ldg(base + ((long long)h * N + i) * N + j);
          f = fusion == 0 ? f + p : (fusion == 1 ? fmaxf(f, p) : fminf(f, p));
        }
        if (fusion == 0) f /= float(H);
        mine[j] += w * 0.5f * (f + (i == j ? 1.f : 0.f));
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < ROLL_WARPS; ++w) acc += part[w * N + j];
      v[j] = acc;
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < N; j += blockDim.x) out[(long long)b * N + j] = v[j];
}
__global__ void rollout_eye_kernel(float* __restrict__ r, int B, int N) {
  const long long total = (long long)B * N * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = int(idx % N);
    const int i = int((idx / N) % N);
    r[idx] = i == j ? 1.f : 0.f;
  }
}

// Class-token heat map (attention_utils.py:50-67): mean over heads of row 0 of an attention map, patch columns only,
// as the sqrt-grid, bilinearly upsampled to the image (F.interpolate(mode='bilinear', align_corners=False) restated:
// src = scale * (dst + 0.5) - 0.5 clamped at 0, the upper neighbour clamped at the border).  One CTA per (row band, image):
// the g x g grid (a few hundred floats) is rebuilt per CTA in shared memory, the band is written as 16-byte stores.
__global__ void __launch_bounds__(256) cls_heatmap_kernel(const float* __restrict__ src, float* __restrict__ out, long long batch_stride,
                                                          long long head_stride, int H, int n_prefix, int g, int OH, int OW,
                                                          int band, float sh, float sw) {
  extern __shared__ float s_grid[];   // [g*g]
  const int b = blockIdx.y;
  const float* row0 = src + (long long)b * batch_stride + n_prefix;
  const float invH = 1.f / float(H);
  for (int j = threadIdx.x; j < g * g; j += blockDim.x) {
    float acc = 0.f;
    for (int h = 0; h < H; ++h) acc += row0[(long long)h * head_stride + j];
    s_grid[j] = H > 1 ? acc * invH : acc;
  }
  __syncthreads();
  const int y_lo = blockIdx.x * band, y_hi = min(OH, y_lo + band);
  float* ob = out + (long long)b * OH * OW;
  auto sample = [&](int oy, int ox) {
    const float fy = fmaxf(sh * (float(oy) + 0.5f) - 0.5f, 0.f), fx = fmaxf(sw * (float(ox) + 0.5f) - 0.5f, 0.f);
    const int y0 = int(fy), x0 = int(fx);
    const int yp = y0 < g - 1 ? g : 0, xp = x0 < g - 1 ? 1 : 0;
    const float ly1 = fy - float(y0), lx1 = fx - float(x0);
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float* r = s_grid + y0 * g + x0;
    return ly0 * (lx0 * r[0] + lx1 * r[xp]) + ly1 * (lx0 * r[yp] + lx1 * r[yp + xp]);
  };
  if ((OW & 3) == 0) {
    const int vw = OW >> 2;
    for (int i = threadIdx.x; i < (y_hi - y_lo) * vw; i += blockDim.x) {
      const int oy = y_lo + i / vw, ox = (i % vw) << 2;
      float4 v;
      v.x = sample(oy, ox);
      v.y = sample(oy, ox + 1);
      v.z = sample(oy, ox + 2);
      v.w = sample(oy, ox + 3);
      *reinterpret_cast<float4*>(ob + (long long)oy * OW + ox) = v;
    }
  } else {
    for (int i = threadIdx.x; i < (y_hi - y_lo) * OW; i += blockDim.x) {
      const int oy = y_lo + i / OW, ox = i % OW;
      ob[(long long)oy * OW + ox] = sample(oy, ox);
    }
  }
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_cast_f32_to_16(const float* src, void* dst_bf16, void* dst_fp16, int64_t n, void* stream) {
  VITK_CHECK_ARG(src && (dst_bf16 || dst_fp16) && n >= 0, "vitk_cast_f32_to_16: bad args");
  if (n == 0) return VITK_OK;
  VITK_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dst_fp16) & 15) == 0,
                 "vitk_cast_f32_to_16: pointers must be 16-byte aligned");
  cast_kernel<<<capped_grid(n / 8 + 1, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<bf16*>(dst_bf16), reinterpret_cast<__half*>(dst_fp16), n);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_colsum16(const void* x, int32_t dtype, float* out, const float* grad_unscale, int64_t rows, int32_t dim,
                             void* stream) {
  VITK_CHECK_ARG(x && out && rows > 0 && dim > 0 && dim % 8 == 0, "vitk_colsum16: dim=%d must be a multiple of 8", dim);
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_colsum16: dtype must be bf16 or fp16");
  const int col_chunks = (dim / 8 + 31) / 32;
  // aim for ~4 CTAs per SM in total
  int slabs = (num_sms() * 4 + col_chunks - 1) / col_chunks;
  if (slabs > (rows + 63) / 64) slabs = (int)((rows + 63) / 64);
  if (slabs < 1) slabs = 1;
  const int rpb = (int)((rows + slabs - 1) / slabs);
  dim3 grid(col_chunks, (unsigned)((rows + rpb - 1) / rpb));
  colsum_kernel<<<grid, dim3(32, 8), 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), int(dtype == VITK_FP16), out, grad_unscale, rows, dim, rpb);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_patchify(const float* images, void* patches, int32_t patches_dtype, int32_t B, int32_t C, int32_t H,
                             int32_t W, int32_t P, void* stream) {
  VITK_CHECK_ARG(images && patches, "vitk_patchify: null pointer");
  VITK_CHECK_ARG(patches_dtype == VITK_BF16 || patches_dtype == VITK_FP16, "vitk_patchify: patches must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && C > 0 && P > 0 && P % 8 == 0 && H % P == 0 && W % P == 0,
                 "vitk_patchify: need P %% 8 == 0 and H, W divisible by P (B=%d C=%d H=%d W=%d P=%d)", B, C, H, W, P);
  const long long total = (long long)B * C * H * (W / 8);
  patchify_kernel<<<capped_grid(total, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      images, reinterpret_cast<bf16*>(patches), int(patches_dtype == VITK_FP16), B, C, H, W, P);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_patchify_hwc(const float* images, void* patches, int32_t patches_dtype, int32_t B, int32_t C, int32_t H,
                                 int32_t W, int32_t P, void* stream) {
  VITK_CHECK_ARG(images && patches, "vitk_patchify_hwc: null pointer");
  VITK_CHECK_ARG(patches_dtype == VITK_BF16 || patches_dtype == VITK_FP16, "vitk_patchify_hwc: patches must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && C > 0 && P > 0 && P % 8 == 0 && H % P == 0 && W % P == 0,
                 "vitk_patchify_hwc: need P %% 8 == 0 and H, W divisible by P (B=%d C=%d H=%d W=%d P=%d)", B, C, H, W, P);
  const long long total = (long long)B * (H / P) * (W / P) * ((long long)C * P * P / 2);
  patchify_hwc_kernel<<<capped_grid(total, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      images, reinterpret_cast<bf16*>(patches), int(patches_dtype == VITK_FP16), B, C, H, W, P);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

static DropSpec spec_of(const vitk_dropout* d) {
  return d != nullptr ? make_drop_spec(d->seed, d->p, d->site) : make_drop_spec(nullptr, 0.f, 0);
}
#define VITK_CHECK_DROP(d, spec, dim, who) \
  VITK_CHECK_ARG((spec).seed == nullptr || ((d)->p < 1.f && (dim) % 8 == 0), who ": dropout needs p < 1 and dim %% 8 == 0")

extern "C" int vitk_prefix_tokens_fwd(float* x, const float* cls_tok, const float* dist_tok, const float* pos, int32_t B,
                                      int32_t T, int32_t dim, int32_t n_prefix, const vitk_dropout* drop, void* stream) {
  VITK_CHECK_ARG(x && pos && dim % 4 == 0 && n_prefix >= 0 && n_prefix <= 2, "vitk_prefix_tokens_fwd: bad args");
  const DropSpec ds = spec_of(drop);
  VITK_CHECK_DROP(drop, ds, dim, "vitk_prefix_tokens_fwd");
  if (n_prefix == 0) return VITK_OK;
  VITK_CHECK_ARG(cls_tok && (n_prefix < 2 || dist_tok), "vitk_prefix_tokens_fwd: missing token");
  const long long total = (long long)B * n_prefix * (dim / 4);
  prefix_tokens_kernel<<<capped_grid(total, 256, 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, cls_tok, dist_tok, pos, B, T,
                                                                                                      dim, n_prefix, ds);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_tokens_bwd(const float* dx, float* dpos, float* dcls, float* ddist, void* dpatch16, int32_t dpatch_dtype,
                               float* dbias_patch, const float* grad_unscale, int32_t B, int32_t T, int32_t dim,
                               int32_t n_prefix, const vitk_dropout* drop, void* stream) {
  VITK_CHECK_ARG(dx && dim % 4 == 0 && n_prefix >= 0 && n_prefix <= 2 && T > n_prefix, "vitk_tokens_bwd: bad args");
  const DropSpec ds = spec_of(drop);
  VITK_CHECK_DROP(drop, ds, dim, "vitk_tokens_bwd");
  int threads = dim / 4;
  if (threads > 256) threads = 256;
  threads = ((threads + 31) / 32) * 32;
  dim3 grid((T + TOK_TOKS - 1) / TOK_TOKS, (B + TOK_IMGS - 1) / TOK_IMGS);
  tokens_bwd_kernel<<<grid, dim3(threads, TOK_TOKS), (size_t)threads * TOK_TOKS * sizeof(float4), reinterpret_cast<cudaStream_t>(stream)>>>(
      dx, dpos, dcls, ddist, reinterpret_cast<bf16*>(dpatch16), int(dpatch_dtype == VITK_FP16), dbias_patch, grad_unscale, B, T,
      dim, n_prefix, ds);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_gather_rows(const void* src, void* dst, int32_t B, int32_t T, int32_t n, int64_t row_bytes, void* stream) {
  VITK_CHECK_ARG(src && dst && B > 0 && T > 0 && n > 0 && n <= T && row_bytes > 0 && row_bytes % 16 == 0,
                 "vitk_gather_rows: bad args (row_bytes %% 16 == 0, 0 < n <= T)");
  VITK_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "vitk_gather_rows: 16-byte alignment");
  const int row_v = (int)(row_bytes / 16);
  const long long total = (long long)B * n * row_v;
  gather_rows_kernel<<<capped_grid(total, 256, 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), total, n * row_v, (long long)T * row_v);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_expand_rows(const void* src, void* dst, int32_t B, int32_t T, int32_t n, int64_t row_bytes, void* stream) {
  VITK_CHECK_ARG(src && dst && B > 0 && T > 0 && n > 0 && n <= T && row_bytes > 0 && row_bytes % 16 == 0,
                 "vitk_expand_rows: bad args (row_bytes %% 16 == 0, 0 < n <= T)");
  VITK_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "vitk_expand_rows: 16-byte alignment");
  const int row_v = (int)(row_bytes / 16);
  const long long total = (long long)B * T * row_v;
  expand_rows_kernel<<<capped_grid(total, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), total, n * row_v, (long long)T * row_v);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_head_fwd(const float* x, const float* gamma, const float* beta, const float* W0, const float* b0,
                             const float* W1, const float* b1, float* logits0, float* logits1, float* xhat, float* rstd,
                             float* pooled, int32_t B, int32_t T, int32_t dim, int32_t C, int32_t n_heads, float eps, void* stream) {
  VITK_CHECK_ARG(x && gamma && beta && W0 && logits0 && xhat && rstd, "vitk_head_fwd: null pointer");
  VITK_CHECK_ARG(n_heads == 1 || (n_heads == 2 && W1 && logits1), "vitk_head_fwd: n_heads must be 1 or 2");
  VITK_CHECK_ARG(T >= n_heads, "vitk_head_fwd: fewer tokens than heads");
  const int warps = B * n_heads;
  head_fwd_kernel<<<(warps + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, gamma, beta, W0, b0, W1, b1, logits0, logits1,
                                                                                      xhat, rstd, pooled, B, T, dim, C, n_heads, eps);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_head_bwd(const float* dlogits0, const float* dlogits1, const float* xhat, const float* rstd,
                             const float* gamma, const float* beta, const float* W0, const float* W1, float* dx,
                             void* dx16, int32_t dx16_dtype, float* dgamma, float* dbeta, float* dW0, float* db0, float* dW1,
                             float* db1, float* dcolsum, const float* loss_scale, const float* branch_scale,
                             const vitk_dropout* branch_drop, int32_t B, int32_t T, int32_t dim, int32_t C, int32_t n_heads,
                             void* stream) {
  VITK_CHECK_ARG(dlogits0 && xhat && rstd && gamma && beta && W0 && dx && dgamma && dbeta && dW0, "vitk_head_bwd: null pointer");
  const DropSpec ds = spec_of(branch_drop);
  VITK_CHECK_DROP(branch_drop, ds, dim, "vitk_head_bwd");
  VITK_CHECK_ARG(n_heads == 1 || (n_heads == 2 && dlogits1 && W1 && dW1), "vitk_head_bwd: n_heads must be 1 or 2");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VITK_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * T * dim * sizeof(float), st));
  if (dx16 != nullptr) VITK_CUDA(cudaMemsetAsync(dx16, 0, (size_t)B * T * dim * 2, st));
  const int warps = B * n_heads;
  head_bwd_kernel<<<(warps + 3) / 4, 128, 0, st>>>(dlogits0, dlogits1, xhat, rstd, gamma, W0, W1, dx,
                                                   reinterpret_cast<bf16*>(dx16), int(dx16_dtype == VITK_FP16), dgamma, dbeta, db0,
                                                   db1, dcolsum, loss_scale, branch_scale, ds, B, T, dim, C, n_heads);
  VITK_LAUNCH_CHECK();
  const long long total = (long long)n_heads * C * dim;
  head_wgrad_kernel<<<dim3(capped_grid(total, 128, 4), (B + 15) / 16), 128, 0, st>>>(dlogits0, dlogits1, xhat, gamma, beta, dW0, dW1,
                                                                                    B, dim, C, n_heads);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_dropout_mask(const vitk_dropout* drop, float* factors, int64_t rows, int32_t cols, void* stream) {
  VITK_CHECK_ARG(drop && drop->seed && factors && rows > 0 && cols > 0 && cols % 8 == 0 && drop->p >= 0.f && drop->p < 1.f,
                 "vitk_dropout_mask: bad args (cols %% 8 == 0, 0 <= p < 1)");
  DropSpec ds = make_drop_spec(drop->seed, drop->p, drop->site);
  ds.seed = static_cast<const unsigned long long*>(static_cast<const void*>(drop->seed));   // p == 0 still yields the all-ones mask
  const long long n4 = rows * (long long)(cols / 4);
  dropout_mask_kernel<<<capped_grid(n4, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(factors, n4, ds);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_droppath_scale(const float* uniform, const float* drop_prob, float* scale, int32_t branches, int32_t B, int32_t T,
                                   void* stream) {
  VITK_CHECK_ARG(uniform && drop_prob && scale && branches > 0 && B > 0 && T > 0, "vitk_droppath_scale: bad args");
  const long long total = (long long)branches * B * T;
  droppath_scale_kernel<<<capped_grid(total, 256, 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(uniform, drop_prob, scale,
                                                                                                       branches, B, T);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_ensemble_probs(const float* logits, const float* weights, float* probs, int64_t* pred, int32_t F, int32_t B,
                                   int32_t C, void* stream) {
  VITK_CHECK_ARG(logits && weights && probs && pred && F > 0 && B > 0 && C > 0, "vitk_ensemble_probs: bad args");
  ensemble_kernel<<<(B + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, weights, probs,
                                                                                      reinterpret_cast<long long*>(pred), F, B, C);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_attention_rollout(const float* probs, float* rollout, float* scratch, int32_t L, int32_t B, int32_t H,
                                      int32_t N, int32_t fusion, void* stream) {
  VITK_CHECK_ARG(probs && rollout && scratch && L > 0 && B > 0 && H > 0 && N > 0 && fusion >= 0 && fusion <= 2,
                 "vitk_attention_rollout: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // scratch: [2][B,N,N] -> a (fused layer matrix) and the ping-pong partner of `rollout`
  float* a = scratch;
  float* other = scratch + (size_t)B * N * N;
  float* cur = (L % 2 == 0) ? rollout : other;  // after L swaps the result lands in `rollout`
  float* nxt = (L % 2 == 0) ? other : rollout;
  rollout_eye_kernel<<<capped_grid((long long)B * N * N, 256, 4), 256, 0, st>>>(cur, B, N);
  VITK_LAUNCH_CHECK();
  for (int l = 0; l < L; ++l) {
    const long long rows = (long long)B * N;
    rollout_fuse_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, st>>>(probs, a, l, B, H, N, fusion);
    VITK_LAUNCH_CHECK();
    dim3 grid((N + 127) / 128, N, B);
    rollout_matmul_kernel<<<grid, 128, 0, st>>>(a, cur, nxt, N);
    VITK_LAUNCH_CHECK();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  return VITK_OK;
}

extern "C" int vitk_attention_rollout_row(const float* probs, float* out, int64_t layer_stride, int64_t batch_stride, int32_t L,
                                          int32_t B, int32_t H, int32_t N, int32_t row, int32_t fusion, void* stream) {
  VITK_CHECK_ARG(layer_stride >= (int64_t)H * N * N && batch_stride >= (int64_t)H * N * N, "vitk_attention_rollout_row: bad strides");
  VITK_CHECK_ARG(probs && out && L > 0 && B > 0 && H > 0 && N > 0 && row >= 0 && row < N && fusion >= 0 && fusion <= 2,
                 "vitk_attention_rollout_row: bad args");
  const size_t smem = (size_t)(ROLL_WARPS + 1) * N * sizeof(float);
  VITK_CHECK_ARG(smem <= 200 * 1024, "vitk_attention_rollout_row: sequence too long (N=%d)", N);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nj = (N + 31) / 32;
#define VITK_ROLL(NJ_, HT_) rollout_row_kernel<NJ_, HT_><<<B, ROLL_WARPS * 32, smem, st>>>(probs, out, layer_stride, batch_stride, L, B, H, N, row, fusion)
  if (nj <= 7) {
    if (H == 3) VITK_ROLL(7, 3);
    else if (H == 6) VITK_ROLL(7, 6);
    else if (H == 12) VITK_ROLL(7, 12);
    else VITK_ROLL(7, 0);
  } else if (nj == 8) {
    if (H == 3) VITK_ROLL(8, 3);
    else if (H == 12) VITK_ROLL(8, 12);
    else VITK_ROLL(8, 0);
  } else {
    static size_t configured = 0;
    if (smem > 48 * 1024 && configured < smem) {
      VITK_CUDA(cudaFuncSetAttribute(rollout_row_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = 200 * 1024;
    }
    VITK_ROLL(0, 0);
  }
#undef VITK_ROLL
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_cls_attention_heatmap(const float* src, float* out, int64_t batch_stride, int64_t head_stride, int32_t B,
                                          int32_t H, int32_t n_prefix, int32_t grid, int32_t out_h, int32_t out_w, void* stream) {
  VITK_CHECK_ARG(src && out && B > 0 && H > 0 && n_prefix >= 0 && grid > 0 && out_h > 0 && out_w > 0,
                 "vitk_cls_attention_heatmap: bad args");
  VITK_CHECK_ARG(grid <= 96, "vitk_cls_attention_heatmap: grid %d x %d too large", grid, grid);
  VITK_CHECK_ARG(batch_stride >= (int64_t)n_prefix + (int64_t)grid * grid && (H == 1 || head_stride > 0),
                 "vitk_cls_attention_heatmap: bad strides");
  VITK_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "vitk_cls_attention_heatmap: out must be 16-byte aligned");
  // bands sized so that the launch covers the SMs a few times over while every CTA still amortises its grid rebuild
  int band = 32;
  while (band > 4 && (long long)B * ((out_h + band - 1) / band) < 4LL * num_sms()) band >>= 1;
  const dim3 g3((out_h + band - 1) / band, B);
  cls_heatmap_kernel<<<g3, 256, (size_t)grid * grid * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      src, out, batch_stride, head_stride, H, n_prefix, grid, out_h, out_w, band, float(grid) / float(out_h),
      float(grid) / float(out_w));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
