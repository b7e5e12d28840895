// layernorm.cu -- LayerNorm forward / backward: a group of LPR lanes (8/16/32) owns one token row, 128-bit
// accesses, shuffle reductions inside the group.  Replaces nn.LayerNorm(dim) (eps 1e-5, affine) at
// vision_transformer_base.py:263,273 (norm1/norm2 of every Block).
//
// The residual stream stays fp32 (x in, dx out); the normalised activations feeding the tensor-core GEMMs are
// emitted in 16 bits.  Backward also folds in the residual gradient (dx = dres + LN'(dy)), emits the 16-bit copy
// the next dgrad/wgrad GEMM consumes and the column sum of dx (= bias gradient of the Linear that wrote the
// residual branch), so the residual gradient is read once and written once per LayerNorm.
//
// Lane mapping: dim/4 float4 vectors per row are spread over LPR lanes x CHUNKS; D=192 -> 16 lanes x 3 (two rows
// per warp, every lane busy), D=768 -> 32 lanes x 6.  HBM roofline (algorithmic bytes per row, D = dim):
//   fwd: 4D (x) + 2D (y) + 8 (stats)            bwd: 2D (dy) + 4D (x) + 4D (dres) + 4D (dx) + 2D (dx16)
#include "vitk_common.cuh"

namespace vitk {
namespace {

constexpr int LN_WARPS = 8;

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR, int CHUNKS, bool PF = false>
__global__ void __launch_bounds__(LN_WARPS * 32)
    ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  __nv_bfloat16* __restrict__ y, int y_fp16, float* __restrict__ mean, float* __restrict__ rstd, long long rows,
                  int dim, float eps) {
  constexpr int RPW = 32 / LPR;  // rows per warp
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, gl = lane % LPR;
  const int nvec = dim >> 2;
  const float inv_dim = 1.f / float(dim);
  float4 gm[CHUNKS], bt[CHUNKS];
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int i = gl + LPR * c;
    gm[c] = i < nvec ? ldg_f4(gamma + 4 * i) : make_float4(0, 0, 0, 0);
    bt[c] = i < nvec ? ldg_f4(beta + 4 * i) : make_float4(0, 0, 0, 0);
  }
  const long long stride = (long long)gridDim.x * LN_WARPS * RPW;
  // PF: the next row's loads are in flight while this row is reduced and stored (a lane group walks only a few rows, so the
  // exposed DRAM latency per row is what bounds the short-row case)
  float4 nv[CHUNKS];
  auto issue = [&](long long r0) {
    const long long row = r0 + sub;
    const bool ok = row < rows;
    const float* xr = x + (ok ? row : 0) * dim;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = gl + LPR * c;
      nv[c] = (ok && i < nvec) ? ldg_f4(xr + 4 * i) : make_float4(0, 0, 0, 0);
    }
  };
  const long long first = ((long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5)) * RPW;
  if (PF && first < rows) issue(first);
  for (long long row0 = first; row0 < rows; row0 += stride) {
    const long long row = row0 + sub;
    const bool ok = row < rows;
    if (!PF) issue(row0);
    float4 v[CHUNKS];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      v[c] = nv[c];
      s += v[c].x + v[c].y + v[c].z + v[c].w;
    }
    if (PF && row0 + stride < rows) issue(row0 + stride);
    const float mu = group_sum<LPR>(s) * inv_dim;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = gl + LPR * c;
      if (i < nvec) {
        const float a = v[c].x - mu, b = v[c].y - mu, cc = v[c].z - mu, d = v[c].w - mu;
        q += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rs = rsqrtf(group_sum<LPR>(q) * inv_dim + eps);
    if (!ok) continue;
    __nv_bfloat16* yr = y + row * dim;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = gl + LPR * c;
      if (i < nvec) {
        const float a = (v[c].x - mu) * rs * gm[c].x + bt[c].x;
        const float b = (v[c].y - mu) * rs * gm[c].y + bt[c].y;
        const float cc = (v[c].z - mu) * rs * gm[c].z + bt[c].z;
        const float d = (v[c].w - mu) * rs * gm[c].w + bt[c].w;
        *reinterpret_cast<uint2*>(yr + 4 * i) = make_uint2(pack16(a, b, y_fp16), pack16(cc, d, y_fp16));
      }
    }
    if (gl == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

template <int LPR, int CHUNKS>
__global__ void __launch_bounds__(LN_WARPS * 32)
    ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_fp16, const float* __restrict__ x, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dres,
                  float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16, int dx_fp16, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, float* __restrict__ dcolsum, const float* __restrict__ unscale,
                  const float* __restrict__ branch_scale, DropSpec drop, long long rows, int dim) {
  constexpr int RPW = 32 / LPR;
  const unsigned long long dseed = drop.seed != nullptr ? __ldg(drop.seed) : 0ull;
  extern __shared__ float red[];  // [3][dim]
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, gl = lane % LPR;
  const int nvec = dim >> 2;
  const float inv_dim = 1.f / float(dim);
  for (int i = threadIdx.x; i < 3 * dim; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  float4 gm[CHUNKS], dg[CHUNKS], db[CHUNKS], dc[CHUNKS];
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int i = gl + LPR * c;
    gm[c] = i < nvec ? ldg_f4(gamma + 4 * i) : make_float4(0, 0, 0, 0);
    dg[c] = db[c] = dc[c] = make_float4(0, 0, 0, 0);
  }
  const long long stride = (long long)gridDim.x * LN_WARPS * RPW;
  for (long long row0 = ((long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5)) * RPW; row0 < rows; row0 += stride) {
    const long long row = row0 + sub;
    const bool ok = row < rows;
    const long long rr = ok ? row : 0;
    const float mu = mean[rr], rs = rstd[rr];
    const float* xr = x + rr * dim;
    const __nv_bfloat16* dyr = dy + rr * dim;
    // issue every load of the row before any arithmetic (memory-level parallelism)
    float4 xv[CHUNKS], rv[CHUNKS];
    uint2 dyu[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = gl + LPR * c;
      const bool act = ok && i < nvec;
      xv[c] = act ? ldg_f4(xr + 4 * i) : make_float4(0, 0, 0, 0);
      dyu[c] = act ? ldg_u2(dyr + 4 * i) : make_uint2(0, 0);
      rv[c] = (act && dres != nullptr) ? ldg_f4(dres + rr * dim + 4 * i) : make_float4(0, 0, 0, 0);
    }
    float4 xh[CHUNKS], g[CHUNKS];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const float2 d01 = unpack16(dyu[c].x, dy_fp16), d23 = unpack16(dyu[c].y, dy_fp16);
      xh[c] = make_float4((xv[c].x - mu) * rs, (xv[c].y - mu) * rs, (xv[c].z - mu) * rs, (xv[c].w - mu) * rs);
      const int i = gl + LPR * c;
      if (!(ok && i < nvec)) xh[c] = make_float4(0, 0, 0, 0);
      g[c] = make_float4(d01.x * gm[c].x, d01.y * gm[c].y, d23.x * gm[c].z, d23.y * gm[c].w);
      dg[c].x += d01.x * xh[c].x; dg[c].y += d01.y * xh[c].y; dg[c].z += d23.x * xh[c].z; dg[c].w += d23.y * xh[c].w;
      db[c].x += d01.x; db[c].y += d01.y; db[c].z += d23.x; db[c].w += d23.y;
      s1 += g[c].x + g[c].y + g[c].z + g[c].w;
      s2 += g[c].x * xh[c].x + g[c].y * xh[c].y + g[c].z * xh[c].z + g[c].w * xh[c].w;
    }
    const float m1 = group_sum<LPR>(s1) * inv_dim;
    const float m2 = group_sum<LPR>(s2) * inv_dim;
    if (!ok) continue;
    float* dxr = dx + row * dim;
    const float bs = branch_scale != nullptr ? __ldg(branch_scale + row) : 1.f;   // stochastic-depth scale of the branch below
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i = gl + LPR * c;
      if (i < nvec) {
        const float4 o = make_float4(rs * (g[c].x - m1 - xh[c].x * m2) + rv[c].x, rs * (g[c].y - m1 - xh[c].y * m2) + rv[c].y,
                                     rs * (g[c].z - m1 - xh[c].z * m2) + rv[c].z, rs * (g[c].w - m1 - xh[c].w * m2) + rv[c].w);
        *reinterpret_cast<float4*>(dxr + 4 * i) = o;
        float4 ob = make_float4(o.x * bs, o.y * bs, o.z * bs, o.w * bs);
        if (drop.seed != nullptr) {   // dropout mask of the branch output this gradient enters
          const float4 m = drop_factors4(drop, dseed, ((unsigned long long)row * dim + 4 * i) >> 2);
          ob.x *= m.x; ob.y *= m.y; ob.z *= m.z; ob.w *= m.w;
        }
        if (dx16 != nullptr)
          *reinterpret_cast<uint2*>(dx16 + row * dim + 4 * i) = make_uint2(pack16(ob.x, ob.y, dx_fp16), pack16(ob.z, ob.w, dx_fp16));
        dc[c].x += ob.x; dc[c].y += ob.y; dc[c].z += ob.z; dc[c].w += ob.w;
      }
    }
  }
  // block reduction of the three column sums (shared atomics), then one global atomic per column per block
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int i = gl + LPR * c;
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


// ------------------------------------------------------------------ backward, "column owner" layout
// A CTA of 192 threads covers 192 / V rows at a time (V = dim / 4 float4 columns per row: 48 for D = 192, 192 for D = 768):
// every thread owns ONE float4 column, so the three column-sum accumulators (dgamma, dbeta, bias gradient) cost 12
// registers instead of 12 per chunk per lane, and R rows per thread are in flight at once (R x 40 bytes of loads per
// thread, 24+ resident warps per SM).  Row statistics are reduced inside G-lane groups by shuffles and across groups
// through a double-buffered shared-memory table (one __syncthreads per batch of R x slots rows).
template <int V, int R, bool DROP, bool PF = false, int MINB = 5>
__global__ void __launch_bounds__(192, MINB)
    ln_bwd_cols_kernel(const __nv_bfloat16* __restrict__ dy, int dy_fp16, const float* __restrict__ x, const float* __restrict__ mean,
                       const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dres,
                       float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16, int dx_fp16, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ dcolsum, const float* __restrict__ unscale,
                       const float* __restrict__ branch_scale, DropSpec drop, long long rows) {
  constexpr int T = 192;
  constexpr int SLOTS = T / V;                 // rows processed side by side
  constexpr int G = (V % 32 == 0) ? 32 : 16;   // lanes per shuffle group (never straddles two rows)
  constexpr int GPR = V / G;                   // groups per row
  constexpr int DIM = V * 4;
  __shared__ float2 part[2][R][SLOTS][GPR];
  __shared__ float4 fin[3][SLOTS][V];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int slot = tid / V, cv = tid - slot * V;
  const int grp = cv / G, gl = cv % G;
  const float inv_dim = 1.f / float(DIM);
  const float4 gm = ldg_f4(gamma + 4 * cv);
  const unsigned long long dseed = DROP ? __ldg(drop.seed) : 0ull;   // DROP: compile-time, so the plain kernel pays no registers
  float4 dg = make_float4(0, 0, 0, 0), db = dg, dc = dg;
  const long long batch = (long long)SLOTS * R;
  int it = 0;
  // PF: the loads of the CTA's NEXT batch are issued before this batch is reduced, so a batch's DRAM latency hides behind
  // the previous batch's shuffles / barrier / stores instead of adding to them (short rows: ~9 dependent batches per CTA)
  float4 nxv[R], nrv[R];
  uint2 ndyu[R];
  float nmu[R], nrs[R];
  auto issue = [&](long long base) {
#pragma unroll
    for (int r = 0; r < R; ++r) {  // every load of the batch in flight before any arithmetic
      const long long row = base + r * SLOTS + slot;
      const bool ok = row < rows;
      const long long rr = ok ? row : 0;
      nxv[r] = ok ? ldg_f4(x + rr * DIM + 4 * cv) : make_float4(0, 0, 0, 0);
      ndyu[r] = ok ? ldg_u2(dy + rr * DIM + 4 * cv) : make_uint2(0, 0);
      nrv[r] = (ok && dres != nullptr) ? ldg_f4(dres + rr * DIM + 4 * cv) : make_float4(0, 0, 0, 0);
      nmu[r] = __ldg(mean + rr);
      nrs[r] = __ldg(rstd + rr);
    }
  };
  const long long step = (long long)gridDim.x * batch;
  if (PF && (long long)blockIdx.x * batch < rows) issue((long long)blockIdx.x * batch);
  for (long long base = (long long)blockIdx.x * batch; base < rows; base += step, ++it) {
    float4 xv[R], rv[R];
    uint2 dyu[R];
    float mu[R], rs[R];
    if (!PF) issue(base);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      xv[r] = nxv[r];
      dyu[r] = ndyu[r];
      rv[r] = nrv[r];
      mu[r] = nmu[r];
      rs[r] = nrs[r];
    }
    if (PF && base + step < rows) issue(base + step);
    float4 xh[R], g[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = base + r * SLOTS + slot;
      const float2 d01 = unpack16(dyu[r].x, dy_fp16), d23 = unpack16(dyu[r].y, dy_fp16);
      xh[r] = make_float4((xv[r].x - mu[r]) * rs[r], (xv[r].y - mu[r]) * rs[r], (xv[r].z - mu[r]) * rs[r], (xv[r].w - mu[r]) * rs[r]);
      if (row >= rows) xh[r] = make_float4(0, 0, 0, 0);
      g[r] = make_float4(d01.x * gm.x, d01.y * gm.y, d23.x * gm.z, d23.y * gm.w);
      dg.x += d01.x * xh[r].x; dg.y += d01.y * xh[r].y; dg.z += d23.x * xh[r].z; dg.w += d23.y * xh[r].w;
      db.x += d01.x; db.y += d01.y; db.z += d23.x; db.w += d23.y;
      float s1 = g[r].x + g[r].y + g[r].z + g[r].w;
      float s2 = g[r].x * xh[r].x + g[r].y * xh[r].y + g[r].z * xh[r].z + g[r].w * xh[r].w;
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (gl == 0) part[it & 1][r][slot][grp] = make_float2(s1, s2);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = base + r * SLOTS + slot;
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int k = 0; k < GPR; ++k) {
        const float2 pp = part[it & 1][r][slot][k];
        m1 += pp.x;
        m2 += pp.y;
      }
      m1 *= inv_dim;
      m2 *= inv_dim;
      if (row < rows) {
        const float4 o = make_float4(rs[r] * (g[r].x - m1 - xh[r].x * m2) + rv[r].x, rs[r] * (g[r].y - m1 - xh[r].y * m2) + rv[r].y,
                                     rs[r] * (g[r].z - m1 - xh[r].z * m2) + rv[r].z, rs[r] * (g[r].w - m1 - xh[r].w * m2) + rv[r].w);
        *reinterpret_cast<float4*>(dx + row * DIM + 4 * cv) = o;
        // the 16-bit copy and the bias gradient belong to the residual BRANCH below: scaled by its stochastic-depth factor
        const float bs = branch_scale != nullptr ? __ldg(branch_scale + row) : 1.f;
        float4 ob = make_float4(o.x * bs, o.y * bs, o.z * bs, o.w * bs);
        if (DROP) {   // ... and by the dropout mask of that branch's output (recomputed, never stored)
          const float4 m = drop_factors4(drop, dseed, ((unsigned long long)row * DIM + 4 * cv) >> 2);
          ob.x *= m.x; ob.y *= m.y; ob.z *= m.z; ob.w *= m.w;
        }
        if (dx16 != nullptr)
          *reinterpret_cast<uint2*>(dx16 + row * DIM + 4 * cv) = make_uint2(pack16(ob.x, ob.y, dx_fp16), pack16(ob.z, ob.w, dx_fp16));
        dc.x += ob.x; dc.y += ob.y; dc.z += ob.z; dc.w += ob.w;
      }
    }
  }
  // column sums: combine the SLOTS partial owners of each column, then one global atomic per column per CTA
  fin[0][slot][cv] = dg;
  fin[1][slot][cv] = db;
  fin[2][slot][cv] = dc;
  __syncthreads();
  const float u = unscale != nullptr ? __ldg(unscale) : 1.f;
  for (int i = tid; i < 3 * DIM; i += T) {
    const int which = i / DIM, c = i - which * DIM;
    if (which == 2 && dcolsum == nullptr) continue;
    float acc = 0.f;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) acc += reinterpret_cast<const float*>(&fin[which][sl][0])[c];
    float* dst = which == 0 ? dgamma : which == 1 ? dbeta : dcolsum;
    atomicAdd(dst + c, acc * u);
  }
}

struct LnCfg {
  int lpr, chunks;
};
// smallest lane group that covers the row with <= 8 float4 per lane, preferring full lane utilisation
LnCfg ln_config(int dim) {
  const int nvec = dim / 4;
  const int lprs[3] = {8, 16, 32};
  for (int lpr : lprs) {
    const int ch = (nvec + lpr - 1) / lpr;
    if (ch <= 4) return {lpr, ch <= 1 ? 1 : ch <= 2 ? 2 : ch <= 3 ? 3 : 4};
  }
  const int ch = (nvec + 31) / 32;
  return {32, ch <= 6 ? 6 : 8};
}

int ln_grid(long long rows, int rows_per_block, int ctas_per_sm) {
  long long blocks = (rows + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)num_sms() * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace vitk

using namespace vitk;

#define LN_CASE(L, C, ...)                      \
  if (cfg.lpr == L && cfg.chunks == C) {        \
    constexpr int L_ = L, C_ = C;               \
    __VA_ARGS__;                                \
    launched = true;                            \
  }
#define LN_DISPATCH(...)                                                                                       \
  {                                                                                                            \
    bool launched = false;                                                                                     \
    LN_CASE(8, 1, __VA_ARGS__) LN_CASE(8, 2, __VA_ARGS__) LN_CASE(8, 3, __VA_ARGS__) LN_CASE(8, 4, __VA_ARGS__) \
    LN_CASE(16, 1, __VA_ARGS__) LN_CASE(16, 2, __VA_ARGS__) LN_CASE(16, 3, __VA_ARGS__) LN_CASE(16, 4, __VA_ARGS__) \
    LN_CASE(32, 1, __VA_ARGS__) LN_CASE(32, 2, __VA_ARGS__) LN_CASE(32, 3, __VA_ARGS__) LN_CASE(32, 4, __VA_ARGS__) \
    LN_CASE(32, 6, __VA_ARGS__) LN_CASE(32, 8, __VA_ARGS__)                                                     \
    if (!launched) {                                                                                           \
      set_error("layernorm: unsupported dim");                                                                 \
      return VITK_ERR_UNSUPPORTED;                                                                             \
    }                                                                                                          \
  }

extern "C" int vitk_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int32_t y_dtype, float* mean,
                                  float* rstd, int64_t rows, int32_t dim, float eps, void* stream) {
  VITK_CHECK_ARG(x && gamma && beta && y && mean && rstd, "vitk_layernorm_fwd: null pointer");
  VITK_CHECK_ARG(y_dtype == VITK_BF16 || y_dtype == VITK_FP16, "vitk_layernorm_fwd: y must be bf16 or fp16");
  VITK_CHECK_ARG(rows > 0 && dim > 0 && dim % 4 == 0 && dim <= 1024, "vitk_layernorm_fwd: dim=%d must be a multiple of 4, <= 1024", dim);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const LnCfg cfg = ln_config(dim);
  const int rpb = LN_WARPS * (32 / cfg.lpr);
  // long rows (>= 4 float4 per lane, e.g. D = 768) prefetch the next row: 51 -> 45 us on ViT-B/16 batch 256; short rows gain
  // nothing (measured) and keep the smaller-register kernel.  VITK_LN_FWD_VARIANT=0|1 forces one of them (A/B runs).
  static const int variant = [] { const char* e = getenv("VITK_LN_FWD_VARIANT"); return e != nullptr ? atoi(e) : -1; }();
  if (variant == 1 || (variant < 0 && cfg.chunks >= 4)) {
    LN_DISPATCH(VITK_CUDA(launch_pdl(ln_fwd_kernel<L_, C_, true>, dim3(ln_grid(rows, rpb, 6)), dim3(LN_WARPS * 32), 0, st, x, gamma,
                                     beta, reinterpret_cast<__nv_bfloat16*>(y), int(y_dtype == VITK_FP16), mean, rstd, (long long)rows,
                                     dim, eps)));
  } else {
    LN_DISPATCH(VITK_CUDA(launch_pdl(ln_fwd_kernel<L_, C_, false>, dim3(ln_grid(rows, rpb, 8)), dim3(LN_WARPS * 32), 0, st, x, gamma,
                                     beta, reinterpret_cast<__nv_bfloat16*>(y), int(y_dtype == VITK_FP16), mean, rstd, (long long)rows,
                                     dim, eps)));
  }
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* mean, const float* rstd,
                                  const float* gamma, const float* dres, float* dx, void* dx16, int32_t dx16_dtype,
                                  float* dgamma, float* dbeta, float* dcolsum, const float* grad_unscale,
                                  const float* branch_scale, const vitk_dropout* branch_drop, int64_t rows, int32_t dim,
                                  void* stream) {
  VITK_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta, "vitk_layernorm_bwd: null pointer");
  const DropSpec drop = branch_drop != nullptr ? make_drop_spec(branch_drop->seed, branch_drop->p, branch_drop->site)
                                               : make_drop_spec(nullptr, 0.f, 0);
  VITK_CHECK_ARG(drop.seed == nullptr || (branch_drop->p < 1.f && dim % 8 == 0), "vitk_layernorm_bwd: dropout needs p < 1, dim %% 8 == 0");
  VITK_CHECK_ARG((dy_dtype == VITK_BF16 || dy_dtype == VITK_FP16) && (dx16_dtype == VITK_BF16 || dx16_dtype == VITK_FP16),
                 "vitk_layernorm_bwd: dy / dx16 must be bf16 or fp16");
  VITK_CHECK_ARG(rows > 0 && dim > 0 && dim % 4 == 0 && dim <= 1024, "vitk_layernorm_bwd: dim=%d must be a multiple of 4, <= 1024", dim);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    // dims whose float4 column count divides 192 take the column-owner kernel (every ViT/DeiT width: 192, 384, 768; and 128)
    const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
    __nv_bfloat16* dx16p = reinterpret_cast<__nv_bfloat16*>(dx16);
    const int f_dy = int(dy_dtype == VITK_FP16), f_dx = int(dx16_dtype == VITK_FP16);
#define LN_LAUNCH(K_)                                                                                                       \
  VITK_CUDA(launch_pdl(K_, dim3((unsigned)blocks), dim3(192), 0, st, dyp, f_dy, x, mean, rstd, gamma, dres, dx, dx16p, f_dx, \
                       dgamma, dbeta, dcolsum, grad_unscale, branch_scale, drop, (long long)rows))
#define LN_COLS(V_, R_, PF_, MINB_)                                                                                        \
  {                                                                                                                        \
    const long long batch = (long long)(192 / V_) * R_;                                                                    \
    long long blocks = (rows + batch - 1) / batch;                                                                         \
    const long long cap = (long long)num_sms() * MINB_;                                                                    \
    if (blocks > cap) blocks = cap;                                                                                        \
    if (drop.seed != nullptr)                                                                                              \
      LN_LAUNCH((ln_bwd_cols_kernel<V_, R_, true, PF_, MINB_>));                                                           \
    else                                                                                                                   \
      LN_LAUNCH((ln_bwd_cols_kernel<V_, R_, false, PF_, MINB_>));                                                          \
    VITK_LAUNCH_CHECK();                                                                                                   \
    return VITK_OK;                                                                                                        \
  }
    // D = 192 (DeiT-tiny / ViT-tiny): prefetching variant on 2 CTAs per SM -- 43.0 -> 38.9 us at batch 256 (fewer CTAs also
    // means fewer same-address atomics on the 3 x 192 column sums); measured alternatives: prefetch on 4 CTAs/SM 41.0, on 3
    // CTAs/SM 45.1, four rows in flight without prefetch 47.2.  D = 768 is at 0.91 of its HBM floor and does not move with
    // prefetching (104.4 vs 104.5 us).  VITK_LN_BWD_VARIANT=0 restores the non-prefetching kernel for A/B runs.
    static const int variant = [] { const char* e = getenv("VITK_LN_BWD_VARIANT"); return e != nullptr ? atoi(e) : -1; }();
    if (dim == 768) LN_COLS(192, 2, false, 5)
    if (dim == 384) LN_COLS(96, 2, false, 5)
    if (dim == 192 && variant == 0) LN_COLS(48, 2, false, 5)
    if (dim == 192) LN_COLS(48, 2, true, 2)
    if (dim == 128) LN_COLS(32, 2, false, 5)
#undef LN_COLS
#undef LN_LAUNCH
  }
  const LnCfg cfg = ln_config(dim);
  const size_t smem = 3 * (size_t)dim * sizeof(float);
  // each CTA should sweep several rows per lane group so the closing 3*dim global atomics are amortised
  const int rpb = LN_WARPS * (32 / cfg.lpr) * 4;
  LN_DISPATCH((ln_bwd_kernel<L_, C_><<<ln_grid(rows, rpb, 6), LN_WARPS * 32, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), int(dy_dtype == VITK_FP16), x, mean, rstd, gamma, dres, dx,
      reinterpret_cast<__nv_bfloat16*>(dx16), int(dx16_dtype == VITK_FP16), dgamma, dbeta, dcolsum, grad_unscale, branch_scale, drop,
      rows, dim)));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
