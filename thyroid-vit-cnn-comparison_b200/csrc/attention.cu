// attention.cu -- fused softmax attention for head_dim 64 (every ViT/DeiT variant of the
// reference: tiny 192/3, small 384/6, base 768/12).
//
// Replaces Attention.forward's q@k^T*scale -> softmax -> attn@v (vision_transformer_base.py:
// 182-191) and its autograd backward.  The [B,H,N,N] score tensor is never written: forward is
// a flash-style online softmax over 64-key blocks and saves only the log-sum-exp; backward
// recomputes P from q, k and lse.
//
// Layouts are the reference's own: qkv [B,N,3,H,64] is what nn.Linear(D,3D) emits (:178), the
// output [B,N,H,64] is (attn@v).transpose(1,2).reshape(B,N,C) (:191) -- no permute copies.
//
// Element type: fp16 or bf16 for ALL of qkv / out / dout / dqkv of one call (template H16); the engine uses
// fp16 (gradients are loss-scaled), bf16 stays available.  Softmax statistics and accumulators are fp32.
//
// Round-1 implementation: warp-level mma.sync m16n8k16 (HMMA) with ldmatrix from
// XOR-swizzled shared tiles.  The tcgen05/TMEM version (S and P resident in TMEM, N<=256 in a
// single tile) is the planned replacement; the C-ABI below is already the one it will keep.
#include <cstdlib>

#include "vitk_common.cuh"

namespace vitk {
namespace {

constexpr int DH = 64;
constexpr int TILE = 64;  // rows per tile (queries per CTA, keys per block)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

typedef __nv_bfloat16 bf16;

// ---- shared tile [64][64] bf16, 16-byte chunks XOR-swizzled by row -----------------------
__device__ __forceinline__ int tile_off(int r, int chunk) { return r * DH + ((chunk ^ (r & 7)) << 3); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Load rows [row0, row0+64) of a strided [rows, 64] bf16 matrix into a swizzled tile;
// rows >= nrows are zero-filled.
__device__ __forceinline__ void load_tile(bf16* tile, const bf16* base, long long row_stride, int row0, int nrows) {
  for (int i = threadIdx.x; i < TILE * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const int gr = row0 + r;
    const bool ok = gr < nrows;
    const bf16* src = base + (long long)(ok ? gr : 0) * row_stride + c * 8;
    cp_async16(tile + tile_off(r, c), src, ok);
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
template <bool H16>
__device__ __forceinline__ void mma_16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (H16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
template <bool H16>
__device__ __forceinline__ uint32_t pk(float lo, float hi) {
  if (H16) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16(lo, hi);
}
template <bool H16>
__device__ __forceinline__ float2 upk(uint32_t u) {
  if (H16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return unpack_bf16(u);
}

// A-operand fragments (16 rows x 64 k) of rows [r0, r0+16) of a tile: 4 k16 steps.
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], const bf16* tile, int r0, int lane) {
  const int r = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(f[ks], tile + tile_off(r, ks * 2 + (lane >> 4)));
}
// B operand "rows are n, k contiguous" (Q/K/V/dO tile used as X in  A * X^T):
// returns fragments for n-tiles 2*np and 2*np+1 at k16 step ks.
__device__ __forceinline__ void load_b_nt(uint32_t (&f)[4], const bf16* tile, int np, int ks, int lane) {
  const int r = np * 16 + (lane & 7) + (lane >> 4) * 8;
  ldsm_x4(f, tile + tile_off(r, ks * 2 + ((lane >> 3) & 1)));
}
// B operand "rows are k, n contiguous" (tile used as X in  A * X): k16 step ks, n-tiles 2*np, 2*np+1.
__device__ __forceinline__ void load_b_kn(uint32_t (&f)[4], const bf16* tile, int ks, int np, int lane) {
  const int r = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
  ldsm_x4_t(f, tile + tile_off(r, np * 2 + (lane >> 4)));
}

// acc[8][4] (16 x 64) += A(16 x 64 via frags) * X^T, X = tile rows as n
template <bool H16>
__device__ __forceinline__ void gemm_a_xt(float (&acc)[8][4], const uint32_t (&a)[4][4], const bf16* tile, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      load_b_nt(b, tile, np, ks, lane);
      mma_16<H16>(acc[2 * np], a[ks], b[0], b[1]);
      mma_16<H16>(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}
// acc[8][4] (16 x 64) += P(16 x 64, given as fp32 C-fragments, converted to bf16) * X, X = tile rows as k
template <bool H16>
__device__ __forceinline__ void gemm_p_x(float (&acc)[8][4], const float (&p)[8][4], const bf16* tile, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    a[0] = pk<H16>(p[2 * ks][0], p[2 * ks][1]);
    a[1] = pk<H16>(p[2 * ks][2], p[2 * ks][3]);
    a[2] = pk<H16>(p[2 * ks + 1][0], p[2 * ks + 1][1]);
    a[3] = pk<H16>(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      load_b_kn(b, tile, ks, np, lane);
      mma_16<H16>(acc[2 * np], a, b[0], b[1]);
      mma_16<H16>(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&a)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}

// ------------------------------------------------------------------ forward
// DROP: nn.Dropout on the softmax output (Attention.attn_drop, vision_transformer_base.py:184): O = (P o M) V with the
// counter-based factors M[b,h,q,key] = 0 | 1/(1-p) of element (((b*H + h)*N + q)*Npad + key) at `drop.site`; the row sum and
// lse stay those of the unmasked P.  The backward kernels re-derive M from the same counters.
template <bool H16, bool DROP>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                       float* __restrict__ lse, int N, int H, float scale_log2, DropSpec drop,
                                                       int Npad) {
  __shared__ __align__(128) bf16 sQ[TILE * DH];
  __shared__ __align__(128) bf16 sK[2][TILE * DH];
  __shared__ __align__(128) bf16 sV[2][TILE * DH];
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long rs = 3LL * H * DH;
  const bf16* qb = qkv + ((long long)b * N * 3 + 0) * H * DH + h * DH;
  const bf16* kb = qb + (long long)H * DH;
  const bf16* vb = kb + (long long)H * DH;

  load_tile(sQ, qb, rs, q0, N);
  load_tile(sK[0], kb, rs, 0, N);
  load_tile(sV[0], vb, rs, 0, N);
  cp_async_commit();

  const int nkv = (N + TILE - 1) / TILE;
  uint32_t qf[4][4];
  float o[8][4];
  zero_acc(o);
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const unsigned long long dseed = DROP ? __ldg(drop.seed) : 0ull;

  for (int j = 0; j < nkv; ++j) {
    const int buf = j & 1;
    if (j + 1 < nkv) {
      load_tile(sK[buf ^ 1], kb, rs, (j + 1) * TILE, N);
      load_tile(sV[buf ^ 1], vb, rs, (j + 1) * TILE, N);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) load_a_frags(qf, sQ, warp * 16, lane);

    float s[8][4];
    zero_acc(s);
    gemm_a_xt<H16>(s, qf, sK[buf], lane);

    // scale into log2 domain, mask keys >= N
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = j * TILE + nt * 8 + 2 * t + (e & 1);
        const float v = key < N ? s[nt][e] * scale_log2 : -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_run[r], mx[r]);
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] - m_new[r]);
      m_run[r] = m_new[r];
    }
    float rsum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(s[nt][e] - m_new[e >> 1]);
        s[nt][e] = pv;
        rsum[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rsum[r];
    if (DROP) {   // this thread's two adjacent keys of every 8-key group share one counter block
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long q = q0 + warp * 16 + g + r * 8;
        const unsigned long long rowblk = ((unsigned long long)(((long long)b * H + h) * N + q) * Npad + j * TILE) >> 3;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint4 bits = drop_bits8(dseed, drop.site, rowblk + nt);
          const float2 f = drop_pair(t == 0 ? bits.x : t == 1 ? bits.y : t == 2 ? bits.z : bits.w, drop.thresh, drop.inv_keep);
          s[nt][2 * r] *= f.x;
          s[nt][2 * r + 1] *= f.y;
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= corr[0]; o[nt][1] *= corr[0]; o[nt][2] *= corr[1]; o[nt][3] *= corr[1];
    }
    gemm_p_x<H16>(o, s, sV[buf], lane);
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int q = q0 + warp * 16 + g + r * 8;
    if (q >= N) continue;
    const float inv = 1.f / l_run[r];
    const long long ooff = ((long long)(b * N + q) * H + h) * DH;
    bf16* orow = out + ooff;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<uint32_t*>(orow + nt * 8 + 2 * t) = pk<H16>(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
    if (t == 0) lse[((long long)b * H + h) * N + q] = m_run[r] * LN2 + logf(l_run[r]);
  }
}

// ------------------------------------------------------------------ backward: delta = rowsum(dO * O)
template <bool H16>
__global__ void attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta,
                                  int B, int N, int H) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (b, n, h)
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * N * H;
  if (row >= total) return;
  const uint32_t a = *reinterpret_cast<const uint32_t*>(out + row * DH + lane * 2);
  const uint32_t d = *reinterpret_cast<const uint32_t*>(dout + row * DH + lane * 2);
  const float2 af = upk<H16>(a), df = upk<H16>(d);
  float v = warp_sum(af.x * df.x + af.y * df.y);
  if (lane == 0) {
    const int h = int(row % H);
    const long long bn = row / H;
    const int n = int(bn % N);
    const int b = int(bn / N);
    delta[((long long)b * H + h) * N + n] = v;
  }
}

// ------------------------------------------------------------------ backward: dK, dV
// One CTA per (64-key block, head, image); each warp owns 16 keys and loops over query blocks.
struct BwdSmem {
  bf16 k[TILE * DH];
  bf16 v[TILE * DH];
  bf16 q[2][TILE * DH];
  bf16 d[2][TILE * DH];
  float lse2[2][TILE];
  float delta[2][TILE];
};

template <bool H16, bool DROP>
__global__ void __launch_bounds__(128) attn_bwd_dkdv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                            const float* __restrict__ lse, const float* __restrict__ delta,
                                                            bf16* __restrict__ dqkv, int N, int H, float scale,
                                                            float scale_log2, DropSpec drop, int Npad) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int k0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long rs = 3LL * H * DH;
  const bf16* qb = qkv + ((long long)b * N * 3) * H * DH + h * DH;
  const bf16* kb = qb + (long long)H * DH;
  const bf16* vb = kb + (long long)H * DH;
  const bf16* dob = dout + ((long long)b * N * H + h) * DH;
  const long long rso = (long long)H * DH;
  const float* lse_b = lse + ((long long)b * H + h) * N;
  const float* del_b = delta + ((long long)b * H + h) * N;

  auto load_q_block = [&](int i, int buf) {
    load_tile(sm.q[buf], qb, rs, i * TILE, N);
    load_tile(sm.d[buf], dob, rso, i * TILE, N);
    if (threadIdx.x < TILE) {
      const int q = i * TILE + threadIdx.x;
      sm.lse2[buf][threadIdx.x] = q < N ? lse_b[q] * LOG2E : 0.f;
      sm.delta[buf][threadIdx.x] = q < N ? del_b[q] : 0.f;
    }
  };

  load_tile(sm.k, kb, rs, k0, N);
  load_tile(sm.v, vb, rs, k0, N);
  load_q_block(0, 0);
  cp_async_commit();

  const int nq = (N + TILE - 1) / TILE;
  uint32_t kf[4][4], vf[4][4];
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);

  for (int i = 0; i < nq; ++i) {
    const int buf = i & 1;
    if (i + 1 < nq) {
      load_q_block(i + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (i == 0) {
      load_a_frags(kf, sm.k, warp * 16, lane);
      load_a_frags(vf, sm.v, warp * 16, lane);
    }
    // S^T (16 keys x 64 queries) = K_w Q^T
    float st[8][4];
    zero_acc(st);
    gemm_a_xt<H16>(st, kf, sm.q[buf], lane);
    // P^T = exp(S^T*scale - lse[q])
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ql = nt * 8 + 2 * t + (e & 1);
        const bool ok = (i * TILE + ql) < N;
        st[nt][e] = ok ? exp2f(st[nt][e] * scale_log2 - sm.lse2[buf][ql]) : 0.f;
      }
    }
    // dropout factors of this thread's (query, key) elements: transposed walk, so one counter block per element
    float mk[DROP ? 8 : 1][4];
    if (DROP) {
      const unsigned long long dseed = __ldg(drop.seed);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const long long q = (long long)i * TILE + nt * 8 + 2 * t + (e & 1);
          const int key = k0 + warp * 16 + g + (e >> 1) * 8;
          const unsigned long long el = (unsigned long long)(((long long)b * H + h) * N + q) * Npad + key;
          const uint4 bits = drop_bits8(dseed, drop.site, el >> 3);
          const int jj = int(el & 7);
          const uint32_t w = jj < 2 ? bits.x : jj < 4 ? bits.y : jj < 6 ? bits.z : bits.w;
          mk[nt][e] = (((jj & 1) ? (w >> 16) : (w & 0xffffu)) >= drop.thresh) ? drop.inv_keep : 0.f;
        }
      }
    }
    // dV += (P o M)^T dO
    if (DROP) {
      float pm[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) pm[nt][e] = st[nt][e] * mk[DROP ? nt : 0][e];
      }
      gemm_p_x<H16>(dv, pm, sm.d[buf], lane);
    } else {
      gemm_p_x<H16>(dv, st, sm.d[buf], lane);
    }
    // dP^T = V_w dO^T (o M)
    float dp[8][4];
    zero_acc(dp);
    gemm_a_xt<H16>(dp, vf, sm.d[buf], lane);
    // dS^T = P^T * (dP^T - delta[q])
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ql = nt * 8 + 2 * t + (e & 1);
        const float dpe = DROP ? dp[nt][e] * mk[DROP ? nt : 0][e] : dp[nt][e];
        dp[nt][e] = st[nt][e] * (dpe - sm.delta[buf][ql]);
      }
    }
    // dK += dS^T Q
    gemm_p_x<H16>(dk, dp, sm.q[buf], lane);
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = k0 + warp * 16 + g + r * 8;
    if (key >= N) continue;
    bf16* dkrow = dqkv + (((long long)(b * N + key) * 3 + 1) * H + h) * DH;
    bf16* dvrow = dqkv + (((long long)(b * N + key) * 3 + 2) * H + h) * DH;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t*>(dkrow + nt * 8 + 2 * t) = pk<H16>(dk[nt][2 * r] * scale, dk[nt][2 * r + 1] * scale);
      *reinterpret_cast<uint32_t*>(dvrow + nt * 8 + 2 * t) = pk<H16>(dv[nt][2 * r], dv[nt][2 * r + 1]);
    }
  }
}

// ------------------------------------------------------------------ backward: dQ
template <bool H16, bool DROP>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                          const float* __restrict__ lse, const float* __restrict__ delta,
                                                          bf16* __restrict__ dqkv, int N, int H, float scale,
                                                          float scale_log2, DropSpec drop, int Npad) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // reuse BwdSmem: k/v fields hold Q and dO of this CTA, q/d double buffers hold K and V blocks
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long rs = 3LL * H * DH;
  const bf16* qb = qkv + ((long long)b * N * 3) * H * DH + h * DH;
  const bf16* kb = qb + (long long)H * DH;
  const bf16* vb = kb + (long long)H * DH;
  const bf16* dob = dout + ((long long)b * N * H + h) * DH;
  const long long rso = (long long)H * DH;

  load_tile(sm.k, qb, rs, q0, N);
  load_tile(sm.v, dob, rso, q0, N);
  load_tile(sm.q[0], kb, rs, 0, N);
  load_tile(sm.d[0], vb, rs, 0, N);
  cp_async_commit();

  float lse2[2], del[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int q = q0 + warp * 16 + g + r * 8;
    lse2[r] = q < N ? lse[((long long)b * H + h) * N + q] * LOG2E : 0.f;
    del[r] = q < N ? delta[((long long)b * H + h) * N + q] : 0.f;
  }

  const int nkv = (N + TILE - 1) / TILE;
  uint32_t qf[4][4], dof[4][4];
  float dq[8][4];
  zero_acc(dq);

  for (int j = 0; j < nkv; ++j) {
    const int buf = j & 1;
    if (j + 1 < nkv) {
      load_tile(sm.q[buf ^ 1], kb, rs, (j + 1) * TILE, N);
      load_tile(sm.d[buf ^ 1], vb, rs, (j + 1) * TILE, N);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
      load_a_frags(qf, sm.k, warp * 16, lane);
      load_a_frags(dof, sm.v, warp * 16, lane);
    }
    float s[8][4];
    zero_acc(s);
    gemm_a_xt<H16>(s, qf, sm.q[buf], lane);  // S = Q K^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = j * TILE + nt * 8 + 2 * t + (e & 1);
        s[nt][e] = key < N ? exp2f(s[nt][e] * scale_log2 - lse2[e >> 1]) : 0.f;
      }
    }
    float dp[8][4];
    zero_acc(dp);
    gemm_a_xt<H16>(dp, dof, sm.d[buf], lane);  // dP = dO V^T
    if (DROP) {                                // dP o M (same counter walk as the forward)
      const unsigned long long dseed = __ldg(drop.seed);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long q = q0 + warp * 16 + g + r * 8;
        const unsigned long long rowblk = ((unsigned long long)(((long long)b * H + h) * N + q) * Npad + j * TILE) >> 3;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint4 bits = drop_bits8(dseed, drop.site, rowblk + nt);
          const float2 f = drop_pair(t == 0 ? bits.x : t == 1 ? bits.y : t == 2 ? bits.z : bits.w, drop.thresh, drop.inv_keep);
          dp[nt][2 * r] *= f.x;
          dp[nt][2 * r + 1] *= f.y;
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) dp[nt][e] = s[nt][e] * (dp[nt][e] - del[e >> 1]);
    }
    gemm_p_x<H16>(dq, dp, sm.q[buf], lane);  // dQ += dS K
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int q = q0 + warp * 16 + g + r * 8;
    if (q >= N) continue;
    bf16* dqrow = dqkv + (((long long)(b * N + q) * 3 + 0) * H + h) * DH;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<uint32_t*>(dqrow + nt * 8 + 2 * t) = pk<H16>(dq[nt][2 * r] * scale, dq[nt][2 * r + 1] * scale);
  }
}

// ------------------------------------------------------------------ eval-only attention maps
// probs[b,h,q,:] = softmax(q.k^T*scale) in fp32 -- the `attention_maps` the reference stores in
// eval mode (vision_transformer_base.py:186-188).  One warp per (b,h,q) row; not on the train path.
template <bool H16>
__global__ void attn_probs_kernel(const bf16* __restrict__ qkv, float* __restrict__ probs, long long batch_stride, int B, int N, int H,
                                  float scale) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (b,h,q)
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * H * N) return;
  const int q = int(row % N);
  const int h = int((row / N) % H);
  const int b = int(row / ((long long)N * H));
  const bf16* qp = qkv + (((long long)(b * N + q) * 3 + 0) * H + h) * DH;
  float qv[DH];
#pragma unroll
  for (int d = 0; d < DH; d += 2) {
    const float2 f = upk<H16>(*reinterpret_cast<const uint32_t*>(qp + d));
    qv[d] = f.x; qv[d + 1] = f.y;
  }
  float* prow = probs + (long long)b * batch_stride + ((long long)h * N + q) * N;
  float mx = -INFINITY;
  for (int k = lane; k < N; k += 32) {
    const bf16* kp = qkv + (((long long)(b * N + k) * 3 + 1) * H + h) * DH;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < DH; d += 8) {
      const uint4 u = ldg_u4(kp + d);
      const float2 f0 = upk<H16>(u.x), f1 = upk<H16>(u.y), f2 = upk<H16>(u.z), f3 = upk<H16>(u.w);
      acc += qv[d] * f0.x + qv[d + 1] * f0.y + qv[d + 2] * f1.x + qv[d + 3] * f1.y + qv[d + 4] * f2.x +
             qv[d + 5] * f2.y + qv[d + 6] * f3.x + qv[d + 7] * f3.y;
    }
    acc *= scale;
    prow[k] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int k = lane; k < N; k += 32) {
    const float e = __expf(prow[k] - mx);
    prow[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int k = lane; k < N; k += 32) prow[k] *= inv;
}

}  // namespace

// tcgen05 / TMEM kernels for sequences of up to 256 tokens (attention_tc.cu)
constexpr int TC_MAX_TOKENS = 256;      // eval-mode maps (one S tile)
constexpr int TC_MAX_TOKENS_FWD = 1 << 30;  // forward: one S tile up to 256 tokens, key tiles beyond (384x384 images: 577; patch 8: 785 / 1025)
constexpr int TC_MAX_TOKENS_BWD = 240;  // backward (shared-memory budget of the pipelined kernel)
int attention_fwd_tc(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, int q_rows, bool fp16, const DropSpec* drop,
                     cudaStream_t st);
int attention_probs_tc(const void* qkv, const float* lse, float* probs, long long batch_stride, int B, int N, int H, float scale,
                       bool fp16, cudaStream_t st);
int attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int B, int N,
                     int H, float scale, int q_rows, bool fp16, const DropSpec* drop, cudaStream_t st);
int attention_bwd_tc_long(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, int B, int N, int H,
                          float scale, int q_rows, bool fp16, const DropSpec* drop, cudaStream_t st);
}  // namespace vitk

using namespace vitk;

template <bool H16>
static int attention_probs_impl(const void* qkv, const float* lse, float* probs, long long batch_stride, int B, int N, int H,
                                float scale, cudaStream_t st) {
  if (N <= TC_MAX_TOKENS) return attention_probs_tc(qkv, lse, probs, batch_stride, B, N, H, scale, H16, st);   // tensor-core S + lse
  const long long rows = (long long)B * H * N;
  attn_probs_kernel<H16><<<(unsigned)((rows + 3) / 4), 128, 0, st>>>(reinterpret_cast<const bf16*>(qkv), probs, batch_stride, B, N, H,
                                                                     scale);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

template <bool H16>
static int attention_fwd_impl(const void* qkv, void* out, float* lse, float* probs, int B, int N, int H, float scale, int q_rows,
                              cudaStream_t st) {
  if (N <= TC_MAX_TOKENS_FWD) {
    const int rc = attention_fwd_tc(qkv, out, lse, B, N, H, scale, probs != nullptr ? 0 : q_rows, H16, nullptr, st);  // maps need every lse
    if (rc != VITK_OK) return rc;
  } else {
    dim3 grid((N + TILE - 1) / TILE, H, B);
    attn_fwd_kernel<H16, false><<<grid, 128, 0, st>>>(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), lse, N, H,
                                                      scale * LOG2E, make_drop_spec(nullptr, 0.f, 0), 0);
    VITK_LAUNCH_CHECK();
  }
  if (probs != nullptr) return attention_probs_impl<H16>(qkv, lse, probs, (long long)H * N * N, B, N, H, scale, st);
  return VITK_OK;
}

extern "C" int vitk_attention_fwd(const void* qkv, void* out, int32_t dtype, float* lse, float* probs, int32_t B, int32_t N,
                                  int32_t H, float scale, int32_t q_rows, void* stream) {
  VITK_CHECK_ARG(q_rows >= 0, "vitk_attention_fwd: q_rows must be >= 0 (0: every query row)");
  VITK_CHECK_ARG(qkv && out && lse, "vitk_attention_fwd: null pointer");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_attention_fwd: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && N > 0 && H > 0, "vitk_attention_fwd: bad shape B=%d N=%d H=%d", B, N, H);
  VITK_CHECK_ARG(H <= 65535 && B <= 65535, "vitk_attention_fwd: grid limit");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == VITK_FP16 ? attention_fwd_impl<true>(qkv, out, lse, probs, B, N, H, scale, q_rows, st)
                            : attention_fwd_impl<false>(qkv, out, lse, probs, B, N, H, scale, q_rows, st);
}

extern "C" int vitk_attention_probs(const void* qkv, int32_t dtype, const float* lse, float* probs, int64_t probs_batch_stride,
                                    int32_t B, int32_t N, int32_t H, float scale, void* stream) {
  VITK_CHECK_ARG(qkv && lse && probs, "vitk_attention_probs: null pointer");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_attention_probs: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "vitk_attention_probs: bad shape B=%d N=%d H=%d", B, N, H);
  VITK_CHECK_ARG(probs_batch_stride >= (int64_t)H * N * N, "vitk_attention_probs: batch stride smaller than one image's maps");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == VITK_FP16 ? attention_probs_impl<true>(qkv, lse, probs, probs_batch_stride, B, N, H, scale, st)
                            : attention_probs_impl<false>(qkv, lse, probs, probs_batch_stride, B, N, H, scale, st);
}

static bool tc_long_bwd_enabled() {     // VITK_ATTN_LEGACY_BWD=1: the mma.sync kernels (A/B measurements only)
  static const bool on = [] {
    const char* e = getenv("VITK_ATTN_LEGACY_BWD");
    return !(e != nullptr && e[0] == '1');
  }();
  return on;
}

template <bool H16, bool DROP>
static int attention_bwd_impl(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                              int B, int N, int H, float scale, DropSpec drop, cudaStream_t st, int q_rows = 0) {
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<H16, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    VITK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<H16, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    configured = true;
  }
  const int Npad = (N + 7) & ~7;
  if (N <= TC_MAX_TOKENS_BWD) return attention_bwd_tc(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, q_rows, H16, DROP ? &drop : nullptr, st);  // delta fused
  const long long rows = (long long)B * N * H;
  attn_delta_kernel<H16><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(reinterpret_cast<const bf16*>(out),
                                                                     reinterpret_cast<const bf16*>(dout), delta, B, N, H);
  VITK_LAUNCH_CHECK();
  if (tc_long_bwd_enabled())
    return attention_bwd_tc_long(qkv, dout, lse, delta, dqkv, B, N, H, scale, q_rows, H16, DROP ? &drop : nullptr, st);
  dim3 grid((N + TILE - 1) / TILE, H, B);
  attn_bwd_dkdv_kernel<H16, DROP><<<grid, 128, sizeof(BwdSmem), st>>>(reinterpret_cast<const bf16*>(qkv),
                                                                      reinterpret_cast<const bf16*>(dout), lse, delta,
                                                                      reinterpret_cast<bf16*>(dqkv), N, H, scale, scale * LOG2E, drop, Npad);
  VITK_LAUNCH_CHECK();
  attn_bwd_dq_kernel<H16, DROP><<<grid, 128, sizeof(BwdSmem), st>>>(reinterpret_cast<const bf16*>(qkv),
                                                                    reinterpret_cast<const bf16*>(dout), lse, delta,
                                                                    reinterpret_cast<bf16*>(dqkv), N, H, scale, scale * LOG2E, drop, Npad);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                                  void* dqkv, int32_t dtype, int32_t B, int32_t N, int32_t H, float scale, int32_t q_rows,
                                  void* stream) {
  VITK_CHECK_ARG(qkv && out && dout && lse && delta && dqkv, "vitk_attention_bwd: null pointer");
  VITK_CHECK_ARG(q_rows >= 0, "vitk_attention_bwd: q_rows must be >= 0 (0: every query row)");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_attention_bwd: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && N > 0 && H > 0, "vitk_attention_bwd: bad shape B=%d N=%d H=%d", B, N, H);
  VITK_CHECK_ARG(H <= 65535 && B <= 65535, "vitk_attention_bwd: grid limit");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const DropSpec none = make_drop_spec(nullptr, 0.f, 0);
  return dtype == VITK_FP16 ? attention_bwd_impl<true, false>(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, none, st, q_rows)
                            : attention_bwd_impl<false, false>(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, none, st, q_rows);
}

// ---- training-mode dropout on the attention probabilities (Attention.attn_drop, vision_transformer_base.py:184).  Every
// ViT / DeiT configuration of the reference sets the rate to 0.  Forward: tcgen05 kernels up to 816 tokens; backward: the tcgen05
// kernel up to 240 tokens (DROP instantiation), the mma.sync kernels beyond.
extern "C" int vitk_attention_dropout_fwd(const void* qkv, void* out, int32_t dtype, float* lse, int32_t B, int32_t N, int32_t H,
                                          float scale, const vitk_dropout* attn_drop, void* stream) {
  VITK_CHECK_ARG(qkv && out && lse && attn_drop && attn_drop->seed, "vitk_attention_dropout_fwd: null pointer");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_attention_dropout_fwd: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "vitk_attention_dropout_fwd: bad shape B=%d N=%d H=%d", B, N, H);
  VITK_CHECK_ARG(attn_drop->p > 0.f && attn_drop->p < 1.f, "vitk_attention_dropout_fwd: need 0 < p < 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const DropSpec ds = make_drop_spec(attn_drop->seed, attn_drop->p, attn_drop->site);
  if (N <= TC_MAX_TOKENS_FWD) return attention_fwd_tc(qkv, out, lse, B, N, H, scale, 0, dtype == VITK_FP16, &ds, st);
  const int Npad = (N + 7) & ~7;
  dim3 grid((N + TILE - 1) / TILE, H, B);
  if (dtype == VITK_FP16)
    attn_fwd_kernel<true, true><<<grid, 128, 0, st>>>(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), lse, N, H,
                                                      scale * LOG2E, ds, Npad);
  else
    attn_fwd_kernel<false, true><<<grid, 128, 0, st>>>(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), lse, N, H,
                                                       scale * LOG2E, ds, Npad);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_attention_dropout_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                                          void* dqkv, int32_t dtype, int32_t B, int32_t N, int32_t H, float scale,
                                          const vitk_dropout* attn_drop, void* stream) {
  VITK_CHECK_ARG(qkv && out && dout && lse && delta && dqkv && attn_drop && attn_drop->seed, "vitk_attention_dropout_bwd: null pointer");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_attention_dropout_bwd: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "vitk_attention_dropout_bwd: bad shape B=%d N=%d H=%d", B, N, H);
  VITK_CHECK_ARG(attn_drop->p > 0.f && attn_drop->p < 1.f, "vitk_attention_dropout_bwd: need 0 < p < 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const DropSpec ds = make_drop_spec(attn_drop->seed, attn_drop->p, attn_drop->site);
  return dtype == VITK_FP16 ? attention_bwd_impl<true, true>(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, ds, st)
                            : attention_bwd_impl<false, true>(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, ds, st);
}
