// attention_tc.cu -- fused softmax attention on the 5th-gen tensor cores (tcgen05.mma, S / P / O and every backward
// accumulator resident in TMEM, operands staged by TMA), head_dim 64.  The first two kernels serve sequences of up to 256 / 240
// tokens (every 224x224 / 256x256 configuration of the reference with N = 197 / 198 / 257 at patch 16); the key-tile forward
// and the two streaming backward kernels further down serve every longer sequence (384x384: 577; patch 8: 785 / 1025 / 2305).
//
// Replaces Attention.forward's q@k^T*scale -> softmax -> attn@v (vision_transformer_base.py:182-191) and its autograd
// backward.  Layouts are the reference's own: qkv [B,N,3,H,64] as nn.Linear(D,3D) emits it (:178), out [B,N,H,64] =
// (attn@v).transpose(1,2).reshape(B,N,C) (:191).  The [B,H,N,N] score tensor never exists in HBM.
//
// Forward  (one CTA per (128-query tile, head, image); 4 softmax warps + 1 TMA/MMA warp; 256 TMEM columns, 2 CTAs/SM)
//   S[128 x KP] = Q K^T            tcgen05.mma SS, fp32 in TMEM columns [0, KP)
//   P = exp2(S*c - max*c)          one thread per query row (tcgen05.ld), written back IN PLACE as 16-bit (tcgen05.st)
//   O[128 x 64] = P V              tcgen05.mma TS (A = P from TMEM, B = V MN-major from smem), TMEM columns [128, 192)
//   out = O / rowsum               TMEM -> registers -> swizzled smem -> TMA store (rows >= N clipped by the tensor map)
//
// Backward (one CTA per (head, image); 8 math warps + 1 TMA/MMA warp; all 512 TMEM columns; N <= 240).  For every
// 128-key tile j and every 64-query chunk c (software-pipelined: the MMAs of chunk t run while chunk t+1 is computed):
//   S^T = K_j Q_c^T, dP^T = V_j dO_c^T             (SS)  columns [0,64), [64,128); lane = key, column = query
//   P^T = exp2(S^T*c - lse), dS^T = P^T o (dP^T - delta)   16-bit, double-buffered in columns [128,256); dS is also
//                                                   staged MN-major in smem (A operand of dQ); delta = rowsum(dO o O)
//                                                   is computed in the kernel from the TMA-staged dO and O tiles
//   dV_j += P^T dO_c, dK_j += dS^T Q_c             (TS)  columns [256,320), [320,384)
//   dQ_i += dS K_j  (128-query tile i = 2 chunks)  (SS)  columns [384,448) / [448,512), accumulated over j
// so K, V, Q, dO, O are read from HBM exactly once and nothing is recomputed or reduced through global memory.
#include "tc_common.cuh"

#ifdef VITK_GEMM_KNOBS
// profiling build only: SM clock stamps of one CTA of a later wave (head 0, image 100) -- [0..63] MMA thread, [64..] math warp 0
__device__ long long g_attn_dbg[256];
#define ASTAMP(i)                                                                   \
  do {                                                                              \
    if (blockIdx.x == 0 && blockIdx.y == (gridDim.y > 100 ? 100 : 0) && (i) < 256) g_attn_dbg[i] = clock64(); \
  } while (0)
#define FSTAMP(i)                                                                                                   \
  do {                                                                                                              \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == (gridDim.z > 100 ? 100 : 0)) g_attn_dbg[200 + (i)] = clock64(); \
  } while (0)
#else
#define FSTAMP(i) \
  do {            \
  } while (0)
#define ASTAMP(i) \
  do {            \
  } while (0)
#endif

namespace vitk {
namespace {

using namespace tc;

constexpr float LOG2E = 1.4426950408889634f;
constexpr int DH = 64;
constexpr int O_COL = 128;  // forward: TMEM column of the O accumulator (P occupies [0, KP/2) <= 128)

template <bool H16>
__device__ __forceinline__ uint32_t pk16(float lo, float hi) {
  if (H16) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16(lo, hi);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ================================================================================================ forward
constexpr int FWD_THREADS = 160;

// DROP: nn.Dropout on the softmax output (Attention.attn_drop, vision_transformer_base.py:184): O = (P o M) V with the
// counter-based factors M[b,h,q,key] = 0 | 1/(1-p) of element (((b*H + h)*N + q)*Npad + key) at `drop.site` (the same function
// vitk_dropout_mask exports); the row sum and lse stay those of the unmasked P.
template <bool H16, bool DROP>
__global__ void __launch_bounds__(FWD_THREADS, 2)
    attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                       const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse, int N, int H, int KP, float scale,
                       float scale_log2, int wave_ctas, DropSpec drop, int Npad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                       // [128][64] 16-bit, later the output staging tile
  uint8_t* sK = sQ + 128 * 128;             // [KP][64]
  uint8_t* sV = sK + KP * 128;              // [KP][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KP * 128);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (warp == 4) {
    if (elect_one()) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmKV);
      prefetch_tmap(&tmO);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      mbar_init_fence();
      pdl_wait();  // qkv is the previous kernel's output
      mbar_expect_tx(bar_qk, (128 + KP) * 128);
      tma_load_3d(sQ, &tmQ, bar_qk, h * DH, q0, b);
      tma_load_3d(sK, &tmKV, bar_qk, (H + h) * DH, 0, b);
      mbar_expect_tx(bar_v, KP * 128);
      tma_load_3d(sV, &tmKV, bar_v, (2 * H + h) * DH, 0, b);
      // operands of the CTA that runs one wave later -> L2
      const long long nxt = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x + wave_ctas;
      if (nxt < (long long)gridDim.x * gridDim.y * gridDim.z) {
        const int nq = int(nxt % gridDim.x), nh = int((nxt / gridDim.x) % gridDim.y), nb = int(nxt / ((long long)gridDim.x * gridDim.y));
        tma_prefetch_3d(&tmQ, nh * DH, nq * 128, nb);
        tma_prefetch_3d(&tmKV, (H + nh) * DH, 0, nb);
        tma_prefetch_3d(&tmKV, (2 * H + nh) * DH, 0, nb);
      }
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  pdl_trigger();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  if (warp == 4) {
    // ===================== MMA issuer (one elected thread) =====================
    if (elect_one()) {
      FSTAMP(0);
      mbar_wait(bar_qk, 0, 1);
      tc_fence_after();
      FSTAMP(1);
      const uint32_t idesc_s = idesc_f16(KP, false, false, H16);
      const uint64_t adesc = smem_desc_kmajor(smem_u32(sQ));
      const uint64_t bdesc = smem_desc_kmajor(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tb, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
      mbar_wait(bar_p, 0, 2);
      mbar_wait(bar_v, 0, 3);
      tc_fence_after();
      FSTAMP(2);
      const uint32_t idesc_o = idesc_f16(DH, false, true, H16);
      const uint64_t vdesc = smem_desc_mnmajor(smem_u32(sV), 8192);
      const int steps = KP >> 4;
      for (int s = 0; s < steps; ++s)
        umma_ts(tb + O_COL, tb + uint32_t(8 * s), vdesc + uint64_t(s * (2048 >> 4)), idesc_o, s > 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue: one thread per query row =====================
    const int row = warp * 32 + lane;
    const int q = q0 + row;
    const uint32_t trow = tb + (uint32_t(warp * 32) << 16);
    // a warp whose 32 query rows are all padding (second tile of a 197/198-token sequence) only keeps the barriers moving;
    // its P / O rows are garbage that no valid row depends on and the TMA store clips
    const bool warp_valid = q0 + warp * 32 < N;
    mbar_wait(bar_s, 0, 4);
    tc_fence_after();
    if (threadIdx.x == 0) FSTAMP(3);
    // pass 1: row maximum over the valid keys.  Four independent running maxima (a single fmaxf chain over 208 columns is
    // ~1000 cycles of pure dependency latency) and the next chunk's TMEM load in flight while this one is reduced.
    float mx;
    {
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      uint32_t va[32], vb[32];
      const int nch = (KP + 31) >> 5;
      auto load_chunk = [&](int c, uint32_t (&v)[32]) {
        if (KP - 32 * c >= 32) tmem_ld32_nowait(trow + uint32_t(32 * c), v);
        else tmem_ld16_nowait(trow + uint32_t(32 * c), v);
      };
      auto reduce_chunk = [&](int c, const uint32_t (&v)[32]) {
        const int c0 = 32 * c;
        if (c0 + 32 <= N) {        // whole chunk valid: no per-element masking
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], (c0 + j < N) ? __uint_as_float(v[j]) : -INFINITY);
        }
      };
      if (warp_valid) {
        load_chunk(0, va);
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait();
          if (c + 1 < nch) load_chunk(c + 1, vb);
          reduce_chunk(c, va);
          if (c + 1 < nch) {
            tmem_ld_wait();
            if (c + 2 < nch) load_chunk(c + 2, va);
            reduce_chunk(c + 1, vb);
          }
        }
      }
      mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    }
    // pass 2: P = exp2((S - max) * scale * log2e), 16-bit, in place
    if (threadIdx.x == 0) FSTAMP(4);
    const float msc = mx * scale_log2;
    float2 sum2 = make_float2(0.f, 0.f), sum2b = make_float2(0.f, 0.f);
    // DROP: counter block of this query row's first 8 keys (Npad % 8 == 0); block + k covers keys [8k, 8k + 8)
    const unsigned long long dseed = DROP ? __ldg(drop.seed) : 0ull;
    const unsigned long long rowblk = DROP ? ((((unsigned long long)b * H + h) * N + (q < N ? q : 0)) * (unsigned long long)Npad) >> 3 : 0ull;
    for (int c0 = 0; warp_valid && c0 < KP; c0 += 32) {
      uint32_t v[32];
      const bool full = KP - c0 >= 32;
      if (full) {
        tmem_ld32_nowait(trow + uint32_t(c0), v);
      } else {
        tmem_ld16_nowait(trow + uint32_t(c0), v);
#pragma unroll
        for (int j = 16; j < 32; ++j) v[j] = 0u;
      }
      uint32_t dw[16];   // DROP: the 16-bit uniforms of the chunk's 32 keys, two per word
      if constexpr (DROP) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint4 bits = drop_bits8(dseed, drop.site, rowblk + (unsigned long long)((c0 >> 3) + t));
          dw[4 * t + 0] = bits.x;
          dw[4 * t + 1] = bits.y;
          dw[4 * t + 2] = bits.z;
          dw[4 * t + 3] = bits.w;
        }
      }
      tmem_ld_wait();
      uint32_t ph[16];
      const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-msc, -msc);
      if (c0 + 32 <= N) {          // whole chunk valid
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, nm2);
          const float2 pp = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          if (j & 2) sum2b = __fadd2_rn(sum2b, pp);
          else sum2 = __fadd2_rn(sum2, pp);
          if constexpr (DROP) {
            const float2 f = drop_pair(dw[j >> 1], drop.thresh, drop.inv_keep);
            ph[j >> 1] = pk16<H16>(pp.x * f.x, pp.y * f.y);
          } else {
            ph[j >> 1] = pk16<H16>(pp.x, pp.y);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, nm2);
          const float p0 = (c0 + j < N) ? ex2_approx(x.x) : 0.f;
          const float p1 = (c0 + j + 1 < N) ? ex2_approx(x.y) : 0.f;
          sum2 = __fadd2_rn(sum2, make_float2(p0, p1));
          if constexpr (DROP) {
            const float2 f = drop_pair(dw[j >> 1], drop.thresh, drop.inv_keep);
            ph[j >> 1] = pk16<H16>(p0 * f.x, p1 * f.y);
          } else {
            ph[j >> 1] = pk16<H16>(p0, p1);
          }
        }
      }
      if (full) tmem_st16_nowait(trow + uint32_t(c0 >> 1), ph);
      else tmem_st8_nowait(trow + uint32_t(c0 >> 1), ph);
    }
    const float sum = (sum2.x + sum2.y) + (sum2b.x + sum2b.y);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar_p);
    if (threadIdx.x == 0) FSTAMP(5);
    if (q < N) lse[((long long)b * H + h) * N + q] = fmaf(mx, scale, logf(sum));
    const float inv = 1.f / sum;
    // epilogue: O / rowsum -> 16-bit -> swizzled staging (the Q tile is dead once S exists) -> TMA store
    mbar_wait(bar_o, 0, 5);
    tc_fence_after();
    if (threadIdx.x == 0) FSTAMP(6);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      if (!warp_valid) break;
      uint32_t v[32];
      tmem_ld32_nowait(trow + uint32_t(O_COL + 32 * hh), v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 o;
        o.x = pk16<H16>(__uint_as_float(v[8 * c + 0]) * inv, __uint_as_float(v[8 * c + 1]) * inv);
        o.y = pk16<H16>(__uint_as_float(v[8 * c + 2]) * inv, __uint_as_float(v[8 * c + 3]) * inv);
        o.z = pk16<H16>(__uint_as_float(v[8 * c + 4]) * inv, __uint_as_float(v[8 * c + 5]) * inv);
        o.w = pk16<H16>(__uint_as_float(v[8 * c + 6]) * inv, __uint_as_float(v[8 * c + 7]) * inv);
        *reinterpret_cast<uint4*>(sQ + swz128(row, 4 * hh + c)) = o;
      }
    }
    fence_proxy_async();
    named_bar_sync(1, 128);
    if (warp == 0 && elect_one()) {
      tma_store_3d(&tmO, sQ, h * DH, q0, b);
      tma_store_commit();
      tma_store_wait_read();   // shared memory must outlive the reads; the writes complete asynchronously
    }
    if (threadIdx.x == 0) FSTAMP(7);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<256>(tb);
}

// ================================================================================================ forward, > 256 tokens
// 384x384 images (577 tokens), patch 8 (785 / 1025 / 2305 tokens): the keys no longer fit one S tile, so they are walked as tiles
// of KT <= 256 keys, in groups of TPG <= 4 tiles whose K and V sit in shared memory together (one CTA per SM; up to 816 tokens
// that is ONE group, i.e. K and V are loaded once, exactly like the short kernel).  S_j reuses TMEM columns [0, KT); every tile of
// a group gets ITS OWN 64-column O accumulator (columns 256 + 64 j) with its own row maximum m_j and row sum l_j, so no
// accumulator is ever rescaled in TMEM: at the end of a group the softmax threads merge the partial results into a
// register-resident running (m, L, O[64]) -- the split-KV form of the online softmax,
//   m' = max(m, max_j m_j),  O' = O 2^((m - m') c) + sum_j 2^((m_j - m') c) O_j,  L' likewise,  out = O / L,  lse = m scale + log L.
// The issuer starts S_{j+1} only after P_j V_j has completed (P_j lives in S's columns) and the next group's loads only after the
// group's last MMA; the group's first P V waits until the previous group's accumulators have been read (bar_free).
// TCOLS = 512: S tile of up to 256 keys + up to 4 accumulators, one CTA per SM; TCOLS = 256: S tile of up to 128 keys + 2
// accumulators, two CTAs per SM (one CTA's softmax runs under the other's MMAs and loads).
constexpr int FWDL_MAX_TILES = 4;

template <bool H16, bool DROP, int TCOLS>
__global__ void __launch_bounds__(FWD_THREADS, TCOLS == 256 ? 2 : 1)
    attn_fwd_tc_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                            const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse, int N, int H, int KT, int NKT, int TPG,
                            float scale, float scale_log2, DropSpec drop, int Npad) {
  constexpr int OL_COL = TCOLS / 2;   // first O accumulator
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int tile_bytes = KT * 128;
  uint8_t* sQ = smem;                            // [128][64] 16-bit, later the output staging tile
  uint8_t* sK = sQ + 128 * 128;                  // TPG x [KT][64]
  uint8_t* sV = sK + TPG * tile_bytes;           // TPG x [KT][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + TPG * tile_bytes);
  uint64_t* bar_qk = bars + 0;                   // phase = group
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;                    // [4] S_j complete
  uint64_t* bar_p = bars + 6;                    // [4] P_j written (128 arrivals)
  uint64_t* bar_o = bars + 10;                   // [4] P_j V_j complete
  uint64_t* bar_free = bars + 14;                // the group's O accumulators have been read (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int NG = (NKT + TPG - 1) / TPG;

  if (warp == 4) {
    if (elect_one()) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmKV);
      prefetch_tmap(&tmO);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      for (int j = 0; j < FWDL_MAX_TILES; ++j) {
        mbar_init(bar_s + j, 1);
        mbar_init(bar_p + j, 128);
        mbar_init(bar_o + j, 1);
      }
      mbar_init(bar_free, 128);
      mbar_init_fence();
      pdl_wait();  // qkv is the previous kernel's output
      const int nt0 = min(TPG, NKT);
      mbar_expect_tx(bar_qk, 128 * 128 + nt0 * tile_bytes);
      tma_load_3d(sQ, &tmQ, bar_qk, h * DH, q0, b);
      for (int j = 0; j < nt0; ++j) tma_load_3d(sK + j * tile_bytes, &tmKV, bar_qk, (H + h) * DH, j * KT, b);
      mbar_expect_tx(bar_v, nt0 * tile_bytes);
      for (int j = 0; j < nt0; ++j) tma_load_3d(sV + j * tile_bytes, &tmKV, bar_v, (2 * H + h) * DH, j * KT, b);
    }
    __syncwarp();
    tmem_alloc<TCOLS>(tmem_slot);
  }
  pdl_trigger();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  if (warp == 4) {
    // ===================== MMA issuer (one elected thread) =====================
    if (elect_one()) {
      const uint32_t idesc_s = idesc_f16(KT, false, false, H16);
      const uint32_t idesc_o = idesc_f16(DH, false, true, H16);
      const uint64_t adesc = smem_desc_kmajor(smem_u32(sQ));
      const int steps = KT >> 4;
      for (int g = 0; g < NG; ++g) {
        const uint32_t par = uint32_t(g & 1);
        const int ntg = min(TPG, NKT - g * TPG);
        mbar_wait(bar_qk, par, 1);
        tc_fence_after();
        for (int j = 0; j < ntg; ++j) {
          if (j > 0) {                             // P_{j-1} (in S's columns) has been consumed
            mbar_wait(bar_o + (j - 1), par, 2);
            tc_fence_after();
          }
          const uint64_t bdesc = smem_desc_kmajor(smem_u32(sK + j * tile_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tb, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(bar_s + j);
          mbar_wait(bar_p + j, par, 3);
          if (j == 0) {
            mbar_wait(bar_v, par, 4);
            if (g > 0) mbar_wait(bar_free, uint32_t((g - 1) & 1), 7);   // the previous group's accumulators have been merged
          }
          tc_fence_after();
          const uint64_t vdesc = smem_desc_mnmajor(smem_u32(sV + j * tile_bytes), 8192);
          for (int st = 0; st < steps; ++st)
            umma_ts(tb + uint32_t(OL_COL + 64 * j), tb + uint32_t(8 * st), vdesc + uint64_t(st * (2048 >> 4)), idesc_o, st > 0 ? 1u : 0u);
          umma_commit(bar_o + j);
        }
        if (g + 1 < NG) {                          // K / V of this group are dead once its last MMA has completed
          mbar_wait(bar_o + (ntg - 1), par, 8);
          tc_fence_after();
          const int t0 = (g + 1) * TPG;
          const int ntn = min(TPG, NKT - t0);
          mbar_expect_tx(bar_qk, ntn * tile_bytes);
          for (int j = 0; j < ntn; ++j) tma_load_3d(sK + j * tile_bytes, &tmKV, bar_qk, (H + h) * DH, (t0 + j) * KT, b);
          mbar_expect_tx(bar_v, ntn * tile_bytes);
          for (int j = 0; j < ntn; ++j) tma_load_3d(sV + j * tile_bytes, &tmKV, bar_v, (2 * H + h) * DH, (t0 + j) * KT, b);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue: one thread per query row =====================
    const int row = warp * 32 + lane;
    const int q = q0 + row;
    const uint32_t trow = tb + (uint32_t(warp * 32) << 16);
    const bool warp_valid = q0 + warp * 32 < N;
    const unsigned long long dseed = DROP ? __ldg(drop.seed) : 0ull;
    const unsigned long long rowblk = DROP ? ((((unsigned long long)b * H + h) * N + (q < N ? q : 0)) * (unsigned long long)Npad) >> 3 : 0ull;
    float m_run = -INFINITY, L_run = 0.f;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    for (int g = 0; g < NG; ++g) {
      const uint32_t par = uint32_t(g & 1);
      const int ntg = min(TPG, NKT - g * TPG);
      float mj[FWDL_MAX_TILES], lj[FWDL_MAX_TILES];
#pragma unroll
      for (int j = 0; j < FWDL_MAX_TILES; ++j) {
        mj[j] = -INFINITY;
        lj[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < FWDL_MAX_TILES; ++j) {
        if (j < ntg) {
          const int key0 = (g * TPG + j) * KT;
          const int nv = min(KT, N - key0);          // valid keys of this tile (>= 1)
          mbar_wait(bar_s + j, par, 5);
          tc_fence_after();
          if (warp_valid) {
            // pass 1: row maximum over the tile's valid keys
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            for (int c0 = 0; c0 < nv; c0 += 32) {
              uint32_t v[32];
              if (KT - c0 >= 32) {
                tmem_ld32_nowait(trow + uint32_t(c0), v);
              } else {
                tmem_ld16_nowait(trow + uint32_t(c0), v);
#pragma unroll
                for (int i = 16; i < 32; ++i) v[i] = 0u;
              }
              tmem_ld_wait();
              if (c0 + 32 <= nv) {
#pragma unroll
                for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], (c0 + i < nv) ? __uint_as_float(v[i]) : -INFINITY);
              }
            }
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            // pass 2: P = exp2((S - max) * scale * log2e), 16-bit, in place; every column of the tile is written (0 beyond nv)
            const float msc = mx * scale_log2;
            const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-msc, -msc);
            float2 sum2 = make_float2(0.f, 0.f);
            for (int c0 = 0; c0 < KT; c0 += 32) {
              uint32_t v[32];
              const bool full = KT - c0 >= 32;
              if (full) {
                tmem_ld32_nowait(trow + uint32_t(c0), v);
              } else {
                tmem_ld16_nowait(trow + uint32_t(c0), v);
#pragma unroll
                for (int i = 16; i < 32; ++i) v[i] = 0u;
              }
              uint32_t dw[16];
              if constexpr (DROP) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const uint4 bits = drop_bits8(dseed, drop.site, rowblk + (unsigned long long)(((key0 + c0) >> 3) + t));
                  dw[4 * t + 0] = bits.x;
                  dw[4 * t + 1] = bits.y;
                  dw[4 * t + 2] = bits.z;
                  dw[4 * t + 3] = bits.w;
                }
              }
              tmem_ld_wait();
              uint32_t ph[16];
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2);
                const float p0 = (c0 + i < nv) ? ex2_approx(x.x) : 0.f;
                const float p1 = (c0 + i + 1 < nv) ? ex2_approx(x.y) : 0.f;
                sum2 = __fadd2_rn(sum2, make_float2(p0, p1));
                if constexpr (DROP) {
                  const float2 f = drop_pair(dw[i >> 1], drop.thresh, drop.inv_keep);
                  ph[i >> 1] = pk16<H16>(p0 * f.x, p1 * f.y);
                } else {
                  ph[i >> 1] = pk16<H16>(p0, p1);
                }
              }
              if (full) tmem_st16_nowait(trow + uint32_t(c0 >> 1), ph);
              else tmem_st8_nowait(trow + uint32_t(c0 >> 1), ph);
            }
            mj[j] = mx;
            lj[j] = sum2.x + sum2.y;
            tmem_st_wait();
          }
          tc_fence_before();
          mbar_arrive(bar_p + j);
        }
      }
      // merge the group's partial results into the running (m, L, O)
      float m_new = m_run;
#pragma unroll
      for (int j = 0; j < FWDL_MAX_TILES; ++j) m_new = fmaxf(m_new, mj[j]);
      const float f_old = warp_valid ? ex2_approx((m_run - m_new) * scale_log2) : 0.f;   // 0 for the first group (m_run = -inf)
      float wj[FWDL_MAX_TILES];
      L_run *= f_old;
#pragma unroll
      for (int j = 0; j < FWDL_MAX_TILES; ++j) {
        wj[j] = (j < ntg && warp_valid) ? ex2_approx((mj[j] - m_new) * scale_log2) : 0.f;
        L_run = fmaf(wj[j], lj[j], L_run);
      }
      m_run = m_new;
      mbar_wait(bar_o + (ntg - 1), par, 6);          // commits are ordered: every earlier P_j V_j of the group is complete as well
      tc_fence_after();
      if (warp_valid) {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] *= f_old;
#pragma unroll
        for (int j = 0; j < FWDL_MAX_TILES; ++j) {
          if (j < ntg) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[32];
              tmem_ld32_nowait(trow + uint32_t(OL_COL + 64 * j + 32 * hh), v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) acc[32 * hh + i] = fmaf(__uint_as_float(v[i]), wj[j], acc[32 * hh + i]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_free);
    }
    if (q < N) lse[((long long)b * H + h) * N + q] = fmaf(m_run, scale, logf(L_run));
    const float inv = 1.f / L_run;
    if (warp_valid) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 o;
        o.x = pk16<H16>(acc[8 * c + 0] * inv, acc[8 * c + 1] * inv);
        o.y = pk16<H16>(acc[8 * c + 2] * inv, acc[8 * c + 3] * inv);
        o.z = pk16<H16>(acc[8 * c + 4] * inv, acc[8 * c + 5] * inv);
        o.w = pk16<H16>(acc[8 * c + 6] * inv, acc[8 * c + 7] * inv);
        *reinterpret_cast<uint4*>(sQ + swz128(row, c)) = o;
      }
    }
    fence_proxy_async();
    named_bar_sync(1, 128);
    if (warp == 0 && elect_one()) {
      tma_store_3d(&tmO, sQ, h * DH, q0, b);
      tma_store_commit();
      tma_store_wait_read();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TCOLS>(tb);
}

// ================================================================================================ eval-mode attention maps
// probs[b,h,q,:] = softmax(q k^T * scale) in fp32 -- the `attention_maps` the reference keeps in eval mode
// (vision_transformer_base.py:186-188).  The forward above already left lse[b,h,q] = log sum_j exp(s_qj * scale), so
// the map is exp2(s * c - lse * log2e): one S = Q K^T on the tensor core (bit-identical to the forward's S), one pass over
// the TMEM row per thread, no max / sum / P / V.  HBM-bound on the B*H*N*N*4 bytes it writes (120 MB per DeiT-tiny layer at
// batch 256): each thread streams its row as 16-byte stores, two consecutive store instructions completing every 32-byte
// sector.  CTA per (128-query tile x key tile, head, image), 2 CTAs / SM like the forward; up to 256 tokens there is one key
// tile (KP = the padded sequence), longer sequences split the keys into tiles of KP <= 256 (blockIdx.x = key tile * NQT + query
// tile) -- every map element depends on its own (q, k) pair and the row's lse only, so the tiles are independent.
template <bool H16>
__global__ void __launch_bounds__(FWD_THREADS, 2)
    attn_probs_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         const float* __restrict__ lse, float* __restrict__ probs, long long batch_stride, int N, int H, int KP,
                         float scale_log2, int NQT) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                       // [128][64] 16-bit
  uint8_t* sK = sQ + 128 * 128;             // [KP][64]
  float* sT = reinterpret_cast<float*>(sK + KP * 128);           // 4 x [32][33] fp32 transpose tiles (one per warp)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + 4 * 32 * 33);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_s = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = (blockIdx.x % NQT) * 128, key0 = (blockIdx.x / NQT) * KP, h = blockIdx.y, b = blockIdx.z;

  if (warp == 4) {
    if (elect_one()) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmKV);
      mbar_init(bar_qk, 1);
      mbar_init(bar_s, 1);
      mbar_init_fence();
      pdl_wait();
      mbar_expect_tx(bar_qk, (128 + KP) * 128);
      tma_load_3d(sQ, &tmQ, bar_qk, h * DH, q0, b);
      tma_load_3d(sK, &tmKV, bar_qk, (H + h) * DH, key0, b);
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  pdl_trigger();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      mbar_wait(bar_qk, 0, 31);
      tc_fence_after();
      const uint32_t idesc_s = idesc_f16(KP, false, false, H16);
      const uint64_t adesc = smem_desc_kmajor(smem_u32(sQ));
      const uint64_t bdesc = smem_desc_kmajor(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tb, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
  } else {
    const int row = warp * 32 + lane;
    const int q = q0 + row;
    const uint32_t trow = tb + (uint32_t(warp * 32) << 16);
    const bool warp_valid = q0 + warp * 32 < N;
    const long long rowid = ((long long)b * H + h) * N + q;
    const float nl2 = q < N ? -lse[rowid] * LOG2E : 0.f;
    mbar_wait(bar_s, 0, 32);
    tc_fence_after();
    if (warp_valid) {
      const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(nl2, nl2);
      // A thread owns a row, but a row-per-lane store touches 32 different 128-byte lines per instruction (the LSU then
      // needs ~32 cycles for it: measured 168 us per 120 MB layer).  Each warp therefore transposes its 32 x 32 block through
      // a padded shared-memory tile and stores it row by row: one fully coalesced 128-byte segment per instruction.
      float* tile = sT + warp * (32 * 33);
      const int r0 = q0 + warp * 32;                       // first query row of this warp
      float* pbase = probs + (long long)b * batch_stride + ((long long)h * N + r0) * N;   // image b's maps start at b * batch_stride
      uint32_t va[32], vb[32];
      const int nch = (KP + 31) >> 5;
      auto load_chunk = [&](int c, uint32_t (&v)[32]) {
        if (KP - 32 * c >= 32) tmem_ld32_nowait(trow + uint32_t(32 * c), v);
        else tmem_ld16_nowait(trow + uint32_t(32 * c), v);
      };
      auto emit_chunk = [&](int c, const uint32_t (&v)[32]) {
        const int c0 = 32 * c;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, nm2);
          tile[lane * 33 + j] = ex2_approx(x.x);
          tile[lane * 33 + j + 1] = ex2_approx(x.y);
        }
        __syncwarp();
        const int col = key0 + c0 + lane;
        if (col < N && c0 + lane < KP) {          // the tile's own keys only (a 16-wide last chunk leaves lanes 16.. without data)
          const int nrows = min(32, N - r0);
          for (int rr = 0; rr < nrows; ++rr) pbase[(long long)rr * N + col] = tile[rr * 33 + lane];
        }
        __syncwarp();
      };
      load_chunk(0, va);
      for (int c = 0; c < nch; c += 2) {
        tmem_ld_wait();
        if (c + 1 < nch) load_chunk(c + 1, vb);
        emit_chunk(c, va);
        if (c + 1 < nch) {
          tmem_ld_wait();
          if (c + 2 < nch) load_chunk(c + 2, va);
          emit_chunk(c + 1, vb);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<256>(tb);
}

// ================================================================================================ backward
constexpr int BWD_THREADS = 288;  // 8 math warps + 1 control warp
// TMEM map (all 512 columns).  S^T / dP^T hold one 64-query chunk (fp32, lane = key); P^T / dS^T are their 16-bit
// versions (A operands of dV / dK), double-buffered so that the MMAs of chunk t run while chunk t+1 is being computed.
constexpr int TM_S = 0, TM_DP = 64, TM_PT = 128 /* +64*buf */, TM_DS = 160 /* +64*buf */, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

// Math of one chunk for one thread: W (32 or 16) query columns starting at column g0 of the chunk, key lane = this thread.
// DROP (attention-probability dropout, see the forward): with the factors m = M[q,key], dV += (P o M)^T dO and
// dS = P o (M o dP - delta) (delta = rowsum(dO o O) is unchanged: sum_k P_k M_k dP_k = dO . O).  The walk is transposed (this
// thread = one key, W queries), so every element needs its own counter block: dblk0 = block of (query qbase + g0, this key),
// dstep = blocks per query row (Npad / 8), dhalf = which 16-bit uniform of the block belongs to this key.
template <bool H16, int W, bool DROP>
__device__ __forceinline__ void bwd_math(const uint32_t (&sv)[32], const uint32_t (&dv)[32], uint32_t trow, int g0, int buf,
                                         const float* __restrict__ s_lse2, const float* __restrict__ s_delta, int qbase,
                                         float scale_log2, bool keyvalid, uint8_t* sDSblk, int keyrow, const DropSpec& drop,
                                         unsigned long long dseed, unsigned long long dblk0, unsigned long long dstep, int dhalf) {
  uint32_t ph[W / 2], dh[W / 2];
  const float2 sc2 = make_float2(scale_log2, scale_log2);
  auto dfac = [&](int i) -> float {   // dropout factor of query (qbase + g0 + i) x this key
    const uint4 bits = drop_bits8(dseed, drop.site, dblk0 + (unsigned long long)i * dstep);
    const uint32_t w = (dhalf >> 1) == 0 ? bits.x : (dhalf >> 1) == 1 ? bits.y : (dhalf >> 1) == 2 ? bits.z : bits.w;
    return (((dhalf & 1) ? (w >> 16) : (w & 0xffffu)) >= drop.thresh) ? drop.inv_keep : 0.f;
  };
#pragma unroll
  for (int e = 0; e < W; e += 4) {
    // s_lse2 / s_delta hold the NEGATED statistics (-lse*log2e, -delta): everything below is packed fp32x2 math
    const float4 l4 = *reinterpret_cast<const float4*>(s_lse2 + qbase + g0 + e);   // warp-uniform address: broadcast
    const float4 d4 = *reinterpret_cast<const float4*>(s_delta + qbase + g0 + e);
    const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(sv[e + 0]), __uint_as_float(sv[e + 1])), sc2, make_float2(l4.x, l4.y));
    const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])), sc2, make_float2(l4.z, l4.w));
    const float2 p01 = make_float2(ex2_approx(x01.x), ex2_approx(x01.y));
    const float2 p23 = make_float2(ex2_approx(x23.x), ex2_approx(x23.y));
    float2 t01, t23;
    if constexpr (DROP) {
      const float2 m01 = make_float2(dfac(e + 0), dfac(e + 1)), m23 = make_float2(dfac(e + 2), dfac(e + 3));
      t01 = __ffma2_rn(make_float2(__uint_as_float(dv[e + 0]), __uint_as_float(dv[e + 1])), m01, make_float2(d4.x, d4.y));
      t23 = __ffma2_rn(make_float2(__uint_as_float(dv[e + 2]), __uint_as_float(dv[e + 3])), m23, make_float2(d4.z, d4.w));
      const float2 q01 = __fmul2_rn(p01, m01), q23 = __fmul2_rn(p23, m23);
      ph[(e >> 1) + 0] = pk16<H16>(q01.x, q01.y);
      ph[(e >> 1) + 1] = pk16<H16>(q23.x, q23.y);
    } else {
      t01 = __fadd2_rn(make_float2(__uint_as_float(dv[e + 0]), __uint_as_float(dv[e + 1])), make_float2(d4.x, d4.y));
      t23 = __fadd2_rn(make_float2(__uint_as_float(dv[e + 2]), __uint_as_float(dv[e + 3])), make_float2(d4.z, d4.w));
      ph[(e >> 1) + 0] = pk16<H16>(p01.x, p01.y);
      ph[(e >> 1) + 1] = pk16<H16>(p23.x, p23.y);
    }
    const float2 s01 = __fmul2_rn(p01, t01), s23 = __fmul2_rn(p23, t23);
    dh[(e >> 1) + 0] = pk16<H16>(s01.x, s01.y);
    dh[(e >> 1) + 1] = pk16<H16>(s23.x, s23.y);
  }
  if (!keyvalid) {  // padded key lane (its S / dP rows are garbage, possibly non-finite): contributes exactly zero
#pragma unroll
    for (int i = 0; i < W / 2; ++i) ph[i] = dh[i] = 0u;
  }
  if constexpr (W == 32) {
    tmem_st16_nowait(trow + uint32_t(TM_PT + 64 * buf + (g0 >> 1)), ph);
    tmem_st16_nowait(trow + uint32_t(TM_DS + 64 * buf + (g0 >> 1)), dh);
  } else {
    tmem_st8_nowait(trow + uint32_t(TM_PT + 64 * buf + (g0 >> 1)), ph);
    tmem_st8_nowait(trow + uint32_t(TM_DS + 64 * buf + (g0 >> 1)), dh);
  }
  // dS, MN-major A operand of dQ = dS K: [128 key rows][64 queries = 128 B], 128B swizzle
#pragma unroll
  for (int ch = 0; ch < W / 8; ++ch)
    *reinterpret_cast<uint4*>(sDSblk + swz128(keyrow, (g0 >> 3) + ch)) = make_uint4(dh[4 * ch], dh[4 * ch + 1], dh[4 * ch + 2], dh[4 * ch + 3]);
}

template <bool H16, bool DROP>
__global__ void __launch_bounds__(BWD_THREADS, 1)
    attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmDQKV,
                       const float* __restrict__ lse, float* __restrict__ delta, int N, int H, int QP, float scale,
                       float scale_log2, int wave_ctas, int q_chunks, DropSpec drop, int Npad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int tile_bytes = QP * 128;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + tile_bytes;
  uint8_t* sV = sK + tile_bytes;
  uint8_t* sDO = sV + tile_bytes;
  uint8_t* sDS = sDO + tile_bytes;   // 2 buffers (one per 128-query tile in flight) x 2 chunks x [128][128 B]
  uint8_t* sOut = sDS + 65536;       // 2 x [128][128 B]: O at start-up, output staging afterwards
  float* s_lse2 = reinterpret_cast<float*>(sOut + 32768);  // [256]
  float* s_delta = s_lse2 + 256;                            // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_delta + 256);
  uint64_t* bar_kq = bars + 0;      // K, Q landed
  uint64_t* bar_vdo = bars + 1;     // V, dO landed
  uint64_t* bar_o = bars + 2;       // O landed
  uint64_t* bar_s = bars + 3;       // S^T, dP^T of chunk t complete
  uint64_t* bar_ld = bars + 4;      // every math thread has S^T, dP^T of chunk t in registers
  // [2] P^T, dS^T (TMEM + staging) of chunk t written, one barrier per buffer (t & 1), phase t >> 1.  A SINGLE barrier with
  // parity t & 1 deadlocks once in ~1e7 CTAs: S^T/dP^T of chunk t+1 are committed before the issuer consumes bar_p(t), so when
  // chunk t+1 is the short tail chunk (16 of 208 queries) the math warps can complete phase t+1 within a few hundred ns of
  // phase t, and an issuer that has not yet polled sees the parity of phase t+2 and waits for ever.  With one barrier per
  // buffer the next phase of the same barrier belongs to chunk t+2, whose operands are only issued after bar_p(t) was seen.
  uint64_t* bar_p = bars + 5;
  uint64_t* bar_m2 = bars + 7;      // [2] dV / dK / dQ MMAs that read P/dS buffer (t & 1) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int NJ = (N + 127) >> 7;    // 128-key tiles
  const int NCH_ALL = (QP + 63) >> 6;   // 64-query chunks
  // q_chunks > 0: dO is zero from query row 64 * q_chunks on (the caller's guarantee), so those chunks contribute nothing to dK /
  // dV and their dQ rows are zero: only the leading chunks are walked (the last block of a class-token model: 1 chunk of 4)
  const int NCH = q_chunks > 0 ? min(q_chunks, NCH_ALL) : NCH_ALL;
  const int T = NJ * NCH;

  if (warp == 8) {
    if (elect_one()) {
      prefetch_tmap(&tmQKV);
      prefetch_tmap(&tmDO);
      prefetch_tmap(&tmO);
      prefetch_tmap(&tmDQKV);
      mbar_init(bar_kq, 1);
      mbar_init(bar_vdo, 1);
      mbar_init(bar_o, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_ld, 256);
      mbar_init(bar_p + 0, 256);
      mbar_init(bar_p + 1, 256);
      mbar_init(bar_m2 + 0, 1);
      mbar_init(bar_m2 + 1, 1);
      mbar_init_fence();
      pdl_wait();  // qkv / dO / O are earlier kernels' outputs
      mbar_expect_tx(bar_kq, 2 * tile_bytes);
      tma_load_3d(sK, &tmQKV, bar_kq, (H + h) * DH, 0, b);
      tma_load_3d(sQ, &tmQKV, bar_kq, h * DH, 0, b);
      mbar_expect_tx(bar_vdo, 2 * tile_bytes);
      tma_load_3d(sV, &tmQKV, bar_vdo, (2 * H + h) * DH, 0, b);
      tma_load_3d(sDO, &tmDO, bar_vdo, h * DH, 0, b);
      mbar_expect_tx(bar_o, tile_bytes);            // O parks in the (still unused) output staging buffer
      tma_load_3d(sOut, &tmO, bar_o, h * DH, 0, b);
      // operands of the CTA that runs one wave later -> L2 (HBM streams while this wave computes)
      const int nxt = blockIdx.y * gridDim.x + blockIdx.x + wave_ctas;
      if (nxt < (int)(gridDim.x * gridDim.y)) {
        const int nh = nxt % gridDim.x, nb = nxt / gridDim.x;
        tma_prefetch_3d(&tmQKV, (H + nh) * DH, 0, nb);
        tma_prefetch_3d(&tmQKV, nh * DH, 0, nb);
        tma_prefetch_3d(&tmQKV, (2 * H + nh) * DH, 0, nb);
        tma_prefetch_3d(&tmDO, nh * DH, 0, nb);
        tma_prefetch_3d(&tmO, nh * DH, 0, nb);
      }
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
    pdl_wait();
  } else {
    pdl_wait();
    // -lse in the log2 domain (-inf masks the padded queries: P = exp2(-inf) = 0)
    const float* lse_b = lse + ((long long)b * H + h) * N;
    const int i = threadIdx.x;  // 0..255
    s_lse2[i] = i < N ? -lse_b[i] * LOG2E : -INFINITY;   // negated (see bwd_math)
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  if (warp == 8) {
    // ===================== MMA issuer (one elected thread) =====================
    if (elect_one()) {
      const uint32_t idesc64_ts = idesc_f16(DH, false, true, H16);   // A from TMEM (K-major), B MN-major
      const uint32_t idesc64_mn = idesc_f16(DH, true, true, H16);    // A, B MN-major from smem
      auto issue_s = [&](int j, int c) {      // S^T = K_j Q_c^T
        const int QC = min(64, QP - 64 * c);
        const uint32_t idesc_s = idesc_f16(QC, false, false, H16);
        const uint64_t kd = smem_desc_kmajor(smem_u32(sK + j * 16384));
        const uint64_t qd = smem_desc_kmajor(smem_u32(sQ + c * 8192));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tb + TM_S, kd + uint64_t(2 * k), qd + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
      };
      auto issue_dp = [&](int j, int c) {     // dP^T = V_j dO_c^T
        const int QC = min(64, QP - 64 * c);
        const uint32_t idesc_s = idesc_f16(QC, false, false, H16);
        const uint64_t vd = smem_desc_kmajor(smem_u32(sV + j * 16384));
        const uint64_t dd = smem_desc_kmajor(smem_u32(sDO + c * 8192));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tb + TM_DP, vd + uint64_t(2 * k), dd + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
      };
      ASTAMP(0);
      mbar_wait(bar_kq, 0, 1);
      tc_fence_after();
      ASTAMP(1);
      issue_s(0, 0);
      mbar_wait(bar_vdo, 0, 2);
      tc_fence_after();
      ASTAMP(2);
      issue_dp(0, 0);
      umma_commit(bar_s);
      int t = 0;
      for (int j = 0; j < NJ; ++j) {
        for (int c = 0; c < NCH; ++c, ++t) {
          // the math warps hold chunk t in registers: S^T / dP^T may be overwritten by chunk t+1 right away
          mbar_wait(bar_ld, t & 1, 3);
          tc_fence_after();
          ASTAMP(4 + 4 * t);
          if (t + 1 < T) {
            const int c1 = (c + 1 == NCH) ? 0 : c + 1;
            const int j1 = (c + 1 == NCH) ? j + 1 : j;
            issue_s(j1, c1);
            issue_dp(j1, c1);
            umma_commit(bar_s);
          }
          mbar_wait(bar_p + (t & 1), (t >> 1) & 1, 4);
          tc_fence_after();
          ASTAMP(5 + 4 * t);
          const int buf = t & 1;
          const int QC = min(64, QP - 64 * c);
          const int qsteps = QC >> 4;
          for (int s = 0; s < qsteps; ++s) {   // dV_j += P^T dO_c
            const uint64_t bd = smem_desc_mnmajor(smem_u32(sDO + (c * 64 + 16 * s) * 128), 8192);
            umma_ts(tb + TM_DV, tb + uint32_t(TM_PT + 64 * buf + 8 * s), bd, idesc64_ts, (c > 0 || s > 0) ? 1u : 0u);
          }
          for (int s = 0; s < qsteps; ++s) {   // dK_j += dS^T Q_c
            const uint64_t bd = smem_desc_mnmajor(smem_u32(sQ + (c * 64 + 16 * s) * 128), 8192);
            umma_ts(tb + TM_DK, tb + uint32_t(TM_DS + 64 * buf + 8 * s), bd, idesc64_ts, (c > 0 || s > 0) ? 1u : 0u);
          }
          if ((c & 1) || c == NCH - 1) {       // the 128-query tile c/2 is staged completely: dQ += dS K_j
            const int qt = c >> 1;
            const int u = j * ((NCH + 1) >> 1) + qt;
            const int KJ = min(128, QP - 128 * j);
            const int ksteps = KJ >> 4;
            for (int s = 0; s < ksteps; ++s) {
              const uint64_t ad = smem_desc_mnmajor(smem_u32(sDS + (u & 1) * 32768 + s * 2048), 16384);
              const uint64_t bd = smem_desc_mnmajor(smem_u32(sK + (j * 128 + 16 * s) * 128), 8192);
              umma_ss(tb + uint32_t(TM_DQ + 64 * qt), ad, bd, idesc64_mn, (j > 0 || s > 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_m2 + buf);
          ASTAMP(6 + 4 * t);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== math warps: lane quadrant = warp % 4, column half = warp / 4 =====================
    const int quad = warp & 3, half = warp >> 2;
    const int keyrow = quad * 32 + lane;
    const uint32_t trow = tb + (uint32_t(quad * 32) << 16);
    const int tid = threadIdx.x;  // 0..255
    {
      // delta[q] = sum_d dO[q,d] * O[q,d] (the softmax-backward row term), one query row per thread, straight from the
      // TMA-staged dO and O tiles -- no separate pass over HBM
      mbar_wait(bar_vdo, 0, 5);
      mbar_wait(bar_o, 0, 6);
      float dsum = 0.f;
      if (tid < N) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 a = *reinterpret_cast<const uint4*>(sDO + swz128(tid, c));
          const uint4 o = *reinterpret_cast<const uint4*>(sOut + swz128(tid, c));
          const uint32_t au[4] = {a.x, a.y, a.z, a.w}, ou[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = unpack16(au[e], H16), y = unpack16(ou[e], H16);
            dsum = fmaf(x.x, y.x, dsum);
            dsum = fmaf(x.y, y.y, dsum);
          }
        }
        delta[((long long)b * H + h) * N + tid] = dsum;
      }
      s_delta[tid] = -dsum;   // negated (see bwd_math)
      named_bar_sync(1, 256);
    }
    if (tid == 0) ASTAMP(64);
    const unsigned long long dseed = DROP ? __ldg(drop.seed) : 0ull;
    const unsigned long long dstep = (unsigned long long)(Npad >> 3);
    const unsigned long long drow0 = ((unsigned long long)b * H + h) * N;   // first query row of this (image, head)
    int t = 0;
    for (int j = 0; j < NJ; ++j) {
      const bool keyvalid = (j * 128 + keyrow) < N;
      const int dkey = j * 128 + keyrow;
      for (int c = 0; c < NCH; ++c, ++t) {
        const int QC = min(64, QP - 64 * c);
        const int g0 = half * 32;                       // this warp's columns: [g0, min(g0 + 32, QC))
        const int w = min(32, QC - g0);                 // 32, 16 or <= 0 (QC is a multiple of 16)
        const int buf = t & 1;
        const int u = j * ((NCH + 1) >> 1) + (c >> 1);
        uint8_t* sDSblk = sDS + (u & 1) * 32768 + (c & 1) * 16384;
        if (tid == 0) ASTAMP(68 + 8 * t);
        mbar_wait(bar_s, t & 1, 7);
        tc_fence_after();
        if (tid == 0) ASTAMP(69 + 8 * t);
        uint32_t sv[32], dv[32];
        if (w == 32) {
          tmem_ld32_nowait(trow + uint32_t(TM_S + g0), sv);
          tmem_ld32_nowait(trow + uint32_t(TM_DP + g0), dv);
        } else if (w == 16) {
          tmem_ld16_nowait(trow + uint32_t(TM_S + g0), sv);
          tmem_ld16_nowait(trow + uint32_t(TM_DP + g0), dv);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar_ld);                            // S^T / dP^T are free for chunk t+1
        if (tid == 0) ASTAMP(70 + 8 * t);
        // the P/dS buffer and the staging block were last read by the MMAs committed two chunks ago
        if (t >= 2) mbar_wait(bar_m2 + buf, ((t >> 1) - 1) & 1, 8);
        tc_fence_after();
        if (tid == 0) ASTAMP(71 + 8 * t);
        // DROP: counter block of (query c*64 + g0, key dkey); padded queries / keys produce P = 0 whatever the factor
        const unsigned long long dblk0 = DROP ? (drow0 + (unsigned long long)(c * 64 + g0)) * dstep + (unsigned long long)(dkey >> 3) : 0ull;
        if (w == 32)
          bwd_math<H16, 32, DROP>(sv, dv, trow, g0, buf, s_lse2, s_delta, c * 64, scale_log2, keyvalid, sDSblk, keyrow, drop, dseed, dblk0,
                                  dstep, dkey & 7);
        else if (w == 16)
          bwd_math<H16, 16, DROP>(sv, dv, trow, g0, buf, s_lse2, s_delta, c * 64, scale_log2, keyvalid, sDSblk, keyrow, drop, dseed, dblk0,
                                  dstep, dkey & 7);
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(bar_p + buf);
        if (tid == 0) ASTAMP(72 + 8 * t);
        if (c == NCH - 1) {
          // ---- key tile j is complete: dV_j (half 0) / dK_j (half 1) -> 16-bit -> staging -> TMA store
          mbar_wait(bar_m2 + buf, (t >> 1) & 1, 9);
          tc_fence_after();
          if (warp == 0 && elect_one()) tma_store_wait_read();  // earlier stores have finished reading sOut
          named_bar_sync(1, 256);
          const float mul = half == 0 ? 1.f : scale;
          uint8_t* dst = sOut + half * 16384;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld32_nowait(trow + uint32_t((half == 0 ? TM_DV : TM_DK) + 32 * hh), v);
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              uint4 o;
              o.x = pk16<H16>(__uint_as_float(v[8 * cc + 0]) * mul, __uint_as_float(v[8 * cc + 1]) * mul);
              o.y = pk16<H16>(__uint_as_float(v[8 * cc + 2]) * mul, __uint_as_float(v[8 * cc + 3]) * mul);
              o.z = pk16<H16>(__uint_as_float(v[8 * cc + 4]) * mul, __uint_as_float(v[8 * cc + 5]) * mul);
              o.w = pk16<H16>(__uint_as_float(v[8 * cc + 6]) * mul, __uint_as_float(v[8 * cc + 7]) * mul);
              *reinterpret_cast<uint4*>(dst + swz128(keyrow, 4 * hh + cc)) = o;
            }
          }
          tc_fence_before();
          fence_proxy_async();
          named_bar_sync(1, 256);
          if (warp == 0 && elect_one()) {
            tma_store_3d(&tmDQKV, sOut, (2 * H + h) * DH, j * 128, b);          // dV_j
            tma_store_3d(&tmDQKV, sOut + 16384, (H + h) * DH, j * 128, b);      // dK_j
            tma_store_commit();
          }
          if (tid == 0) ASTAMP(73 + 8 * t);
        }
      }
    }
    // ---- dQ tiles (accumulated over all key tiles; every MMA is complete: the last bar_m2 wait above covered them)
    const int NQT = (NCH + 1) >> 1;           // tiles with an accumulator
    const int NQT_ALL = (NCH_ALL + 1) >> 1;   // tiles of the output (rows outside the walked chunks are stored as zeros)
    if (warp == 0 && elect_one()) tma_store_wait_read();
    named_bar_sync(1, 256);
    if (half < NQT_ALL) {
      uint8_t* dst = sOut + half * 16384;
      const bool have = half < NQT;                              // warp-uniform
      const bool live = have && keyrow < 64 * (NCH - 2 * half);  // this thread's query row lies inside a walked chunk
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        if (have) {
          tmem_ld32_nowait(trow + uint32_t(TM_DQ + 64 * half + 32 * hh), v);
          tmem_ld_wait();
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (live) {
            o.x = pk16<H16>(__uint_as_float(v[8 * cc + 0]) * scale, __uint_as_float(v[8 * cc + 1]) * scale);
            o.y = pk16<H16>(__uint_as_float(v[8 * cc + 2]) * scale, __uint_as_float(v[8 * cc + 3]) * scale);
            o.z = pk16<H16>(__uint_as_float(v[8 * cc + 4]) * scale, __uint_as_float(v[8 * cc + 5]) * scale);
            o.w = pk16<H16>(__uint_as_float(v[8 * cc + 6]) * scale, __uint_as_float(v[8 * cc + 7]) * scale);
          }
          *reinterpret_cast<uint4*>(dst + swz128(keyrow, 4 * hh + cc)) = o;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    named_bar_sync(1, 256);
    if (warp == 0 && elect_one()) {
      for (int qt = 0; qt < NQT_ALL; ++qt) tma_store_3d(&tmDQKV, sOut + qt * 16384, h * DH, qt * 128, b);
      tma_store_commit();
      tma_store_wait_read();   // shared memory must outlive the reads; the writes themselves complete asynchronously
    }
    if (tid == 0) ASTAMP(65);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tb);
}

// ================================================================================================ backward, > 240 tokens
// Longer sequences (384x384 images: 577 tokens; patch 8: 785 / 1025 ...) do not fit the single-CTA kernel above (it keeps Q, K, V,
// dO and O of one (image, head) in shared memory).  Two streaming kernels replace it, each with 2 CTAs per SM (256 TMEM columns):
//   dK/dV: CTA per (128-key tile, head, image); K_j, V_j resident, 64-query chunks of Q / dO through a 2-stage TMA ring:
//          S^T = K_j Q_c^T, dP^T = V_j dO_c^T (SS) -> P^T, dS^T 16-bit IN PLACE (lane = key) -> dV += P^T dO_c, dK += dS^T Q_c (TS)
//   dQ:    CTA per (128-query tile, head, image); Q, dO resident, 64-key chunks of K / V through the ring:
//          S = Q K_c^T, dP = dO V_c^T (SS) -> dS 16-bit in place (lane = query) -> dQ += dS K_c (TS)
// delta = rowsum(dO o O) comes from attn_delta_kernel (attention.cu).  The chunk steps of a CTA are serial (MMA -> math -> MMA);
// the second CTA of the SM fills the gaps.  q_limit > 0: dO is zero from query row q_limit on (and out / lse / delta are not
// valid there): the dK/dV kernel walks only the chunks below it, the dQ kernel writes zeros for those rows.
constexpr int BL_THREADS = 160;
constexpr int BL_S = 0, BL_DP = 64, BL_ACC0 = 128, BL_ACC1 = 192;

// P^T / dS^T (or dS) of W columns for one TMEM lane: sv / dv = S and dP values, nl2[i] = -lse*log2e and nd[i] = -delta of column i's
// query (dK/dV: lane = key, columns = queries) -- or one pair for every column (dQ: lane = query).  fac[i] = dropout factor.
template <bool H16, int W>
__device__ __forceinline__ void bl_pack(const uint32_t* sv, const uint32_t* dv, const float* nl2, const float* nd, const float* fac,
                                        bool per_col, bool have_fac, float scale_log2, uint32_t valid_mask, uint32_t* ph, uint32_t* dh) {
#pragma unroll
  for (int i = 0; i < W; i += 2) {
    const float l0 = per_col ? nl2[i] : nl2[0], l1 = per_col ? nl2[i + 1] : nl2[0];
    const float d0 = per_col ? nd[i] : nd[0], d1 = per_col ? nd[i + 1] : nd[0];
    float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), scale_log2, l0));
    float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), scale_log2, l1));
    if (!((valid_mask >> i) & 1u)) p0 = 0.f;
    if (!((valid_mask >> (i + 1)) & 1u)) p1 = 0.f;
    const float m0 = have_fac ? fac[i] : 1.f, m1 = have_fac ? fac[i + 1] : 1.f;
    const float s0 = p0 * fmaf(__uint_as_float(dv[i]), m0, d0), s1 = p1 * fmaf(__uint_as_float(dv[i + 1]), m1, d1);
    ph[i >> 1] = pk16<H16>(p0 * m0, p1 * m1);
    dh[i >> 1] = pk16<H16>(s0, s1);
  }
}

template <bool H16, bool DROP>
__global__ void __launch_bounds__(BL_THREADS, 2)
    attn_bwd_dkdv_long_kernel(const __grid_constant__ CUtensorMap tmKV128, const __grid_constant__ CUtensorMap tmQ64,
                              const __grid_constant__ CUtensorMap tmDO64, const __grid_constant__ CUtensorMap tmDQKV,
                              const float* __restrict__ lse, const float* __restrict__ delta, int N, int H, float scale,
                              float scale_log2, int q_limit, DropSpec drop, int Npad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;                 // [128][64]; dK staging at the end
  uint8_t* sV = sK + 16384;           // [128][64]; dV staging at the end
  uint8_t* sQc = sV + 16384;          // 2 x [64][64]
  uint8_t* sDOc = sQc + 16384;        // 2 x [64][64]
  float* s_nl2 = reinterpret_cast<float*>(sDOc + 16384);   // [2][64]  -lse * log2e of the chunk's queries
  float* s_nd = s_nl2 + 128;                               // [2][64]  -delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_nd + 128);
  uint64_t* bar_kv = bars + 0;
  uint64_t* bar_q = bars + 1;         // [2] ring stage filled
  uint64_t* bar_s = bars + 3;         // S^T, dP^T of chunk c complete
  uint64_t* bar_p = bars + 4;         // P^T, dS^T written (128 arrivals)
  uint64_t* bar_m = bars + 5;         // dV / dK MMAs of chunk c complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int QP = (N + 15) & ~15;
  const int q_end = q_limit > 0 && q_limit < QP ? ((q_limit + 15) & ~15) : QP;   // queries walked (multiple of 16)
  const int NC = (q_end + 63) >> 6;

  if (warp == 4) {
    if (elect_one()) {
      prefetch_tmap(&tmKV128);
      prefetch_tmap(&tmQ64);
      prefetch_tmap(&tmDO64);
      prefetch_tmap(&tmDQKV);
      mbar_init(bar_kv, 1);
      mbar_init(bar_q + 0, 1);
      mbar_init(bar_q + 1, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_m, 1);
      mbar_init_fence();
      pdl_wait();
      mbar_expect_tx(bar_kv, 2 * 16384);
      tma_load_3d(sK, &tmKV128, bar_kv, (H + h) * DH, jt * 128, b);
      tma_load_3d(sV, &tmKV128, bar_kv, (2 * H + h) * DH, jt * 128, b);
      for (int c = 0; c < 2 && c < NC; ++c) {
        mbar_expect_tx(bar_q + c, 2 * 8192);
        tma_load_3d(sQc + c * 8192, &tmQ64, bar_q + c, h * DH, c * 64, b);
        tma_load_3d(sDOc + c * 8192, &tmDO64, bar_q + c, h * DH, c * 64, b);
      }
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  pdl_trigger();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      const uint32_t idesc64_ts = idesc_f16(DH, false, true, H16);
      const uint64_t kd = smem_desc_kmajor(smem_u32(sK)), vd = smem_desc_kmajor(smem_u32(sV));
      mbar_wait(bar_kv, 0, 1);
      for (int c = 0; c < NC; ++c) {
        const int st = c & 1;
        const int QC = min(64, q_end - 64 * c);
        mbar_wait(bar_q + st, (c >> 1) & 1, 2);
        tc_fence_after();
        const uint32_t idesc_s = idesc_f16(QC, false, false, H16);
        const uint64_t qd = smem_desc_kmajor(smem_u32(sQc + st * 8192)), dd = smem_desc_kmajor(smem_u32(sDOc + st * 8192));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tb + BL_S, kd + uint64_t(2 * k), qd + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tb + BL_DP, vd + uint64_t(2 * k), dd + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
        mbar_wait(bar_p, c & 1, 3);
        tc_fence_after();
        const int qsteps = QC >> 4;
        for (int s = 0; s < qsteps; ++s) {     // dV += P^T dO_c
          const uint64_t bd = smem_desc_mnmajor(smem_u32(sDOc + st * 8192 + 16 * s * 128), 8192);
          umma_ts(tb + BL_ACC0, tb + uint32_t(BL_S + 8 * s), bd, idesc64_ts, (c > 0 || s > 0) ? 1u : 0u);
        }
        for (int s = 0; s < qsteps; ++s) {     // dK += dS^T Q_c
          const uint64_t bd = smem_desc_mnmajor(smem_u32(sQc + st * 8192 + 16 * s * 128), 8192);
          umma_ts(tb + BL_ACC1, tb + uint32_t(BL_DP + 8 * s), bd, idesc64_ts, (c > 0 || s > 0) ? 1u : 0u);
        }
        umma_commit(bar_m);
        mbar_wait(bar_m, c & 1, 4);              // serial: S^T / dP^T columns and the ring stage are free again
        tc_fence_after();
        if (c + 2 < NC) {
          mbar_expect_tx(bar_q + st, 2 * 8192);
          tma_load_3d(sQc + st * 8192, &tmQ64, bar_q + st, h * DH, (c + 2) * 64, b);
          tma_load_3d(sDOc + st * 8192, &tmDO64, bar_q + st, h * DH, (c + 2) * 64, b);
        }
      }
    }
    __syncwarp();
  } else {
    const int keyrow = warp * 32 + lane;
    const int key = jt * 128 + keyrow;
    const bool keyvalid = key < N;
    const uint32_t trow = tb + (uint32_t(warp * 32) << 16);
    const int tid = threadIdx.x;   // 0..127
    const long long bh = (long long)b * H + h;
    const unsigned long long dseed = DROP ? __ldg(drop.seed) : 0ull;
    const unsigned long long dstep = (unsigned long long)(Npad >> 3);
    for (int c = 0; c < NC; ++c) {
      const int st = c & 1;
      const int QC = min(64, q_end - 64 * c);
      {   // the chunk's row statistics -> shared memory (the stage's previous readers finished two chunks ago)
        const int qq = 64 * c + (tid & 63);
        if (tid < 64) s_nl2[st * 64 + tid] = qq < N ? -lse[bh * N + qq] * LOG2E : -INFINITY;
        else s_nd[st * 64 + (tid - 64)] = qq < N ? -delta[bh * N + qq] : 0.f;
      }
      named_bar_sync(1, 128);
      mbar_wait(bar_s, c & 1, 5);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int g0 = 32 * half;
        const int w = min(32, QC - g0);        // 32, 16 or <= 0 (warp-uniform)
        if (w <= 0) break;
        uint32_t sv[32], dv[32], ph[16], dh[16];
        if (w == 32) {
          tmem_ld32_nowait(trow + uint32_t(BL_S + g0), sv);
          tmem_ld32_nowait(trow + uint32_t(BL_DP + g0), dv);
        } else {
          tmem_ld16_nowait(trow + uint32_t(BL_S + g0), sv);
          tmem_ld16_nowait(trow + uint32_t(BL_DP + g0), dv);
        }
        float fac[32];
        if constexpr (DROP) {
          const unsigned long long blk0 = ((unsigned long long)bh * N + (unsigned long long)(64 * c + g0)) * dstep + (unsigned long long)(key >> 3);
          const int dhalf = key & 7;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < w) {
              const uint4 bits = drop_bits8(dseed, drop.site, blk0 + (unsigned long long)i * dstep);
              const uint32_t wd = (dhalf >> 1) == 0 ? bits.x : (dhalf >> 1) == 1 ? bits.y : (dhalf >> 1) == 2 ? bits.z : bits.w;
              fac[i] = (((dhalf & 1) ? (wd >> 16) : (wd & 0xffffu)) >= drop.thresh) ? drop.inv_keep : 0.f;
            } else {
              fac[i] = 0.f;
            }
          }
        }
        tmem_ld_wait();
        const uint32_t vmask = keyvalid ? 0xffffffffu : 0u;
        if (w == 32) bl_pack<H16, 32>(sv, dv, s_nl2 + st * 64 + g0, s_nd + st * 64 + g0, fac, true, DROP, scale_log2, vmask, ph, dh);
        else bl_pack<H16, 16>(sv, dv, s_nl2 + st * 64 + g0, s_nd + st * 64 + g0, fac, true, DROP, scale_log2, vmask, ph, dh);
        if (w == 32) {
          tmem_st16_nowait(trow + uint32_t(BL_S + (g0 >> 1)), ph);
          tmem_st16_nowait(trow + uint32_t(BL_DP + (g0 >> 1)), dh);
        } else {
          tmem_st8_nowait(trow + uint32_t(BL_S + (g0 >> 1)), ph);
          tmem_st8_nowait(trow + uint32_t(BL_DP + (g0 >> 1)), dh);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    // ---- dV (x 1) and dK (x scale) -> 16-bit -> staging (K / V are dead) -> TMA store (rows >= N clipped)
    mbar_wait(bar_m, (NC - 1) & 1, 6);
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint8_t* dst = which == 0 ? sV : sK;
      const float mul = which == 0 ? 1.f : scale;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32_nowait(trow + uint32_t((which == 0 ? BL_ACC0 : BL_ACC1) + 32 * hh), v);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint4 o;
          o.x = pk16<H16>(__uint_as_float(v[8 * cc + 0]) * mul, __uint_as_float(v[8 * cc + 1]) * mul);
          o.y = pk16<H16>(__uint_as_float(v[8 * cc + 2]) * mul, __uint_as_float(v[8 * cc + 3]) * mul);
          o.z = pk16<H16>(__uint_as_float(v[8 * cc + 4]) * mul, __uint_as_float(v[8 * cc + 5]) * mul);
          o.w = pk16<H16>(__uint_as_float(v[8 * cc + 6]) * mul, __uint_as_float(v[8 * cc + 7]) * mul);
          *reinterpret_cast<uint4*>(dst + swz128(keyrow, 4 * hh + cc)) = o;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    named_bar_sync(1, 128);
    if (warp == 0 && elect_one()) {
      tma_store_3d(&tmDQKV, sV, (2 * H + h) * DH, jt * 128, b);
      tma_store_3d(&tmDQKV, sK, (H + h) * DH, jt * 128, b);
      tma_store_commit();
      tma_store_wait_read();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<256>(tb);
}

template <bool H16, bool DROP>
__global__ void __launch_bounds__(BL_THREADS, 2)
    attn_bwd_dq_long_kernel(const __grid_constant__ CUtensorMap tmQ128, const __grid_constant__ CUtensorMap tmDO128,
                            const __grid_constant__ CUtensorMap tmKV64, const __grid_constant__ CUtensorMap tmDQKV,
                            const float* __restrict__ lse, const float* __restrict__ delta, int N, int H, float scale,
                            float scale_log2, int q_limit, DropSpec drop, int Npad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                 // [128][64]; dQ staging at the end
  uint8_t* sDO = sQ + 16384;          // [128][64]
  uint8_t* sKc = sDO + 16384;         // 2 x [64][64]
  uint8_t* sVc = sKc + 16384;         // 2 x [64][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sVc + 16384);
  uint64_t* bar_qdo = bars + 0;
  uint64_t* bar_kv = bars + 1;        // [2]
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_p = bars + 4;
  uint64_t* bar_m = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int KP = (N + 15) & ~15;
  const int NJ = (KP + 63) >> 6;
  const int q_live = q_limit > 0 && q_limit < N ? q_limit : N;   // query rows with a gradient
  const bool tile_live = q0 < q_live;                             // block-uniform

  if (warp == 4) {
    if (elect_one()) {
      prefetch_tmap(&tmQ128);
      prefetch_tmap(&tmDO128);
      prefetch_tmap(&tmKV64);
      prefetch_tmap(&tmDQKV);
      mbar_init(bar_qdo, 1);
      mbar_init(bar_kv + 0, 1);
      mbar_init(bar_kv + 1, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_m, 1);
      mbar_init_fence();
      pdl_wait();
      if (tile_live) {
        mbar_expect_tx(bar_qdo, 2 * 16384);
        tma_load_3d(sQ, &tmQ128, bar_qdo, h * DH, q0, b);
        tma_load_3d(sDO, &tmDO128, bar_qdo, h * DH, q0, b);
        for (int j = 0; j < 2 && j < NJ; ++j) {
          mbar_expect_tx(bar_kv + j, 2 * 8192);
          tma_load_3d(sKc + j * 8192, &tmKV64, bar_kv + j, (H + h) * DH, j * 64, b);
          tma_load_3d(sVc + j * 8192, &tmKV64, bar_kv + j, (2 * H + h) * DH, j * 64, b);
        }
      }
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  pdl_trigger();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  if (warp == 4) {
    if (tile_live && elect_one()) {
      const uint32_t idesc64_ts = idesc_f16(DH, false, true, H16);
      const uint64_t qd = smem_desc_kmajor(smem_u32(sQ)), dd = smem_desc_kmajor(smem_u32(sDO));
      mbar_wait(bar_qdo, 0, 1);
      for (int j = 0; j < NJ; ++j) {
        const int st = j & 1;
        const int KC = min(64, KP - 64 * j);
        mbar_wait(bar_kv + st, (j >> 1) & 1, 2);
        tc_fence_after();
        const uint32_t idesc_s = idesc_f16(KC, false, false, H16);
        const uint64_t kd = smem_desc_kmajor(smem_u32(sKc + st * 8192)), vd = smem_desc_kmajor(smem_u32(sVc + st * 8192));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tb + BL_S, qd + uint64_t(2 * k), kd + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tb + BL_DP, dd + uint64_t(2 * k), vd + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
        mbar_wait(bar_p, j & 1, 3);
        tc_fence_after();
        const int ksteps = KC >> 4;
        for (int s = 0; s < ksteps; ++s) {     // dQ += dS K_c
          const uint64_t bd = smem_desc_mnmajor(smem_u32(sKc + st * 8192 + 16 * s * 128), 8192);
          umma_ts(tb + BL_ACC0, tb + uint32_t(BL_S + 8 * s), bd, idesc64_ts, (j > 0 || s > 0) ? 1u : 0u);
        }
        umma_commit(bar_m);
        mbar_wait(bar_m, j & 1, 4);
        tc_fence_after();
        if (j + 2 < NJ) {
          mbar_expect_tx(bar_kv + st, 2 * 8192);
          tma_load_3d(sKc + st * 8192, &tmKV64, bar_kv + st, (H + h) * DH, (j + 2) * 64, b);
          tma_load_3d(sVc + st * 8192, &tmKV64, bar_kv + st, (2 * H + h) * DH, (j + 2) * 64, b);
        }
      }
    }
    __syncwarp();
  } else {
    const int row = warp * 32 + lane;
    const int q = q0 + row;
    const uint32_t trow = tb + (uint32_t(warp * 32) << 16);
    const long long bh = (long long)b * H + h;
    const bool rowlive = q < q_live;
    if (tile_live) {
      float nl2 = rowlive ? -lse[bh * N + q] * LOG2E : -INFINITY;
      float nd = rowlive ? -delta[bh * N + q] : 0.f;
      const unsigned long long dseed = DROP ? __ldg(drop.seed) : 0ull;
      const unsigned long long rowblk = DROP ? (((unsigned long long)bh * N + (rowlive ? q : 0)) * (unsigned long long)Npad) >> 3 : 0ull;
      for (int j = 0; j < NJ; ++j) {
        const int KC = min(64, KP - 64 * j);
        const int key0 = 64 * j;
        mbar_wait(bar_s, j & 1, 5);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int g0 = 32 * half;
          const int w = min(32, KC - g0);
          if (w <= 0) break;
          uint32_t sv[32], dv[32], ph[16], dh[16];
          if (w == 32) {
            tmem_ld32_nowait(trow + uint32_t(BL_S + g0), sv);
            tmem_ld32_nowait(trow + uint32_t(BL_DP + g0), dv);
          } else {
            tmem_ld16_nowait(trow + uint32_t(BL_S + g0), sv);
            tmem_ld16_nowait(trow + uint32_t(BL_DP + g0), dv);
          }
          float fac[32];
          if constexpr (DROP) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint4 bits = drop_bits8(dseed, drop.site, rowblk + (unsigned long long)(((key0 + g0) >> 3) + t));
              const uint32_t wd[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = drop_pair(wd[e], drop.thresh, drop.inv_keep);
                fac[8 * t + 2 * e] = f.x;
                fac[8 * t + 2 * e + 1] = f.y;
              }
            }
          }
          tmem_ld_wait();
          const int nvalid = N - (key0 + g0);                        // keys of this half below N
          const uint32_t vmask = !rowlive ? 0u : nvalid >= 32 ? 0xffffffffu : nvalid <= 0 ? 0u : ((1u << nvalid) - 1u);
          if (w == 32) bl_pack<H16, 32>(sv, dv, &nl2, &nd, fac, false, DROP, scale_log2, vmask, ph, dh);
          else bl_pack<H16, 16>(sv, dv, &nl2, &nd, fac, false, DROP, scale_log2, vmask, ph, dh);
          // only dS is an operand here (dQ += dS K); it goes where S was
          if (w == 32) tmem_st16_nowait(trow + uint32_t(BL_S + (g0 >> 1)), dh);
          else tmem_st8_nowait(trow + uint32_t(BL_S + (g0 >> 1)), dh);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_p);
      }
      mbar_wait(bar_m, (NJ - 1) & 1, 6);
      tc_fence_after();
    }
    // ---- dQ x scale -> 16-bit -> staging (Q is dead) -> TMA store; rows without a gradient are written as zeros
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];
      if (tile_live) {
        tmem_ld32_nowait(trow + uint32_t(BL_ACC0 + 32 * hh), v);
        tmem_ld_wait();
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (rowlive) {
          o.x = pk16<H16>(__uint_as_float(v[8 * cc + 0]) * scale, __uint_as_float(v[8 * cc + 1]) * scale);
          o.y = pk16<H16>(__uint_as_float(v[8 * cc + 2]) * scale, __uint_as_float(v[8 * cc + 3]) * scale);
          o.z = pk16<H16>(__uint_as_float(v[8 * cc + 4]) * scale, __uint_as_float(v[8 * cc + 5]) * scale);
          o.w = pk16<H16>(__uint_as_float(v[8 * cc + 6]) * scale, __uint_as_float(v[8 * cc + 7]) * scale);
        }
        *reinterpret_cast<uint4*>(sQ + swz128(row, 4 * hh + cc)) = o;
      }
    }
    tc_fence_before();
    fence_proxy_async();
    named_bar_sync(1, 128);
    if (warp == 0 && elect_one()) {
      tma_store_3d(&tmDQKV, sQ, h * DH, q0, b);
      tma_store_commit();
      tma_store_wait_read();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<256>(tb);
}

// ------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// 3-D 16-bit tensor map over [B][N][width] (row-major), box = 1 x box_rows x 64 columns, 128B swizzle.  Rows >= N are
// zero-filled on load and clipped on store, so ragged sequences (197 / 198 tokens) need no masking in the data path.
int make_tmap_3d(CUtensorMap* tm, const void* base, int width, int N, int B, int box_rows, bool fp16) {
  auto fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VITK_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)N * width * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed: CUresult %d (base=%p width=%d N=%d B=%d box_rows=%d)", (int)r, base, width, N, B,
              box_rows);
    return VITK_ERR_CUDA;
  }
  return VITK_OK;
}

}  // namespace

// Forward for N <= 256.  Returns VITK_OK or an error; the caller (attention.cu) owns argument validation.
template <bool H16, bool DROP>
int attention_fwd_tc_impl(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, int q_rows, DropSpec drop,
                          cudaStream_t st) {
  const int KP = (N + 15) & ~15;
  CUtensorMap tmQ, tmKV, tmO;
  int rc;
  if ((rc = make_tmap_3d(&tmQ, qkv, 3 * H * DH, N, B, 128, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmKV, qkv, 3 * H * DH, N, B, KP, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmO, out, H * DH, N, B, 128, H16)) != VITK_OK) return rc;
  const int smem = 128 * 128 + 2 * KP * 128 + 64 + 1024;
  auto kfn = attn_fwd_tc_kernel<H16, DROP>;
  static int configured = 0;
  if (configured < smem) {
    VITK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 + 2 * 256 * 128 + 64 + 1024));
    configured = 128 * 128 + 2 * 256 * 128 + 64 + 1024;
  }
  // q_rows > 0: only the first q_rows query rows of out / lse are needed -> only the 128-query tiles that hold them are launched
  const int q_need = q_rows > 0 && q_rows < N ? q_rows : N;
  dim3 grid((q_need + 127) / 128, H, B);
  VITK_CUDA(launch_pdl(kfn, grid, dim3(FWD_THREADS), (size_t)smem, st, tmQ, tmKV, tmO, lse, N, H, KP, scale, scale * LOG2E,
                       2 * num_sms(), drop, (N + 7) & ~7));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

// Forward for N > 256 (key tiles, see attn_fwd_tc_long_kernel).
template <bool H16, bool DROP, int TCOLS>
int attention_fwd_tc_long_launch(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const CUtensorMap& tmO, float* lse, int B, int N, int H,
                                 int KT, int NKT, int TPG, float scale, int q_rows, DropSpec drop, cudaStream_t st) {
  const int smem = 128 * 128 + 2 * TPG * KT * 128 + 192 + 1024;
  if (smem > 227 * 1024) {
    set_error("attention_fwd_tc_long: N=%d needs %d bytes of shared memory", N, smem);
    return VITK_ERR_INVALID;
  }
  auto kfn = attn_fwd_tc_long_kernel<H16, DROP, TCOLS>;
  static int configured = 0;
  if (configured < smem) {
    VITK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  const int q_need = q_rows > 0 && q_rows < N ? q_rows : N;
  dim3 grid((q_need + 127) / 128, H, B);
  VITK_CUDA(launch_pdl(kfn, grid, dim3(FWD_THREADS), (size_t)smem, st, tmQ, tmKV, tmO, lse, N, H, KT, NKT, TPG, scale, scale * LOG2E,
                       drop, (N + 7) & ~7));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

template <bool H16, bool DROP>
int attention_fwd_tc_long_impl(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, int q_rows, DropSpec drop,
                               cudaStream_t st) {
  // Two CTAs per SM: 128-key tiles in groups of two, 256 TMEM columns, 83 KB of shared memory -- one CTA's softmax runs under the
  // other's MMAs and loads.  Measured at 32 x 12 heads x 577 tokens: 129 us, against 190 us for the one-CTA-per-SM
  // configuration of the same kernel (TCOLS = 512: 256-key tiles, all of K / V resident up to 816 tokens) and 182 us for the
  // round-1 mma.sync kernel (profiles/r02_attention_long_sequences.txt).
  const int NKT = (N + 127) / 128;
  const int KT = (((N + NKT - 1) / NKT) + 15) & ~15;
  const int TPG = 2;
  if (KT > 128 || (NKT - 1) * KT >= N) {
    set_error("attention_fwd_tc_long: N=%d out of range", N);
    return VITK_ERR_INVALID;
  }
  CUtensorMap tmQ, tmKV, tmO;
  int rc;
  if ((rc = make_tmap_3d(&tmQ, qkv, 3 * H * DH, N, B, 128, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmKV, qkv, 3 * H * DH, N, B, KT, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmO, out, H * DH, N, B, 128, H16)) != VITK_OK) return rc;
  return attention_fwd_tc_long_launch<H16, DROP, 256>(tmQ, tmKV, tmO, lse, B, N, H, KT, NKT, TPG, scale, q_rows, drop, st);
}

template <bool H16, bool DROP>
int attention_bwd_tc_impl(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int B, int N,
                          int H, float scale, int q_rows, DropSpec drop, cudaStream_t st) {
  const int QP = (N + 15) & ~15;
  CUtensorMap tmQKV, tmDO, tmO, tmDQKV;
  int rc;
  if ((rc = make_tmap_3d(&tmQKV, qkv, 3 * H * DH, N, B, QP, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmDO, dout, H * DH, N, B, QP, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmO, out, H * DH, N, B, QP, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmDQKV, dqkv, 3 * H * DH, N, B, 128, H16)) != VITK_OK) return rc;
  const int smem = 4 * QP * 128 + 98304 + 2048 + 128 + 1024;
  const int smem_max = 4 * 240 * 128 + 98304 + 2048 + 128 + 1024;
  auto kfn = attn_bwd_tc_kernel<H16, DROP>;
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    configured = true;
  }
  dim3 grid(H, B);
  const int q_chunks = q_rows > 0 && q_rows < N ? (q_rows + 63) / 64 : 0;
  VITK_CUDA(launch_pdl(kfn, grid, dim3(BWD_THREADS), (size_t)smem, st, tmQKV, tmDO, tmO, tmDQKV, lse, delta, N, H, QP, scale,
                       scale * LOG2E, num_sms(), q_chunks, drop, (N + 7) & ~7));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

// Backward for N > 240: delta must already hold rowsum(dO o O) (attn_delta_kernel).  q_rows as in attention_bwd_tc_impl.
template <bool H16, bool DROP>
int attention_bwd_tc_long_impl(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, int B, int N, int H,
                               float scale, int q_rows, DropSpec drop, cudaStream_t st) {
  CUtensorMap tmQKV128, tmQKV64, tmDO128, tmDO64, tmDQKV;
  int rc;
  if ((rc = make_tmap_3d(&tmQKV128, qkv, 3 * H * DH, N, B, 128, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmQKV64, qkv, 3 * H * DH, N, B, 64, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmDO128, dout, H * DH, N, B, 128, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmDO64, dout, H * DH, N, B, 64, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmDQKV, dqkv, 3 * H * DH, N, B, 128, H16)) != VITK_OK) return rc;
  const int smem_kv = 4 * 16384 + 1024 + 64 + 1024, smem_q = 4 * 16384 + 64 + 1024;
  auto kkv = attn_bwd_dkdv_long_kernel<H16, DROP>;
  auto kq = attn_bwd_dq_long_kernel<H16, DROP>;
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
    VITK_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q));
    configured = true;
  }
  const int q_limit = q_rows > 0 && q_rows < N ? q_rows : 0;
  const int Npad = (N + 7) & ~7;
  dim3 grid((N + 127) / 128, H, B);
  VITK_CUDA(launch_pdl(kkv, grid, dim3(BL_THREADS), (size_t)smem_kv, st, tmQKV128, tmQKV64, tmDO64, tmDQKV, lse, delta, N, H, scale,
                       scale * LOG2E, q_limit, drop, Npad));
  VITK_LAUNCH_CHECK();
  VITK_CUDA(launch_pdl(kq, grid, dim3(BL_THREADS), (size_t)smem_q, st, tmQKV128, tmDO128, tmQKV64, tmDQKV, lse, delta, N, H, scale,
                       scale * LOG2E, q_limit, drop, Npad));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

#ifdef VITK_GEMM_KNOBS
extern "C" int vitk_debug_read_attn(long long* dst) { return (int)cudaMemcpyFromSymbol(dst, g_attn_dbg, sizeof(g_attn_dbg)); }
#endif

// Eval-mode attention maps (needs the lse the forward just wrote).
template <bool H16>
int attention_probs_tc_impl(const void* qkv, const float* lse, float* probs, long long batch_stride, int B, int N, int H, float scale,
                            cudaStream_t st) {
  const int NKT = (N + 255) / 256;
  const int KP = (((N + NKT - 1) / NKT) + 15) & ~15;       // keys per tile (<= 256); one tile up to 256 tokens
  const int NQT = (N + 127) / 128;
  CUtensorMap tmQ, tmKV;
  int rc;
  if ((rc = make_tmap_3d(&tmQ, qkv, 3 * H * DH, N, B, 128, H16)) != VITK_OK) return rc;
  if ((rc = make_tmap_3d(&tmKV, qkv, 3 * H * DH, N, B, KP, H16)) != VITK_OK) return rc;
  const int smem = 128 * 128 + KP * 128 + 4 * 32 * 33 * 4 + 64 + 1024;
  auto kfn = attn_probs_tc_kernel<H16>;
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 + 256 * 128 + 4 * 32 * 33 * 4 + 64 + 1024));
    configured = true;
  }
  dim3 grid(NQT * NKT, H, B);
  VITK_CUDA(launch_pdl(kfn, grid, dim3(FWD_THREADS), (size_t)smem, st, tmQ, tmKV, lse, probs, batch_stride, N, H, KP, scale * LOG2E, NQT));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
int attention_probs_tc(const void* qkv, const float* lse, float* probs, long long batch_stride, int B, int N, int H, float scale,
                       bool fp16, cudaStream_t st) {
  return fp16 ? attention_probs_tc_impl<true>(qkv, lse, probs, batch_stride, B, N, H, scale, st)
              : attention_probs_tc_impl<false>(qkv, lse, probs, batch_stride, B, N, H, scale, st);
}

// drop == nullptr: no attention-probability dropout (every configuration the reference ships)
int attention_fwd_tc(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, int q_rows, bool fp16, const DropSpec* drop,
                     cudaStream_t st) {
  const DropSpec none = make_drop_spec(nullptr, 0.f, 0);
  const bool dropping = drop != nullptr && drop->seed != nullptr;
  if (N > 256) {
    if (dropping)
      return fp16 ? attention_fwd_tc_long_impl<true, true>(qkv, out, lse, B, N, H, scale, q_rows, *drop, st)
                  : attention_fwd_tc_long_impl<false, true>(qkv, out, lse, B, N, H, scale, q_rows, *drop, st);
    return fp16 ? attention_fwd_tc_long_impl<true, false>(qkv, out, lse, B, N, H, scale, q_rows, none, st)
                : attention_fwd_tc_long_impl<false, false>(qkv, out, lse, B, N, H, scale, q_rows, none, st);
  }
  if (dropping)
    return fp16 ? attention_fwd_tc_impl<true, true>(qkv, out, lse, B, N, H, scale, q_rows, *drop, st)
                : attention_fwd_tc_impl<false, true>(qkv, out, lse, B, N, H, scale, q_rows, *drop, st);
  return fp16 ? attention_fwd_tc_impl<true, false>(qkv, out, lse, B, N, H, scale, q_rows, none, st)
              : attention_fwd_tc_impl<false, false>(qkv, out, lse, B, N, H, scale, q_rows, none, st);
}
int attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int B, int N,
                     int H, float scale, int q_rows, bool fp16, const DropSpec* drop, cudaStream_t st) {
  const DropSpec none = make_drop_spec(nullptr, 0.f, 0);
  if (drop != nullptr && drop->seed != nullptr)
    return fp16 ? attention_bwd_tc_impl<true, true>(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, q_rows, *drop, st)
                : attention_bwd_tc_impl<false, true>(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, q_rows, *drop, st);
  return fp16 ? attention_bwd_tc_impl<true, false>(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, q_rows, none, st)
              : attention_bwd_tc_impl<false, false>(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, q_rows, none, st);
}

int attention_bwd_tc_long(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, int B, int N, int H,
                          float scale, int q_rows, bool fp16, const DropSpec* drop, cudaStream_t st) {
  const DropSpec none = make_drop_spec(nullptr, 0.f, 0);
  if (drop != nullptr && drop->seed != nullptr)
    return fp16 ? attention_bwd_tc_long_impl<true, true>(qkv, dout, lse, delta, dqkv, B, N, H, scale, q_rows, *drop, st)
                : attention_bwd_tc_long_impl<false, true>(qkv, dout, lse, delta, dqkv, B, N, H, scale, q_rows, *drop, st);
  return fp16 ? attention_bwd_tc_long_impl<true, false>(qkv, dout, lse, delta, dqkv, B, N, H, scale, q_rows, none, st)
              : attention_bwd_tc_long_impl<false, false>(qkv, dout, lse, delta, dqkv, B, N, H, scale, q_rows, none, st);
}

}  // namespace vitk
