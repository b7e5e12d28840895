// ingest.cu -- GPU-side input pipeline, the step immediately before the hot path.
//
// Restates, for a batch of raw single-channel uint16 tiles already in HBM:
//   CARSThyroidDataset._preprocess_image   src/data/dataset.py:533-551    cv2.resize(INTER_LINEAR) on uint16, then / 65535
//   AdaptiveNormalization('percentile')    src/data/quality_preprocessing.py:282-326
//                                          per-image torch.quantile(1 %, 99 %), clamp, (x - lo) / (hi - lo + 1e-8)
//   gray -> 3 channels + T.Normalize       src/data/vit_transforms.py:381-393
//   MixUp / CutMix                          src/data/vit_transforms.py:396-462  (lam, permutation and box come from the host RNG)
// All of it is HBM-bound byte work: one coalesced pass per stage, 128-bit stores, grids capped at a multiple of the SM
// count.  The percentile needs exact order statistics: a radix select over the float bit patterns (4 passes of 8 bits
// per rank through a shared-memory histogram, one CTA per image; the 200 KB plane stays in L2).
#include "vitk_common.cuh"

namespace vitk {
namespace {

inline int capped_grid(long long work_items, int threads, int ctas_per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// cv2 INTER_LINEAR source coordinate: fx = (d + 0.5) * scale - 0.5 (computed in double, rounded to float), clamped to the
// image: (index of the left/top tap, weight of the right/bottom tap)
__device__ __forceinline__ void linear_tap(int d, double scale, int n_src, int& i0, float& w1) {
  float f = (float)((d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) {
    s = 0;
    f = 0.f;
  }
  if (s >= n_src - 1) {
    s = n_src - 1;
    f = 0.f;
  }
  i0 = s;
  w1 = f;
}

// gray[b, y, x] = round_to_u16(bilinear(raw[b])) / 65535 -- horizontal pass first, then vertical, in fp32 without FMA
// contraction, rounded half-to-even to the uint16 grid like cv2's saturate_cast before the division by 65535.
__global__ void __launch_bounds__(256)
    resize_u16_kernel(const unsigned short* __restrict__ raw, float* __restrict__ gray, int B, int Hs, int Ws, int H, int W,
                      double sx, double sy) {
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % W);
    const int y = int((i / W) % H);
    const long long b = i / ((long long)W * H);
    float v;
    const unsigned short* src = raw + b * (long long)Hs * Ws;
    if (Hs == H && Ws == W) {
      v = (float)src[(long long)y * Ws + x];   // dataset.py:537: resize only "if needed"
    } else {
      int x0, y0;
      float ax, ay;
      linear_tap(x, sx, Ws, x0, ax);
      linear_tap(y, sy, Hs, y0, ay);
      const int x1 = min(x0 + 1, Ws - 1), y1 = min(y0 + 1, Hs - 1);
      const float a0 = 1.f - ax, b0 = 1.f - ay;
      const float r0 = __fadd_rn(__fmul_rn((float)src[(long long)y0 * Ws + x0], a0), __fmul_rn((float)src[(long long)y0 * Ws + x1], ax));
      const float r1 = __fadd_rn(__fmul_rn((float)src[(long long)y1 * Ws + x0], a0), __fmul_rn((float)src[(long long)y1 * Ws + x1], ax));
      v = rintf(__fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, ay)));
      v = fminf(fmaxf(v, 0.f), 65535.f);
    }
    gray[i] = __fdiv_rn(v, 65535.f);
  }
}

// ---- exact order statistics of one image (n floats) by radix select on the monotone bit key
__device__ __forceinline__ unsigned int float_key(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
constexpr int SEL_THREADS = 1024;
constexpr long long SEL_SMEM_MAX_ELEMS = 53248;   // 208 KB of the 227 KB a CTA may own: a 224 x 224 fp32 tile (50 176) fits
// k-th smallest (0-based) of x[0..n) -- x points to the CTA's shared-memory copy of the image when it fits, else to global
// memory.  Every thread of the CTA calls it, every thread gets the result.
__device__ float select_kth(const float* x, long long n, long long k, unsigned int* hist /* smem [256] */,
                            unsigned int* bcast /* smem [2] */) {
  unsigned int prefix = 0, mask = 0;
  long long remaining = k;
  const int lane = threadIdx.x & 31;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int t = threadIdx.x; t < 256; t += SEL_THREADS) hist[t] = 0;
    __syncthreads();
    // image data is concentrated in a few exponent bins: aggregate equal digits inside the warp before touching the
    // histogram; four independent loads are in flight per thread before the first of them is consumed
    for (long long base = 0; base < n; base += 4 * SEL_THREADS) {
      float v[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = base + u * SEL_THREADS + threadIdx.x;
        ok[u] = i < n;
        v[u] = ok[u] ? x[i] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        unsigned int digit = 256u;   // sentinel: not a candidate
        const unsigned int key = float_key(v[u]);
        if (ok[u] && (key & mask) == prefix) digit = (key >> shift) & 255u;
        if (__ballot_sync(0xffffffffu, digit != 256u) == 0u) continue;   // later passes: most warps hold no candidate at all
        const unsigned int peers = __match_any_sync(0xffffffffu, digit);
        if (digit != 256u && lane == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned)__popc(peers));
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // warp 0: find the bin that holds rank `remaining` (8 bins per lane + a warp scan)
      unsigned int h[8], mine = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        h[j] = hist[8 * lane + j];
        mine += h[j];
      }
      unsigned int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned int excl = incl - mine;
      if ((long long)excl <= remaining && remaining < (long long)incl) {   // exactly one lane
        unsigned int acc = excl;
        int d = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (remaining >= (long long)(acc + h[j])) {
            acc += h[j];
            d = j + 1;
          } else {
            break;
          }
        }
        bcast[0] = (unsigned int)(8 * lane + d);
        bcast[1] = acc;
      }
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 255u << shift;
    remaining -= bcast[1];
    __syncthreads();
  }
  return key_float(prefix);
}

// torch.lerp for floats (ATen Lerp.h): one formula per half so that lerp(a, b, 1) == b exactly
__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
  const float d = b - a;
  return w < 0.5f ? a + w * d : b - d * (1.f - w);
}

// bounds[b] = (quantile(x_b, q_lo), quantile(x_b, q_hi)) with torch.quantile's default 'linear' interpolation
__global__ void __launch_bounds__(SEL_THREADS)
    percentile_bounds_kernel(const float* __restrict__ x, long long n, float q_lo, float q_hi, float* __restrict__ bounds) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int bcast[2];
  extern __shared__ float4 s_img4[];
  const float* xb = x + (long long)blockIdx.x * n;
  if (n <= SEL_SMEM_MAX_ELEMS) {   // stage the image once: all 16+ selection passes then run out of shared memory
    float* s_img = reinterpret_cast<float*>(s_img4);
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0) {
      for (long long i = threadIdx.x; i < (n >> 2); i += SEL_THREADS) s_img4[i] = ldg_f4(xb + 4 * i);
    } else {
      for (long long i = threadIdx.x; i < n; i += SEL_THREADS) s_img[i] = __ldg(xb + i);
    }
    __syncthreads();
    xb = s_img;
  }
  float res[2];
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const float q = which == 0 ? q_lo : q_hi;
    const float rank = q * (float)(n - 1);            // ranks are computed in the tensor's dtype (fp32), as torch does
    const float below = floorf(rank);
    const long long k0 = (long long)below;
    const long long k1 = (long long)ceilf(rank);
    const float v0 = select_kth(xb, n, k0, hist, bcast);
    const float v1 = k1 == k0 ? v0 : select_kth(xb, n, k1, hist, bcast);
    res[which] = torch_lerp(v0, v1, rank - below);
  }
  if (threadIdx.x == 0) {
    bounds[2 * blockIdx.x] = res[0];
    bounds[2 * blockIdx.x + 1] = res[1];
  }
}

// out[b, c, y, x] = (mix(norm(gray[b]), norm(gray[perm[b]])) - mean[c]) / std[c]
//   norm: optional clamp to [lo_b, hi_b] and (v - lo) / (hi - lo + 1e-8)       (AdaptiveNormalization)
//   mix : MixUp lam * a + (1 - lam) * b everywhere, or CutMix: b inside the box [y1,y2) x [x1,x2), a outside
struct FinishParams {
  const float* gray;
  const float* bounds;   // [B,2] or nullptr
  const int* perm;       // [B] or nullptr (no mixing)
  float* out;
  int B, C, H, W;
  float mean[4], stdv[4];
  int has_norm;
  int cutmix;            // 0: MixUp with lam; 1: CutMix box
  float lam;
  int x1, y1, x2, y2;
};
__device__ __forceinline__ float adapt(float v, const float* __restrict__ bounds, long long b) {
  if (bounds == nullptr) return v;
  const float lo = __ldg(bounds + 2 * b), hi = __ldg(bounds + 2 * b + 1);
  v = fminf(fmaxf(v, lo), hi);
  return __fdiv_rn(v - lo, (hi - lo) + 1e-8f);
}
__global__ void __launch_bounds__(256) finish_tiles_kernel(FinishParams p) {
  const int W4 = p.W >> 2;
  const long long total = (long long)p.B * p.H * W4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xv = int(i % W4);
    const int y = int((i / W4) % p.H);
    const long long b = i / ((long long)W4 * p.H);
    const long long plane = (long long)p.H * p.W;
    const float4 g = ldg_f4(p.gray + b * plane + (long long)y * p.W + 4 * xv);
    float v[4] = {adapt(g.x, p.bounds, b), adapt(g.y, p.bounds, b), adapt(g.z, p.bounds, b), adapt(g.w, p.bounds, b)};
    float o[4][4];   // [channel][x]
    long long pb = -1;
    float u[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.perm != nullptr) {
      pb = __ldg(p.perm + b);
      const float4 h = ldg_f4(p.gray + pb * plane + (long long)y * p.W + 4 * xv);
      u[0] = adapt(h.x, p.bounds, pb); u[1] = adapt(h.y, p.bounds, pb); u[2] = adapt(h.z, p.bounds, pb); u[3] = adapt(h.w, p.bounds, pb);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= p.C) break;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // Normalize first (the reference mixes already-normalised loader batches), then mix
        float a = p.has_norm ? __fdiv_rn(v[j] - p.mean[c], p.stdv[c]) : v[j];   // T.Normalize: sub, then div
        if (p.perm != nullptr) {
          const float bb = p.has_norm ? __fdiv_rn(u[j] - p.mean[c], p.stdv[c]) : u[j];
          if (p.cutmix) {
            const int x = 4 * xv + j;
            a = (x >= p.x1 && x < p.x2 && y >= p.y1 && y < p.y2) ? bb : a;
          } else {
            a = __fadd_rn(__fmul_rn(p.lam, a), __fmul_rn(1.f - p.lam, bb));
          }
        }
        o[c][j] = a;
      }
      *reinterpret_cast<float4*>(p.out + (b * p.C + c) * plane + (long long)y * p.W + 4 * xv) =
          make_float4(o[c][0], o[c][1], o[c][2], o[c][3]);
    }
  }
}

// ------------------------------------------------------------------ gray tiles -> 16-bit patch matrix (no NCHW round trip)
// patches[(b, py, px), c*P*P + ky*P + kx] = 16-bit( (norm(tile[b, py*P+ky, px*P+kx]) - mean[c]) / std[c] )
// i.e. finish_tiles (gray -> C replicated channels + T.Normalize, same operation order: the fp32 values are bit-identical)
// followed by the patch gather of PatchEmbed.proj (vision_transformer_base.py:95-101,136-138), in ONE pass: the fp32
// [B,C,H,W] batch never exists, the step's input is the single-channel tile (2 bytes/pixel over PCIe / from HBM), and
// the patch GEMM reads what this kernel wrote.  8 pixels per thread: one 16-byte load (u16 / 16-bit) or two (fp32),
// one 16-byte store per channel.
struct TilePatchParams {
  const void* tiles;
  const float* bounds;   // [B,2] or nullptr
  void* patches;
  int src_kind;          // 0 fp32 [0,1], 1 uint16 (/65535), 2 fp16, 3 bf16
  int fp16_out;
  int B, C, H, W, P;
  float mean[4], stdv[4];
  int has_norm;
};
__global__ void __launch_bounds__(256) tiles_to_patches_kernel(TilePatchParams p) {
  const int wv = p.W >> 3;
  const long long total = (long long)p.B * p.H * wv;
  const int gw = p.W / p.P, gh = p.H / p.P;
  const int PP = p.P * p.P;
  const int Kdim = p.C * PP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xv = int(i % wv);
    const int y = int((i / wv) % p.H);
    const long long b = i / ((long long)wv * p.H);
    const long long off = (b * p.H + y) * (long long)p.W + 8 * xv;
    float g[8];
    if (p.src_kind == 0) {
      const float4 a0 = ldg_f4(reinterpret_cast<const float*>(p.tiles) + off), a1 = ldg_f4(reinterpret_cast<const float*>(p.tiles) + off + 4);
      g[0] = a0.x; g[1] = a0.y; g[2] = a0.z; g[3] = a0.w; g[4] = a1.x; g[5] = a1.y; g[6] = a1.z; g[7] = a1.w;
    } else {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.tiles) + off));
      const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (p.src_kind == 1) {   // _preprocess_image: astype(float32) / 65535
          g[2 * j] = __fdiv_rn((float)(w4[j] & 0xffffu), 65535.f);
          g[2 * j + 1] = __fdiv_rn((float)(w4[j] >> 16), 65535.f);
        } else {
          const float2 f = unpack16(w4[j], p.src_kind == 2);
          g[2 * j] = f.x;
          g[2 * j + 1] = f.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = adapt(g[j], p.bounds, b);
    const int x = 8 * xv;
    const int py = y / p.P, ky = y - py * p.P, px = x / p.P, kx = x - px * p.P;
    uint16_t* dst = reinterpret_cast<uint16_t*>(p.patches) + ((b * gh + py) * gw + px) * (long long)Kdim + ky * p.P + kx;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= p.C) break;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = p.has_norm ? __fdiv_rn(g[j] - p.mean[c], p.stdv[c]) : g[j];
      *reinterpret_cast<uint4*>(dst + c * PP) = make_uint4(pack16(o[0], o[1], p.fp16_out), pack16(o[2], o[3], p.fp16_out),
                                                           pack16(o[4], o[5], p.fp16_out), pack16(o[6], o[7], p.fp16_out));
    }
  }
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_resize_u16(const uint16_t* raw, float* gray, int32_t B, int32_t Hs, int32_t Ws, int32_t H, int32_t W,
                               void* stream) {
  VITK_CHECK_ARG(raw && gray && B > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0, "vitk_resize_u16: bad args");
  // cv2: inv_scale = dsize / ssize (double), scale = 1 / inv_scale
  const double sx = 1.0 / ((double)W / (double)Ws), sy = 1.0 / ((double)H / (double)Hs);
  const long long total = (long long)B * H * W;
  resize_u16_kernel<<<capped_grid(total, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(raw, gray, B, Hs, Ws, H, W, sx,
                                                                                                 sy);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_percentile_bounds(const float* x, int32_t B, int64_t n, float q_lo, float q_hi, float* bounds, void* stream) {
  VITK_CHECK_ARG(x && bounds && B > 0 && n > 0, "vitk_percentile_bounds: bad args");
  VITK_CHECK_ARG(q_lo >= 0.f && q_lo <= 1.f && q_hi >= 0.f && q_hi <= 1.f, "vitk_percentile_bounds: quantiles must lie in [0, 1]");
  const size_t smem = n <= SEL_SMEM_MAX_ELEMS ? (size_t)((n + 3) / 4 * 4) * sizeof(float) : 0;
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(percentile_bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(SEL_SMEM_MAX_ELEMS * sizeof(float))));
    configured = true;
  }
  percentile_bounds_kernel<<<B, SEL_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, (long long)n, q_lo, q_hi, bounds);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_finish_tiles(const float* gray, const float* bounds, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                                 const float* mean, const float* stdv, const int32_t* perm, int32_t cutmix, float lam,
                                 int32_t x1, int32_t y1, int32_t x2, int32_t y2, void* stream) {
  VITK_CHECK_ARG(gray && out && B > 0 && C >= 1 && C <= 4 && H > 0 && W > 0 && W % 4 == 0, "vitk_finish_tiles: bad args (1 <= C <= 4, W %% 4 == 0)");
  VITK_CHECK_ARG((mean == nullptr) == (stdv == nullptr), "vitk_finish_tiles: mean and std come together");
  VITK_CHECK_ARG(gray != out, "vitk_finish_tiles: in-place operation is not supported");
  FinishParams p;
  p.gray = gray; p.bounds = bounds; p.perm = perm; p.out = out;
  p.B = B; p.C = C; p.H = H; p.W = W;
  p.has_norm = mean != nullptr;
  for (int c = 0; c < 4; ++c) {
    p.mean[c] = (mean != nullptr && c < C) ? mean[c] : 0.f;      // HOST arrays (a handful of floats from the transform config)
    p.stdv[c] = (stdv != nullptr && c < C) ? stdv[c] : 1.f;
  }
  p.cutmix = cutmix; p.lam = lam;
  p.x1 = x1; p.y1 = y1; p.x2 = x2; p.y2 = y2;
  const long long total = (long long)B * H * (W / 4);
  finish_tiles_kernel<<<capped_grid(total, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_tiles_to_patches(const void* tiles, int32_t tiles_kind, const float* bounds, const float* mean,
                                     const float* stdv, void* patches, int32_t patches_dtype, int32_t B, int32_t C,
                                     int32_t H, int32_t W, int32_t P, void* stream) {
  VITK_CHECK_ARG(tiles && patches && B > 0 && C >= 1 && C <= 4, "vitk_tiles_to_patches: bad args (1 <= C <= 4)");
  VITK_CHECK_ARG(tiles_kind >= 0 && tiles_kind <= 3, "vitk_tiles_to_patches: tiles_kind 0 fp32, 1 uint16, 2 fp16, 3 bf16");
  VITK_CHECK_ARG(patches_dtype == VITK_BF16 || patches_dtype == VITK_FP16, "vitk_tiles_to_patches: patches must be bf16 or fp16");
  VITK_CHECK_ARG(P > 0 && P % 8 == 0 && H % P == 0 && W % P == 0,
                 "vitk_tiles_to_patches: need P %% 8 == 0 and H, W divisible by P (H=%d W=%d P=%d)", H, W, P);
  VITK_CHECK_ARG((mean == nullptr) == (stdv == nullptr), "vitk_tiles_to_patches: mean and std come together");
  TilePatchParams p;
  p.tiles = tiles; p.bounds = bounds; p.patches = patches;
  p.src_kind = tiles_kind; p.fp16_out = int(patches_dtype == VITK_FP16);
  p.B = B; p.C = C; p.H = H; p.W = W; p.P = P;
  p.has_norm = mean != nullptr;
  for (int c = 0; c < 4; ++c) {
    p.mean[c] = (mean != nullptr && c < C) ? mean[c] : 0.f;      // HOST arrays, as in vitk_finish_tiles
    p.stdv[c] = (stdv != nullptr && c < C) ? stdv[c] : 1.f;
  }
  const long long total = (long long)B * H * (W / 8);
  tiles_to_patches_kernel<<<capped_grid(total, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
