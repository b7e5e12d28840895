// bottleneck_tc.cu -- the 1x1 "bottleneck" convolution of a frozen DenseNet dense layer, fused with the eval-mode
// BatchNorm + ReLU before it and the (folded) BatchNorm + ReLU after it, on the 5th-gen tensor cores.
//
// torchvision densenet.py `_DenseLayer` (the teacher of the distillation step, lightning_modules.py:943-947, densenet.py:24-45):
//     t = relu2(norm2(conv1(relu1(norm1(cat(previous features))))))          conv1: 1x1, C -> 128, no bias
// In eval mode norm1 is a per-channel affine map over the first C channels of the block's NHWC concatenation buffer and
// norm2 folds into conv1 (weights scaled per output channel + a bias).  As a GEMM over pixels:
//     T[p, 0:128] = relu( relu(X[p, 0:C] * s1 + h1) @ W1'^T + b1' )           X: [P pixels, pitch], W1': [128, C]
// The round-1 executor ran this as vitk_affine_relu_nhwc (read C, write C channels of every pixel) + a cuDNN 1x1
// convolution (read C again): three passes over the growing concatenation per layer, 18 GB of the teacher's HBM traffic
// at batch 256.  Here the affine + ReLU is applied to the A tile IN SHARED MEMORY between its TMA load and the MMA that
// reads it, so every layer reads its C input channels exactly once and writes its 128 outputs once.
//
//   warp 0     TMA producer: A block [128 pixels x 64 channels] + W block [128 x 64] per stage, 4-stage ring
//   warp 1     MMA issuer (one elected thread): tcgen05.mma M=128 N=128 K=16 x 4 per block, fp32 accumulator in TMEM
//   warps 2-9  two threads per pixel row (32 channels each): transform the landed A block in place with packed 16-bit
//              fma.rn.relu (scale and shift rounded to the operand type once at start-up; the first version unpacked to
//              fp32 -- bit-identical to vitk_affine_relu_nhwc, but ~55 instructions per 16-byte cell made this pass, not HBM,
//              set the pace at 1.5 k cycles per k-block)
//   warps 10-13 tile epilogue (TMEM -> bias + ReLU -> 16-bit -> swizzled staging -> TMA store) of tile t while the other
//              warps already work on tile t+1 (two accumulator stages in TMEM)
// HBM-bound by construction: per k-block a CTA moves 16 KB of activations (W comes from L2) against 256 cycles of MMA.
#include <cudaTypedefs.h>

#include "tc_common.cuh"

namespace vitk {
namespace {

using namespace tc;

constexpr int BT_THREADS = 448;   // TMA warp, MMA warp, 8 transform warps, 4 epilogue warps
constexpr int BT_STAGES = 4;
constexpr int BT_N = 128;                 // bottleneck width (bn_size * growth_rate = 4 * 32 in every torchvision DenseNet)
constexpr int BT_KMAX = 2048;             // channels of the widest concatenation served (DenseNet169: 1664, 201: 1920)
constexpr int BT_STAGE_BYTES = 2 * 128 * 128;   // A block + W block, [128 rows][64 x 16-bit], 128B swizzle

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
template <bool H16>
__device__ __forceinline__ uint16_t to_bits16(float v) {
  if (H16) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
// max(x * s + h, 0) on a pair of 16-bit values (fma.rn.relu: the product-sum is exact, rounded once)
template <bool H16>
__device__ __forceinline__ uint32_t fma_relu2(uint32_t x, uint32_t s, uint32_t h) {
  uint32_t d;
  if (H16) asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(s), "r"(h));
  else asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(s), "r"(h));
  return d;
}

template <bool H16>
__global__ void __launch_bounds__(BT_THREADS, 1)
    dense_bottleneck_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                            const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ scale,
                            const float* __restrict__ shift, const float* __restrict__ bias, int C, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sStage = smem;                                         // BT_STAGES x (A 16 KB + W 16 KB)
  uint8_t* sOut = sStage + BT_STAGES * BT_STAGE_BYTES;            // 2 x [128][64 x 16-bit] output halves
  uint16_t* sScale = reinterpret_cast<uint16_t*>(sOut + 2 * 16384);   // [KP] norm1 scale in the operand's 16-bit type
  uint16_t* sShift = sScale + BT_KMAX;                               // [KP] norm1 shift
  float* sBias = reinterpret_cast<float*>(sShift + BT_KMAX);         // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + BT_N);
  uint64_t* full = bars;                     // [S] A and W block landed
  uint64_t* ready = bars + BT_STAGES;        // [S] A block transformed (256 arrivals)
  uint64_t* empty = bars + 2 * BT_STAGES;    // [S] the block's MMAs have completed
  uint64_t* tfull = bars + 3 * BT_STAGES;    // [2] accumulator of the tile complete
  uint64_t* tempty = tfull + 2;              // [2] accumulator read out (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (C + 63) >> 6;
  const int KP = nkb * 64;

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&tmX);
      prefetch_tmap(&tmW);
      prefetch_tmap(&tmOut);
      for (int s = 0; s < BT_STAGES; ++s) {
        mbar_init(full + s, 1);
        mbar_init(ready + s, 256);
        mbar_init(empty + s, 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull + a, 1);
        mbar_init(tempty + a, 128);
      }
      mbar_init_fence();
    }
    __syncwarp();
  } else if (warp == 1) {
    tmem_alloc<256>(tmem_slot);
  }
  // per-channel affine of norm1 (zero beyond C: the W columns there are zero-filled by TMA, the products vanish) and bias
  for (int i = threadIdx.x; i < KP; i += BT_THREADS) {
    sScale[i] = i < C ? to_bits16<H16>(__ldg(scale + i)) : uint16_t(0);
    sShift[i] = i < C ? to_bits16<H16>(__ldg(shift + i)) : uint16_t(0);
  }
  for (int i = threadIdx.x; i < BT_N; i += BT_THREADS) sBias[i] = __ldg(bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int g = 0;   // global k-block counter (ring position)
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % BT_STAGES;
          if (g >= BT_STAGES) mbar_wait(empty + s, ((g / BT_STAGES) - 1) & 1, 40);
          uint8_t* st = sStage + s * BT_STAGE_BYTES;
          mbar_expect_tx(full + s, BT_STAGE_BYTES);
          tma_load_3d(st, &tmX, full + s, kb * 64, tile * 128, 0);
          tma_load_3d(st + 16384, &tmW, full + s, kb * 64, 0, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = idesc_f16(BT_N, false, false, H16);
      int g = 0, t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int acc = t & 1;
        if (t >= 2) {
          mbar_wait(tempty + acc, ((t >> 1) - 1) & 1, 41);
          tc_fence_after();
        }
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % BT_STAGES;
          mbar_wait(ready + s, (g / BT_STAGES) & 1, 42);
          tc_fence_after();
          uint8_t* st = sStage + s * BT_STAGE_BYTES;
          const uint64_t adesc = smem_desc_kmajor(smem_u32(st));
          const uint64_t bdesc = smem_desc_kmajor(smem_u32(st + 16384));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tb + uint32_t(BT_N * acc), adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty + s);
        }
        umma_commit(tfull + acc);
      }
    }
    __syncwarp();
  } else if (warp < 10) {
    // ===================== transform: two threads per pixel row =====================
    const int row = ((warp - 2) & 3) * 32 + lane;
    const int half = (warp - 2) >> 2;              // 16-byte cells [4 * half, 4 * half + 4) of the row
    int g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const int s = g % BT_STAGES;
        mbar_wait(full + s, (g / BT_STAGES) & 1, 43);
        uint8_t* a = sStage + s * BT_STAGE_BYTES;
        const uint16_t* sc = sScale + kb * 64;
        const uint16_t* sh = sShift + kb * 64;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = 4 * half + cc;
          uint4* cell = reinterpret_cast<uint4*>(a + swz128(row, c));
          const uint4 in = *cell;
          const uint4 s4 = *reinterpret_cast<const uint4*>(sc + 8 * c);
          const uint4 h4 = *reinterpret_cast<const uint4*>(sh + 8 * c);
          uint4 o;   // packed fma + relu: two channels per instruction, one rounding of the exact x * s + h
          o.x = fma_relu2<H16>(in.x, s4.x, h4.x);
          o.y = fma_relu2<H16>(in.y, s4.y, h4.y);
          o.z = fma_relu2<H16>(in.z, s4.z, h4.z);
          o.w = fma_relu2<H16>(in.w, s4.w, h4.w);
          *cell = o;
        }
        fence_proxy_async();          // the generic-proxy writes above must be visible to the tensor core's smem reads
        mbar_arrive(ready + s);
      }
    }
  } else {
    // ===================== epilogue: T = relu(acc + b1') -> 16-bit -> swizzled staging -> TMA store =====================
    const int quad = warp & 3;                       // TMEM lane quadrant of this warp (warp id % 4)
    const int row = quad * 32 + lane;
    const uint32_t trow = tb + (uint32_t(quad * 32) << 16);
    const bool leader = warp == 10;
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      mbar_wait(tfull + acc, (t >> 1) & 1, 44);
      tc_fence_after();
      if (leader && elect_one()) tma_store_wait_read();     // the previous tile's store has finished reading the staging tiles
      named_bar(1, 128);
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        uint32_t v[32];
        tmem_ld32_nowait(trow + uint32_t(BT_N * acc + 32 * hh), v);
        tmem_ld_wait();
        uint8_t* dst = sOut + (hh >> 1) * 16384;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 b0 = *reinterpret_cast<const float4*>(sBias + 32 * hh + 8 * c);
          const float4 b1 = *reinterpret_cast<const float4*>(sBias + 32 * hh + 8 * c + 4);
          uint4 o;
          o.x = pack16(fmaxf(__uint_as_float(v[8 * c + 0]) + b0.x, 0.f), fmaxf(__uint_as_float(v[8 * c + 1]) + b0.y, 0.f), H16);
          o.y = pack16(fmaxf(__uint_as_float(v[8 * c + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(v[8 * c + 3]) + b0.w, 0.f), H16);
          o.z = pack16(fmaxf(__uint_as_float(v[8 * c + 4]) + b1.x, 0.f), fmaxf(__uint_as_float(v[8 * c + 5]) + b1.y, 0.f), H16);
          o.w = pack16(fmaxf(__uint_as_float(v[8 * c + 6]) + b1.z, 0.f), fmaxf(__uint_as_float(v[8 * c + 7]) + b1.w, 0.f), H16);
          *reinterpret_cast<uint4*>(dst + swz128(row, 4 * (hh & 1) + c)) = o;
        }
      }
      tc_fence_before();
      mbar_arrive(tempty + acc);
      fence_proxy_async();
      named_bar(1, 128);
      if (leader && elect_one()) {
        tma_store_3d(&tmOut, sOut, 0, tile * 128, 0);      // rows beyond the pixel count are clipped by the tensor map
        tma_store_3d(&tmOut, sOut + 16384, 64, tile * 128, 0);
        tma_store_commit();
      }
    }
    if (leader && elect_one()) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tb);
}

PFN_cuTensorMapEncodeTiled_v12000 bt_encode_fn() {
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

// 16-bit [rows][cols] matrix with a row pitch (elements), viewed as a 3-D map {cols, rows, 1}; box = 64 columns x 128 rows,
// 128B swizzle.  Out-of-range rows / columns are zero-filled on load and clipped on store.
int bt_tmap(CUtensorMap* tm, const void* base, int cols, long long rows, long long pitch, bool fp16) {
  auto fn = bt_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VITK_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)rows * (cuuint64_t)pitch * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(bottleneck) failed: CUresult %d (cols=%d rows=%lld pitch=%lld)", (int)r, cols, rows, pitch);
    return VITK_ERR_CUDA;
  }
  return VITK_OK;
}

template <bool H16>
int launch_bottleneck(const void* x, long long x_ld, const float* scale, const float* shift, const void* w, const float* bias,
                      void* out, long long pixels, int C, cudaStream_t st) {
  CUtensorMap tmX, tmW, tmOut;
  int rc;
  if ((rc = bt_tmap(&tmX, x, C, pixels, x_ld, H16)) != VITK_OK) return rc;
  if ((rc = bt_tmap(&tmW, w, C, BT_N, C, H16)) != VITK_OK) return rc;
  if ((rc = bt_tmap(&tmOut, out, BT_N, pixels, BT_N, H16)) != VITK_OK) return rc;
  const int smem = BT_STAGES * BT_STAGE_BYTES + 2 * 16384 + 2 * BT_KMAX * 2 + BT_N * 4 + (3 * BT_STAGES + 4) * 8 + 16 + 1024;
  auto kfn = dense_bottleneck_kernel<H16>;
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const long long tiles = (pixels + 127) / 128;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  kfn<<<grid, BT_THREADS, smem, st>>>(tmX, tmW, tmOut, scale, shift, bias, C, (int)tiles);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_dense_bottleneck(const void* x, int64_t x_ld, const float* scale, const float* shift, const void* w,
                                     const float* bias, void* out, int64_t pixels, int32_t C, int32_t dtype, void* stream) {
  VITK_CHECK_ARG(x && scale && shift && w && bias && out, "vitk_dense_bottleneck: null pointer");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_dense_bottleneck: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(pixels > 0 && pixels < (1ll << 31) - 128 && C >= 8 && C % 8 == 0 && C <= BT_KMAX && x_ld >= C && x_ld % 8 == 0,
                 "vitk_dense_bottleneck: C=%d (multiple of 8, <= %d), pitch=%lld (>= C, multiple of 8)", C, BT_KMAX, (long long)x_ld);
  VITK_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)out % 16 == 0),
                 "vitk_dense_bottleneck: 16-byte aligned pointers required");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == VITK_FP16 ? launch_bottleneck<true>(x, x_ld, scale, shift, w, bias, out, pixels, C, st)
                            : launch_bottleneck<false>(x, x_ld, scale, shift, w, bias, out, pixels, C, st);
}
