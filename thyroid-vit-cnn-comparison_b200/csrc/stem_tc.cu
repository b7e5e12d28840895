// stem_tc.cu -- the stem convolution of the frozen DenseNet teacher (7x7, stride 2, padding 3, 3 -> 64 channels, with the
// eval-mode norm0 folded in and relu0 applied) as an implicit GEMM on the 5th-gen tensor cores.
//
// torchvision densenet.py `features.conv0 / norm0 / relu0` (the teacher of the distillation step, lightning_modules.py:943-947).
// cuDNN runs a 3-channel convolution on a pre-tensor-core kernel (1.8 ms at batch 256); the first GEMM form of this executor
// wrote the patch matrix [B*112*112, 152] to HBM and read it back (vitk_im2col_rows + vitk_gemm: 1.3 ms, 2 GB of traffic for
// 77 MB of input).  Here the patch rows never leave the SM: a tile of 8 x 16 output pixels needs a 21-row x 37-pixel window
// of the NHWC input (4.6 KB, one TMA box whose out-of-image part is zero-filled -- the convolution's padding), eight warps
// expand it in shared memory into the K-major, 128B-swizzled A operand [128 pixels x 192] and one thread issues 12 MMAs
// against the filters, which stay resident in shared memory.
//
//   K layout: k = ky * 24 + 1 + (kx * 3 + c), zero filters at slots 0, 22, 23 of every ky segment and at k >= 168.  A segment is
//   then 24 consecutive input elements (48 bytes) that start one element left of the 7-wide tap: with the window's 16-byte
//   aligned origin that makes every segment 4-byte aligned in shared memory; the three neighbours meet zero weights.
//
//   warp 0      TMA producer: the filters once, then one input window per tile into an 8-deep ring
//   warp 1      MMA issuer: M=128 (pixels) N=64 (filters) K=16 x 12 per tile, two accumulator stages in TMEM
//   warps 2-9   A-tile builders: two threads per output pixel, 21 16-byte chunks per pixel row (4 x LDS.32 -> STS.128)
//   warps 10-13 epilogue: TMEM -> + bias -> ReLU -> 16-bit -> swizzled staging -> one 4-D TMA store [8 rows][16 px][64 ch]
// Per tile a CTA moves 4.6 KB in and 16 KB out against ~500 cycles of MMA: HBM-bound (B*112*112*64*2 bytes written).
#include <cudaTypedefs.h>

#include "tc_common.cuh"

namespace vitk {
namespace {

using namespace tc;

constexpr int ST_THREADS = 448;
constexpr int ST_STAGES = 8;
constexpr int ST_WIN_ROWS = 21;            // input rows of a tile: 2 * 8 + 5
constexpr int ST_WIN_ELEMS = 120;          // elements per window row: 6 + 6 * 15 + 24
constexpr int ST_WIN_BYTES = ST_WIN_ROWS * ST_WIN_ELEMS * 2;   // 5040: what one TMA box delivers
constexpr int ST_WIN_PITCH = 5120;         // ring slot (128-byte aligned)
constexpr int ST_KB = 3;                   // k-blocks of 64: K = 192
constexpr int ST_A_BYTES = ST_KB * 16384;  // [128 rows][64 x 16-bit] x 3
constexpr int ST_N = 64;
constexpr int ST_W_BYTES = ST_KB * ST_N * 128;
constexpr int ST_CHUNKS = 21;              // 16-byte chunks of real K per pixel row (7 segments x 48 bytes)

__device__ __forceinline__ void st_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

template <bool H16>
__global__ void __launch_bounds__(ST_THREADS, 1)
    stem_conv_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ bias, int tiles_x, int tiles_y,
                     int n_tiles, int relu) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;                                    // 2 x ST_A_BYTES
  uint8_t* sW = sA + 2 * ST_A_BYTES;                     // filters, K-major [64][64 x 16-bit] x 3
  uint8_t* sOut = sW + ST_W_BYTES;                       // [128 pixels][64 ch x 16-bit]
  uint8_t* sWin = sOut + 16384;                          // ST_STAGES input windows
  float* sBias = reinterpret_cast<float*>(sWin + ST_STAGES * ST_WIN_PITCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + ST_N);
  uint64_t* full = bars;                      // [S] window landed
  uint64_t* wfree = bars + ST_STAGES;         // [S] window expanded (256 arrivals)
  uint64_t* aready = bars + 2 * ST_STAGES;    // [2] A tile built (256 arrivals)
  uint64_t* afree = aready + 2;               // [2] the tile's MMAs have completed
  uint64_t* tfull = afree + 2;                // [2] accumulator complete
  uint64_t* tempty = tfull + 2;               // [2] accumulator read out (128 arrivals)
  uint64_t* wfull = tempty + 2;               // filters landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_img = tiles_x * tiles_y;

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&tmIn);
      prefetch_tmap(&tmW);
      prefetch_tmap(&tmOut);
      for (int s = 0; s < ST_STAGES; ++s) {
        mbar_init(full + s, 1);
        mbar_init(wfree + s, 256);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(aready + a, 256);
        mbar_init(afree + a, 1);
        mbar_init(tfull + a, 1);
        mbar_init(tempty + a, 128);
      }
      mbar_init(wfull, 1);
      mbar_init_fence();
    }
    __syncwarp();
  } else if (warp == 1) {
    tmem_alloc<128>(tmem_slot);
  }
  // both A buffers start as zeros: the builders never touch k >= 168 (and those filter columns are zero as well)
  for (int i = threadIdx.x; i < 2 * ST_A_BYTES / 16; i += ST_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < ST_N; i += ST_THREADS) sBias[i] = __ldg(bias + i);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(wfull, ST_W_BYTES);
      for (int kb = 0; kb < ST_KB; ++kb) tma_load_3d(sW + kb * ST_N * 128, &tmW, wfull, kb * 64, 0, 0);
      int t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int s = t % ST_STAGES;
        if (t >= ST_STAGES) mbar_wait(wfree + s, ((t / ST_STAGES) - 1) & 1, 50);
        const int b = tile / tiles_img, r = tile - b * tiles_img;
        const int ty = r / tiles_x, tx = r - ty * tiles_x;
        mbar_expect_tx(full + s, ST_WIN_BYTES);
        // window origin: input row 2*oy0 - 3, element 6*ox0 - 16 (the first tap sits at 6*ox0 - 9; a TMA box must start on a
        // 16-byte boundary of the row -- an odd element offset is an illegal instruction); outside the image: zeros
        tma_load_3d(sWin + s * ST_WIN_PITCH, &tmIn, full + s, 96 * tx - 16, 16 * ty - 3, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = idesc_f16(ST_N, false, false, H16);
      mbar_wait(wfull, 0, 51);
      tc_fence_after();
      int t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int ab = t & 1;
        if (t >= 2) {
          mbar_wait(tempty + ab, ((t >> 1) - 1) & 1, 52);
          tc_fence_after();
        }
        mbar_wait(aready + ab, (t >> 1) & 1, 53);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < ST_KB; ++kb) {
          const uint64_t adesc = smem_desc_kmajor(smem_u32(sA + ab * ST_A_BYTES + kb * 16384));
          const uint64_t bdesc = smem_desc_kmajor(smem_u32(sW + kb * ST_N * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tb + uint32_t(ST_N * ab), adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(afree + ab);
        umma_commit(tfull + ab);
      }
    }
    __syncwarp();
  } else if (warp < 10) {
    // ===================== A-tile builders: two threads per output pixel =====================
    const int tid = threadIdx.x - 64;           // 0..255
    const int row = tid & 127, half = tid >> 7;
    const int py = row >> 4, px = row & 15;     // pixel of the 8 x 16 tile; MMA row = py * 16 + px
    const uint32_t src0 = uint32_t(2 * py) * (ST_WIN_ELEMS * 2) + uint32_t(12 * px + 12);   // element 6*px + 6 of the window row
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int s = t % ST_STAGES, ab = t & 1;
      mbar_wait(full + s, (t / ST_STAGES) & 1, 54);
      if (t >= 2) mbar_wait(afree + ab, ((t >> 1) - 1) & 1, 55);
      const uint8_t* win = sWin + s * ST_WIN_PITCH + src0;
      uint8_t* a = sA + ab * ST_A_BYTES;
#pragma unroll
      for (int i = 0; i < (ST_CHUNKS + 1) / 2; ++i) {
        const int q = 2 * i + half;             // chunk of the pixel's K row: segment ky = q / 3, 16-byte part e = q % 3
        if (q < ST_CHUNKS) {
          const int ky = q / 3, e = q - 3 * ky;
          const uint32_t* src = reinterpret_cast<const uint32_t*>(win + ky * (ST_WIN_ELEMS * 2) + 16 * e);
          const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
          *reinterpret_cast<uint4*>(a + (q >> 3) * 16384 + swz128(row, q & 7)) = v;
        }
      }
      fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's shared-memory reads
      mbar_arrive(aready + ab);
      mbar_arrive(wfree + s);
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t trow = tb + (uint32_t(quad * 32) << 16);
    const bool leader = warp == 10;
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int ab = t & 1;
      mbar_wait(tfull + ab, (t >> 1) & 1, 56);
      tc_fence_after();
      if (leader && elect_one()) tma_store_wait_read();
      st_named_bar(1, 128);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32_nowait(trow + uint32_t(ST_N * ab + 32 * hh), v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 b0 = *reinterpret_cast<const float4*>(sBias + 32 * hh + 8 * c);
          const float4 b1 = *reinterpret_cast<const float4*>(sBias + 32 * hh + 8 * c + 4);
          float f[8] = {__uint_as_float(v[8 * c + 0]) + b0.x, __uint_as_float(v[8 * c + 1]) + b0.y, __uint_as_float(v[8 * c + 2]) + b0.z,
                        __uint_as_float(v[8 * c + 3]) + b0.w, __uint_as_float(v[8 * c + 4]) + b1.x, __uint_as_float(v[8 * c + 5]) + b1.y,
                        __uint_as_float(v[8 * c + 6]) + b1.z, __uint_as_float(v[8 * c + 7]) + b1.w};
          if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          uint4 o;
          o.x = pack16(f[0], f[1], H16);
          o.y = pack16(f[2], f[3], H16);
          o.z = pack16(f[4], f[5], H16);
          o.w = pack16(f[6], f[7], H16);
          *reinterpret_cast<uint4*>(sOut + swz128(row, 4 * hh + c)) = o;
        }
      }
      tc_fence_before();
      mbar_arrive(tempty + ab);
      fence_proxy_async();
      st_named_bar(1, 128);
      if (leader && elect_one()) {
        const int b = tile / tiles_img, r = tile - b * tiles_img;
        const int ty = r / tiles_x, tx = r - ty * tiles_x;
        tma_store_4d(&tmOut, sOut, 0, 16 * tx, 8 * ty, b);    // pixels beyond the output image are clipped by the tensor map
        tma_store_commit();
      }
    }
    if (leader && elect_one()) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tb);
}

PFN_cuTensorMapEncodeTiled_v12000 st_encode_fn() {
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

int st_encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
              CUtensorMapSwizzle swz, bool fp16, const char* what) {
  auto fn = st_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VITK_ERR_CUDA;
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(stem %s) failed: CUresult %d", what, (int)r);
    return VITK_ERR_CUDA;
  }
  return VITK_OK;
}

template <bool H16>
int launch_stem(const void* x, const void* w, const float* bias, void* out, int B, int H, int W, int relu, cudaStream_t st) {
  const int OH = H / 2, OW = W / 2;
  CUtensorMap tmIn, tmW, tmOut;
  int rc;
  {  // input [B][H][W*3] 16-bit, box = 120 elements x 21 rows, dense (unswizzled) in shared memory
    cuuint64_t dims[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)W * 6, (cuuint64_t)H * W * 6};
    cuuint32_t box[3] = {ST_WIN_ELEMS, ST_WIN_ROWS, 1};
    if ((rc = st_encode(&tmIn, x, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, H16, "input")) != VITK_OK) return rc;
  }
  {  // filters [64][192] 16-bit, box = 64 x 64, 128B swizzle (K-major B operand)
    cuuint64_t dims[3] = {(cuuint64_t)(ST_KB * 64), (cuuint64_t)ST_N, 1};
    cuuint64_t strides[2] = {(cuuint64_t)(ST_KB * 64) * 2, (cuuint64_t)ST_N * (ST_KB * 64) * 2};
    cuuint32_t box[3] = {64, ST_N, 1};
    if ((rc = st_encode(&tmW, w, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, H16, "filters")) != VITK_OK) return rc;
  }
  {  // output [B][OH][OW][64] 16-bit, box = 64 ch x 16 px x 8 rows
    cuuint64_t dims[4] = {(cuuint64_t)ST_N, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ST_N * 2, (cuuint64_t)OW * ST_N * 2, (cuuint64_t)OH * OW * ST_N * 2};
    cuuint32_t box[4] = {ST_N, 16, 8, 1};
    if ((rc = st_encode(&tmOut, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, H16, "output")) != VITK_OK) return rc;
  }
  const int smem = 2 * ST_A_BYTES + ST_W_BYTES + 16384 + ST_STAGES * ST_WIN_PITCH + ST_N * 4 + (2 * ST_STAGES + 9) * 8 + 16 + 1024;
  auto kfn = stem_conv_kernel<H16>;
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles_x = (OW + 15) / 16, tiles_y = (OH + 7) / 8;
  const long long tiles = (long long)B * tiles_x * tiles_y;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  kfn<<<grid, ST_THREADS, smem, st>>>(tmIn, tmW, tmOut, bias, tiles_x, tiles_y, (int)tiles, relu);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_stem_conv7(const void* x, const void* w, const float* bias, void* out, int32_t B, int32_t H, int32_t W,
                               int32_t relu, int32_t dtype, void* stream) {
  VITK_CHECK_ARG(x && w && bias && out, "vitk_stem_conv7: null pointer");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_stem_conv7: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && H >= 8 && W >= 8 && H % 2 == 0 && W % 8 == 0 && (long long)B * (H / 2) * (W / 2) < (1ll << 31) - 128,
                 "vitk_stem_conv7: bad shape B=%d H=%d W=%d (H even, W %% 8 == 0: 16-byte input rows)", B, H, W);
  VITK_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)out % 16 == 0),
                 "vitk_stem_conv7: 16-byte aligned pointers required");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == VITK_FP16 ? launch_stem<true>(x, w, bias, out, B, H, W, relu, st)
                            : launch_stem<false>(x, w, bias, out, B, H, W, relu, st);
}
