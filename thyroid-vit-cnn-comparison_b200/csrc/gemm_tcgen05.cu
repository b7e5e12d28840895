// gemm_tcgen05.cu -- bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring.
//
//   D[M,N] = A[M,K] * B[N,K]^T   with a fused epilogue (bias / residual / GELU / dGELU /
//                                  split-K atomic accumulate / patch->token scatter + pos_embed)
//
// Replaces the nn.Linear / Conv2d(k=16,s=16) call sites of the reference encoder block
// (vision_transformer_base.py:95-101, :166-168, :212-222) in forward, dgrad and wgrad.
// Both operands can be read either K-major (row-major [rows,K]) or MN-major (row-major
// [K,rows]) straight from the tensors autograd already holds, so backward needs no transposes:
//   forward  Y  = X  * W^T      A = X  [M,K]  K-major     B = W  [N,K]  K-major
//   dgrad    dX = dY * W        A = dY [M,N]  K-major     B = W  [N,K]  read MN-major
//   wgrad    dW = dY^T * X      A = dY [M,N]  MN-major    B = X  [M,K]  MN-major (split-K)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quadrant each).  One 128 x BN output tile per CTA; the
// shallow-pipeline variants fit two CTAs per SM so one CTA's epilogue overlaps the other's
// mainloop (DeiT-tiny GEMMs have K = 192: three k-blocks, epilogue-dominated).
#include <cudaTypedefs.h>

#include "vitk_common.cuh"

namespace vitk {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

struct GemmParams {
  int M, N, K;
  int num_kblocks;
  int kblocks_per_split;
  int epilogue, out_dtype, aux_dtype;
  uint32_t idesc;
  float alpha;
  const float* alpha_dev;
  const float* bias;
  const float* residual;
  long long ldr;
  void* out;
  long long ldo;
  void* out2;
  long long ldo2;
  const void* aux;
  long long ldaux;
  int rows_per_img, tokens_per_img, prefix;
  const float* pos;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ull) {  // 2 s
        printf("vitk gemm: mbarrier wait timeout tag=%d block=(%d,%d,%d) thread=%d parity=%u\n", tag,
               blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, parity);
        __trap();
      }
    }
  }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128B swizzle.
//   K-major : rows of 128 B (64 bf16 of K), 8-row groups 1024 B apart        -> SBO = 1024
//   MN-major: rows of 128 B (64 bf16 of M/N) per k, 8-k groups 1024 B apart   -> SBO = 1024,
//             next 64-wide M/N block one whole TMA box (64 k * 128 B) further -> LBO = 8192
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  constexpr uint64_t lbo = MN_MAJOR ? (8192u >> 4) : 1u;
  constexpr uint64_t sbo = 1024u >> 4;
  return static_cast<uint64_t>((saddr >> 4) & 0x3fffu) | (lbo << 16) | (sbo << 32) | (1ull << 46) |
         (2ull << 61);
}

// Instruction descriptor for kind::f16: {bf16|fp16} x {bf16|fp16} -> fp32, M = 128, N = bn.
// a_format / b_format: 0 = F16, 1 = BF16 (may differ between A and B).
inline uint32_t make_idesc(int bn, bool a_mn, bool b_mn, bool a_fp16, bool b_fp16) {
  return (1u << 4) | (uint32_t(a_fp16 ? 0 : 1) << 7) | (uint32_t(b_fp16 ? 0 : 1) << 10) | (uint32_t(a_mn) << 15) |
         (uint32_t(b_mn) << 16) | (uint32_t(bn >> 3) << 17) | (uint32_t(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}

template <int BN>
constexpr int tmem_cols() {
  return BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
}

// ------------------------------------------------------------------ the kernel
template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const GemmParams p) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_BYTES = BN * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = tmem_cols<BN>();

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * BN;
  const int kb_begin = blockIdx.z * p.kblocks_per_split;
  const int kb_end = min(p.num_kblocks, kb_begin + p.kblocks_per_split);
  const int nk = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1, 1);
      if (lane == 0) {
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int kc = (kb_begin + kb) * BLOCK_K;
        uint8_t* a_dst = sA + s * A_BYTES;
        uint8_t* b_dst = sB + s * B_BYTES;
        if (!A_MN) {
          tma_load_2d(a_dst, &tmA, &full_bar[s], kc, m0);
        } else {
#pragma unroll
          for (int i = 0; i < BLOCK_M / 64; ++i) tma_load_2d(a_dst + i * 8192, &tmA, &full_bar[s], m0 + 64 * i, kc);
        }
        if (!B_MN) {
          tma_load_2d(b_dst, &tmB, &full_bar[s], kc, n0);
        } else {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) tma_load_2d(b_dst + i * 8192, &tmB, &full_bar[s], n0 + 64 * i, kc);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = p.idesc;
    constexpr uint32_t a_adv = A_MN ? (2048u >> 4) : (32u >> 4);  // desc.lo step per UMMA_K
    constexpr uint32_t b_adv = B_MN ? (2048u >> 4) : (32u >> 4);
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full_bar[s], ph, 2);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_smem_desc<A_MN>(smem_u32(sA + s * A_BYTES));
        const uint64_t bdesc = make_smem_desc<B_MN>(smem_u32(sB + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          umma_bf16(tmem_base, adesc + uint64_t(k * a_adv), bdesc + uint64_t(k * b_adv), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (kb == nk - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = m0 + quad * 32 + lane;
    const bool row_ok = row < p.M;
    if (nk > 0) {
      mbar_wait(tmem_full_bar, 0, 3);
      tcgen05_fence_after();
    }
    long long orow = row;
    const float* pos_row = nullptr;
    if (p.epilogue == VITK_EPI_TOKENS && row_ok) {
      const int b = row / p.rows_per_img;
      const int pi = row - b * p.rows_per_img;
      orow = (long long)b * p.tokens_per_img + p.prefix + pi;
      pos_row = p.pos + (long long)(p.prefix + pi) * p.N;
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      const int col0 = n0 + c0;
      if (col0 >= p.N) break;  // warp-uniform
      uint32_t v[32];
      if (nk > 0) {
        tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(c0), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (!row_ok) continue;
      float f[32];
      const float alpha = p.alpha_dev != nullptr ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * alpha;
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (col0 + j < p.N) {
            const float4 b4 = ldg_f4(p.bias + col0 + j);
            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
          }
        }
      }
      switch (p.epilogue) {
        case VITK_EPI_STORE:
        case VITK_EPI_TOKENS: {
          if (p.epilogue == VITK_EPI_TOKENS) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < p.N) {
                const float4 q4 = ldg_f4(pos_row + col0 + j);
                f[j] += q4.x; f[j + 1] += q4.y; f[j + 2] += q4.z; f[j + 3] += q4.w;
              }
            }
          }
          if (p.residual != nullptr) {
            const float* r = p.residual + orow * p.ldr + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < p.N) {
                const float4 r4 = *reinterpret_cast<const float4*>(r + j);
                f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
              }
            }
          }
          if (p.out_dtype == VITK_FP32) {
            float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (col0 + j < p.N) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            const bool h = p.out_dtype == VITK_FP16;
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (col0 + j < p.N)
                st_global_v4(o + j, pack16(f[j], f[j + 1], h), pack16(f[j + 2], f[j + 3], h),
                             pack16(f[j + 4], f[j + 5], h), pack16(f[j + 6], f[j + 7], h));
          }
        } break;
        case VITK_EPI_GELU: {
          const bool h = p.out_dtype == VITK_FP16;
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col0;
          __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(p.out2) + orow * p.ldo2 + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < p.N) {
              st_global_v4(o + j, pack16(f[j], f[j + 1], h), pack16(f[j + 2], f[j + 3], h),
                           pack16(f[j + 4], f[j + 5], h), pack16(f[j + 6], f[j + 7], h));
              float g[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) g[t] = gelu_erf(f[j + t]);
              st_global_v4(o2 + j, pack16(g[0], g[1], h), pack16(g[2], g[3], h), pack16(g[4], g[5], h),
                           pack16(g[6], g[7], h));
            }
          }
        } break;
        case VITK_EPI_DGELU: {
          const __nv_bfloat16* ax = reinterpret_cast<const __nv_bfloat16*>(p.aux) + orow * p.ldaux + col0;
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < p.N) {
              const uint4 a4 = ldg_u4(ax + j);
              const bool ah = p.aux_dtype == VITK_FP16;
              const float2 a0 = unpack16(a4.x, ah), a1 = unpack16(a4.y, ah), a2 = unpack16(a4.z, ah),
                           a3 = unpack16(a4.w, ah);
              const float g0 = f[j] * gelu_erf_grad(a0.x), g1 = f[j + 1] * gelu_erf_grad(a0.y);
              const float g2 = f[j + 2] * gelu_erf_grad(a1.x), g3 = f[j + 3] * gelu_erf_grad(a1.y);
              const float g4 = f[j + 4] * gelu_erf_grad(a2.x), g5 = f[j + 5] * gelu_erf_grad(a2.y);
              const float g6 = f[j + 6] * gelu_erf_grad(a3.x), g7 = f[j + 7] * gelu_erf_grad(a3.y);
              const bool h = p.out_dtype == VITK_FP16;
              st_global_v4(o + j, pack16(g0, g1, h), pack16(g2, g3, h), pack16(g4, g5, h), pack16(g6, g7, h));
            }
          }
        } break;
        case VITK_EPI_ATOMIC_ADD: {
          float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) atomicAdd(o + j, f[j]);
        } break;
        default:
          break;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

// 2-D bf16 tensor map over a row-major [outer, inner] matrix, 128B swizzle, zero OOB fill.
int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                 uint32_t box_inner, uint32_t box_outer, bool fp16) {
  auto fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VITK_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (base=%p inner=%llu outer=%llu pitch=%llu box=%ux%u)",
              (int)r, base, (unsigned long long)inner, (unsigned long long)outer,
              (unsigned long long)pitch_elems, box_inner, box_outer);
    return VITK_ERR_CUDA;
  }
  return VITK_OK;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, dim3 grid, cudaStream_t st) {
  constexpr int SMEM = STAGES * (BLOCK_M * BLOCK_K * 2 + BN * BLOCK_K * 2) + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool configured = false;
  auto kfn = gemm_tcgen05_kernel<BN, STAGES, A_MN, B_MN>;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  kfn<<<grid, GEMM_THREADS, SMEM, st>>>(tmA, tmB, p);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

template <bool A_MN, bool B_MN>
int dispatch_gemm(int bn, bool deep, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, dim3 grid,
                  cudaStream_t st) {
  switch (bn) {
    case 64:
      return deep ? launch_gemm<64, 8, A_MN, B_MN>(tmA, tmB, p, grid, st)
                  : launch_gemm<64, 4, A_MN, B_MN>(tmA, tmB, p, grid, st);
    case 128:
      return deep ? launch_gemm<128, 6, A_MN, B_MN>(tmA, tmB, p, grid, st)
                  : launch_gemm<128, 3, A_MN, B_MN>(tmA, tmB, p, grid, st);
    case 192:
      return deep ? launch_gemm<192, 5, A_MN, B_MN>(tmA, tmB, p, grid, st)
                  : launch_gemm<192, 2, A_MN, B_MN>(tmA, tmB, p, grid, st);
    case 256:
      return deep ? launch_gemm<256, 4, A_MN, B_MN>(tmA, tmB, p, grid, st)
                  : launch_gemm<256, 2, A_MN, B_MN>(tmA, tmB, p, grid, st);
    default:
      set_error("unsupported BLOCK_N %d", bn);
      return VITK_ERR_UNSUPPORTED;
  }
}

// Pick the N tile: the widest of {256,192,128,64} that tiles N with the least padding.
int pick_bn(int N) {
  const int cands[4] = {256, 192, 128, 64};
  int best = 64;
  long best_cost = -1;
  for (int c : cands) {
    const long tiles = (N + c - 1) / c;
    const long cost = tiles * c;  // padded width; ties -> wider tile (listed first)
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = c;
    }
  }
  return best;
}

}  // namespace

}  // namespace vitk

using namespace vitk;

extern "C" int vitk_gemm(const vitk_gemm_args* a, void* stream) {
  VITK_CHECK_ARG(a != nullptr, "vitk_gemm: null args");
  VITK_CHECK_ARG(a->A && a->B && a->out, "vitk_gemm: null operand");
  VITK_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "vitk_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  VITK_CHECK_ARG(a->N % 8 == 0, "vitk_gemm: N=%d must be a multiple of 8", a->N);
  VITK_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0, "vitk_gemm: lda/ldb must be multiples of 8 (16-byte TMA pitch)");
  VITK_CHECK_ARG((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0,
                 "vitk_gemm: operands must be 16-byte aligned");
  VITK_CHECK_ARG(a->split_k >= 1, "vitk_gemm: split_k must be >= 1");
  VITK_CHECK_ARG(a->split_k == 1 || a->epilogue == VITK_EPI_ATOMIC_ADD,
                 "vitk_gemm: split_k > 1 needs VITK_EPI_ATOMIC_ADD");
  VITK_CHECK_ARG(a->epilogue >= 0 && a->epilogue <= VITK_EPI_TOKENS, "vitk_gemm: bad epilogue %d", a->epilogue);
  VITK_CHECK_ARG(a->out_dtype >= VITK_BF16 && a->out_dtype <= VITK_FP16, "vitk_gemm: bad out_dtype %d", a->out_dtype);
  VITK_CHECK_ARG((a->a_dtype == VITK_BF16 || a->a_dtype == VITK_FP16) && (a->b_dtype == VITK_BF16 || a->b_dtype == VITK_FP16),
                 "vitk_gemm: operands must be bf16 or fp16");
  VITK_CHECK_ARG(a->a_dtype == a->b_dtype,
                 "vitk_gemm: A and B must share one element type (tcgen05 kind::f16 with mixed fp16/bf16 operands is an "
                 "illegal instruction on sm_100a)");
  const bool out_fp32 = a->out_dtype == VITK_FP32;
  if (a->epilogue == VITK_EPI_GELU) VITK_CHECK_ARG(a->out2 != nullptr && !out_fp32, "GELU epilogue needs 16-bit out and out2");
  if (a->epilogue == VITK_EPI_DGELU)
    VITK_CHECK_ARG(a->aux != nullptr && !out_fp32 && (a->aux_dtype == VITK_BF16 || a->aux_dtype == VITK_FP16),
                   "DGELU epilogue needs a 16-bit aux and 16-bit out");
  if (a->epilogue == VITK_EPI_ATOMIC_ADD) VITK_CHECK_ARG(out_fp32, "ATOMIC_ADD epilogue needs fp32 out");
  if (a->epilogue == VITK_EPI_TOKENS)
    VITK_CHECK_ARG(a->pos != nullptr && a->rows_per_img > 0 && a->tokens_per_img >= a->rows_per_img + a->prefix,
                   "TOKENS epilogue needs pos / rows_per_img / tokens_per_img");
  const int vec = out_fp32 ? 4 : 8;
  VITK_CHECK_ARG(a->ldo % vec == 0, "vitk_gemm: ldo must keep rows 16-byte aligned");

  const int bn = pick_bn(a->N);
  const int num_kblocks = (a->K + BLOCK_K - 1) / BLOCK_K;
  int kpb = (num_kblocks + a->split_k - 1) / a->split_k;
  const int splits = (num_kblocks + kpb - 1) / kpb;

  GemmParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_kblocks = num_kblocks;
  p.kblocks_per_split = kpb;
  p.epilogue = a->epilogue; p.out_dtype = a->out_dtype; p.aux_dtype = a->aux_dtype;
  p.idesc = make_idesc(bn, a->a_mn_major != 0, a->b_mn_major != 0, a->a_dtype == VITK_FP16, a->b_dtype == VITK_FP16);
  p.alpha = a->alpha; p.alpha_dev = a->alpha_dev;
  p.bias = a->bias; p.residual = a->residual; p.ldr = a->ldr;
  p.out = a->out; p.ldo = a->ldo; p.out2 = a->out2; p.ldo2 = a->ldo2;
  p.aux = a->aux; p.ldaux = a->ldaux;
  p.rows_per_img = a->rows_per_img; p.tokens_per_img = a->tokens_per_img; p.prefix = a->prefix; p.pos = a->pos;

  CUtensorMap tmA, tmB;
  int rc;
  if (!a->a_mn_major) rc = make_tmap_2d(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BLOCK_K, BLOCK_M, a->a_dtype == VITK_FP16);
  else                rc = make_tmap_2d(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BLOCK_K, a->a_dtype == VITK_FP16);
  if (rc != VITK_OK) return rc;
  if (!a->b_mn_major) rc = make_tmap_2d(&tmB, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BLOCK_K, (uint32_t)bn, a->b_dtype == VITK_FP16);
  else                rc = make_tmap_2d(&tmB, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BLOCK_K, a->b_dtype == VITK_FP16);
  if (rc != VITK_OK) return rc;

  dim3 grid((a->M + BLOCK_M - 1) / BLOCK_M, (a->N + bn - 1) / bn, splits);
  const bool deep = kpb >= 6;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!a->a_mn_major && !a->b_mn_major) return dispatch_gemm<false, false>(bn, deep, tmA, tmB, p, grid, st);
  if (!a->a_mn_major && a->b_mn_major) return dispatch_gemm<false, true>(bn, deep, tmA, tmB, p, grid, st);
  if (a->a_mn_major && a->b_mn_major) return dispatch_gemm<true, true>(bn, deep, tmA, tmB, p, grid, st);
  return dispatch_gemm<true, false>(bn, deep, tmA, tmB, p, grid, st);
}
