// gemm_tcgen05.cu -- 16-bit GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
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
// Structure (v2): PERSISTENT, one CTA per SM, 64 + 32*EPI_WARPS threads:
//   warp 0      TMA producer   -- runs ahead over tiles, STAGES-deep smem ring (full/empty mbarriers)
//   warp 1      MMA issuer     -- one elected lane issues tcgen05.mma; owns the TMEM allocation
//   warps 2..13 epilogue       -- three warps per TMEM lane quadrant (interleaved 32-column chunks): the exact-erf
//                                 GELU / dGELU epilogues are ALU work that needs the extra warps to hide latency
// TMEM holds TWO accumulator stages (2 x BN fp32 columns), so the MMAs of tile i+1 overlap the epilogue of
// tile i.  The epilogue transposes each 32x32 fp32 chunk through shared memory so that every global access
// of a warp covers whole 64/128-byte row segments (residual reads, 16-bit / fp32 stores, vector reductions).
// DeiT-tiny GEMMs (K = 192) are HBM-bound: what matters is bytes in flight (the ring) and coalescing;
// ViT-B GEMMs (K = 768/3072) are tensor-bound: what matters is that the MMA warp never waits for the epilogue.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "vitk_common.cuh"

#ifdef VITK_GEMM_KNOBS
#include <stdlib.h>
#define KNOB(x) ((p.knobs & (x)) != 0)
// per-tile timeline of CTA 0 (profiling build only): [role][tile][event] SM clock stamps
__device__ long long g_vitk_dbg[3 * 64 * 4];
#define DBG_STAMP(role, t, ev)                                                              \
  do {                                                                                      \
    if (KNOB(64) && blockIdx.x == 0 && (t) < 64) g_vitk_dbg[((role) * 64 + (t)) * 4 + (ev)] = clock64(); \
  } while (0)
__device__ long long g_vitk_dbg2[8 * 4 * 8];  // [tile][unit][event] of epilogue warp 2
#define DBG_UNIT(t, u, ev)                                                                                   \
  do {                                                                                                       \
    if (KNOB(64) && blockIdx.x == 0 && warp == 2 && lane == 0 && (t) < 8 && (u) < 4) g_vitk_dbg2[((t) * 4 + (u)) * 8 + (ev)] = clock64(); \
  } while (0)
// per-CTA wall-clock stamps (globaltimer ns): 0 kernel entry, 1 set-up done, 2 last epilogue done, 3 exit
__device__ unsigned long long g_vitk_dbg3[256 * 4];
#define DBG_G(ev)                                                                  \
  do {                                                                             \
    if (KNOB(64) && blockIdx.x < 256) {                                            \
      unsigned long long t_;                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                       \
      g_vitk_dbg3[blockIdx.x * 4 + (ev)] = t_;                                     \
    }                                                                              \
  } while (0)
#else
#define KNOB(x) false
#define DBG_G(ev) \
  do {            \
  } while (0)
#define DBG_UNIT(t, u, ev) \
  do {                     \
  } while (0)
#define DBG_STAMP(role, t, ev) \
  do {                         \
  } while (0)
#endif

namespace vitk {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 12;  // three warps per TMEM lane quadrant, interleaved over the 32-column chunks
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_PITCH = 20;  // floats per staged row (16 + 4 pad: conflict-free 128-bit accesses)
constexpr int EPI_STAGE_FLOATS = 1024;  // per warp: 32 x EPI_PITCH floats (fp32 path) or two 2 KB TMA-store buffers (16-bit path)
static_assert(EPI_STAGE_FLOATS >= 32 * EPI_PITCH, "staging buffer too small");
constexpr int ONES_OFFSET = 2816;   // floats: the last 2 KB of the bias region hold the ones tile of the column-sum MMA
constexpr int MAX_BIAS_SMEM = 3328;  // floats (13 KB): bias of every supported layer (N <= 3072, rounded up to the N tile)

struct GemmParams {
  int M, N, K;
  int num_kblocks;
  int kblocks_per_split;
  int num_m_tiles, num_n_tiles, num_splits;
  int step_split, step_mt, step_nt;  // mixed-radix digits of the grid size over (split, m tile, n tile)
  const float* row_scale;  // optional [M]: out = residual + row_scale[m] * (acc*alpha + bias)  (stochastic depth)
  DropSpec drop;           // optional nn.Dropout on the linear's output (after GELU for the GELU epilogue); seed == nullptr: off
  float* colsum;     // optional [M]: += alpha * sum_k A[m, k], from one extra N=16 MMA per k-step against a tile of ones
  int acc_stride;    // TMEM columns per accumulator stage (BN, or BN + 16 with colsum)
  int nacc;          // accumulator stages: 2 (MMAs of tile i+1 overlap epilogue i), or 1 when 2 x acc_stride > 512 columns
  int tmem_cols;     // TMEM allocation (power of two >= 2 * acc_stride)
  uint32_t idesc_ones;
  int epilogue, out_dtype, aux_dtype;
  int tma_epi;     // STORE / GELU / DGELU epilogues: row-per-lane math, 2 KB swizzled staging units, TMA loads (residual / saved
                   // derivative) and TMA stores -- no per-thread global access in the epilogue
  int knobs;  // profiling only (-DVITK_GEMM_KNOBS build, tools/build_dbg.py): 1 no TMA stores, 2 no epilogue math/stores,
              // 4 B operand loaded for the first tile of a CTA only, 8 A operand likewise, 16 epilogue does not even read TMEM,
              // 32 no MMAs issued (results are then wrong)
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
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.  The budget is 8 s of
// ACCUMULATED waiting, each sample clamped to 1 ms: %globaltimer is a wall clock, and a step of that clock must not be
// mistaken for a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long prev = 0, waited = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prev));
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      const unsigned long long d = now - prev;
      prev = now;
      waited += d < 1000000ull ? d : 1000000ull;
      if (waited > 8000000000ull) {
        printf("vitk gemm: mbarrier wait timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
        __trap();
      }
    }
  }
}

// ---- CTA pair (cluster of 2, tcgen05 cta_group::2): one 256 x BN tile per pair, each CTA holds 128 accumulator rows and
// stages its own A rows plus HALF of the B tile; the leader CTA issues the MMAs for both.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is counted on a barrier that may live in the peer CTA (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrives (once all earlier MMAs of this thread are complete) on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// true in exactly one (converged) lane of the warp; lets ptxas issue the uniform-datapath TMA / MMA instructions
// without the per-instruction "waterfall" loops it emits around `if (lane == 0)`
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// TMA reduction: global[box] += smem[box] (fp32 add performed at the L2, one bulk request instead of 512 per-lane REDs)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128B swizzle.
//   K-major : rows of 128 B (64 elements of K), 8-row groups 1024 B apart       -> SBO = 1024
//   MN-major: rows of 128 B (64 elements of M/N) per k, 8-k groups 1024 B apart  -> SBO = 1024,
//             next 64-wide M/N block one whole TMA box (64 k * 128 B) further    -> LBO = 8192
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  constexpr uint64_t lbo = MN_MAJOR ? (8192u >> 4) : 1u;
  constexpr uint64_t sbo = 1024u >> 4;
  return static_cast<uint64_t>((saddr >> 4) & 0x3fffu) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}

// Instruction descriptor for kind::f16: {bf16|fp16} x same -> fp32, M = 128, N = bn.  a/b_format: 0 = F16, 1 = BF16.
inline uint32_t make_idesc(int bn, bool a_mn, bool b_mn, bool a_fp16, bool b_fp16, int m = BLOCK_M) {
  return (1u << 4) | (uint32_t(a_fp16 ? 0 : 1) << 7) | (uint32_t(b_fp16 ? 0 : 1) << 10) | (uint32_t(a_mn) << 15) |
         (uint32_t(b_mn) << 16) | (uint32_t(bn >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

constexpr int tmem_cols_for(int bn) {
  return 2 * bn <= 32 ? 32 : 2 * bn <= 64 ? 64 : 2 * bn <= 128 ? 128 : 2 * bn <= 256 ? 256 : 512;
}

// ------------------------------------------------------------------ epilogue on one float4 (row, 4 consecutive columns)
// Split in two phases so that a warp first ISSUES the global loads of all its rows (residual / pos_embed / saved
// pre-activation) and only then does arithmetic and stores: `out` may alias `residual`, so the compiler cannot hoist
// those loads across the stores by itself, and a serialised load->math->store chain exposes HBM latency per row.
__device__ __forceinline__ long long epilogue_out_row(const GemmParams& p, int row) {
  if (p.epilogue != VITK_EPI_TOKENS) return row;
  const int b = row / p.rows_per_img;
  return (long long)b * p.tokens_per_img + p.prefix + (row - b * p.rows_per_img);
}
__device__ __forceinline__ float4 epilogue_load(const GemmParams& p, int row, long long orow, int col) {
  float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.epilogue == VITK_EPI_TOKENS) {
    const int b = row / p.rows_per_img;
    e = ldg_f4(p.pos + (long long)(p.prefix + row - b * p.rows_per_img) * p.N + col);
  } else if (p.epilogue == VITK_EPI_STORE) {
    if (p.residual != nullptr) e = *reinterpret_cast<const float4*>(p.residual + orow * p.ldr + col);
  } else if (p.epilogue == VITK_EPI_DGELU) {
    const uint2 a = ldg_u2(reinterpret_cast<const __nv_bfloat16*>(p.aux) + orow * p.ldaux + col);
    e.x = __uint_as_float(a.x);
    e.y = __uint_as_float(a.y);
  }
  return e;
}
template <bool DROP>
__device__ __forceinline__ void epilogue_apply(const GemmParams& p, float4 f, const float4& e, long long orow, int col, float alpha,
                                               const float4& bias4, unsigned long long dseed) {
  f.x = fmaf(f.x, alpha, bias4.x); f.y = fmaf(f.y, alpha, bias4.y); f.z = fmaf(f.z, alpha, bias4.z); f.w = fmaf(f.w, alpha, bias4.w);
  switch (p.epilogue) {
    case VITK_EPI_TOKENS:
    case VITK_EPI_STORE: {
      f.x += e.x; f.y += e.y; f.z += e.z; f.w += e.w;
      if (DROP && p.drop.seed != nullptr) {  // pos_drop: dropout of (patch embedding + pos_embed); host admits it for TOKENS only
        const float4 m = drop_factors4(p.drop, dseed, ((unsigned long long)orow * p.N + col) >> 2);
        f.x *= m.x; f.y *= m.y; f.z *= m.z; f.w *= m.w;
      }
      if (p.out_dtype == VITK_FP32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + col) = f;
      } else {
        const bool h = p.out_dtype == VITK_FP16;
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col) =
            make_uint2(pack16(f.x, f.y, h), pack16(f.z, f.w, h));
      }
    } break;
    case VITK_EPI_GELU: {
      const bool h = p.out_dtype == VITK_FP16;
      float4 g, d;
      gelu_erf_both(f.x, g.x, d.x); gelu_erf_both(f.y, g.y, d.y); gelu_erf_both(f.z, g.z, d.z); gelu_erf_both(f.w, g.w, d.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col) =
          make_uint2(pack16(d.x, d.y, h), pack16(d.z, d.w, h));
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + orow * p.ldo2 + col) =
          make_uint2(pack16(g.x, g.y, h), pack16(g.z, g.w, h));
    } break;
    case VITK_EPI_DGELU: {
      const bool ah = p.aux_dtype == VITK_FP16, h = p.out_dtype == VITK_FP16;
      const float2 a0 = unpack16(__float_as_uint(e.x), ah), a1 = unpack16(__float_as_uint(e.y), ah);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col) =
          make_uint2(pack16(f.x * a0.x, f.y * a0.y, h), pack16(f.z * a1.x, f.w * a1.y, h));
    } break;
    case VITK_EPI_ATOMIC_ADD:
      red_add_v4(reinterpret_cast<float*>(p.out) + orow * p.ldo + col, f);
      break;
    default:
      break;
  }
}

// 32 rows x 32 16-bit columns (64-byte rows) staged in the TMA SWIZZLE_64B layout: 16-byte chunk c of row r lives at
// r*64 + ((c ^ ((r >> 1) & 3)) << 4)  -- conflict-free for row-per-lane 128-bit accesses.
__device__ __forceinline__ uint32_t swz64(int r, int c) { return uint32_t(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// Stage one packed 32x32 16-bit chunk (h[16] = this lane's row) and hand it to the TMA store engine.  Every warp
// alternates between two 2 KB buffers, so "at most one newer store still reading" means this buffer is free.
__device__ __forceinline__ void stage_and_store(uint8_t* sbuf, const uint32_t (&h)[16], const CUtensorMap* tm, int col0, int row0,
                                                int lane, bool wait_free = true, bool no_store = false) {
  if (wait_free) {
    if (elect_one()) tma_store_wait_read1();  // bulk groups are per thread: the elected lane issues AND waits
    __syncwarp();
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(sbuf + swz64(lane, c)) = make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
  fence_proxy_async();
  __syncwarp();
  if (elect_one() && !no_store) tma_store_2d(tm, sbuf, col0, row0);
}

// 32 accumulator values of one row x 32 saved-derivative values (a[4] = 32 x 16-bit) -> 16 packed 16-bit pairs
template <bool AUX_H16>
__device__ __forceinline__ void mul_pack(uint32_t (&h)[16], const float (&f)[32], const uint4 (&a)[4], bool h16) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float2 a0 = unpack16(a[c].x, AUX_H16), a1 = unpack16(a[c].y, AUX_H16), a2 = unpack16(a[c].z, AUX_H16),
                 a3 = unpack16(a[c].w, AUX_H16);
    h[4 * c + 0] = pack16(f[8 * c + 0] * a0.x, f[8 * c + 1] * a0.y, h16);
    h[4 * c + 1] = pack16(f[8 * c + 2] * a1.x, f[8 * c + 3] * a1.y, h16);
    h[4 * c + 2] = pack16(f[8 * c + 4] * a2.x, f[8 * c + 5] * a2.y, h16);
    h[4 * c + 3] = pack16(f[8 * c + 6] * a3.x, f[8 * c + 7] * a3.y, h16);
  }
}
// 32 accumulators of one row -> *alpha + bias -> 16 packed 16-bit pairs
__device__ __forceinline__ void bias_pack(uint32_t (&h)[16], const uint32_t (&v)[32], const float* __restrict__ sbias, float alpha,
                                          bool h16) {
  float4 b4[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) b4[c] = *reinterpret_cast<const float4*>(sbias + 4 * c);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float f0 = fmaf(__uint_as_float(v[4 * c + 0]), alpha, b4[c].x), f1 = fmaf(__uint_as_float(v[4 * c + 1]), alpha, b4[c].y);
    const float f2 = fmaf(__uint_as_float(v[4 * c + 2]), alpha, b4[c].z), f3 = fmaf(__uint_as_float(v[4 * c + 3]), alpha, b4[c].w);
    h[2 * c + 0] = h16 ? pack_f16(f0, f1) : pack_bf16(f0, f1);
    h[2 * c + 1] = h16 ? pack_f16(f2, f3) : pack_bf16(f2, f3);
  }
}
// hd = gelu'(f), hg = gelu(f), packed 16-bit pairs
template <bool H16>
__device__ __forceinline__ void gelu_pack(uint32_t (&hd)[16], uint32_t (&hg)[16], const float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float2 g, d;
    gelu_erf_both2(make_float2(f[2 * j], f[2 * j + 1]), g, d);
    hd[j] = H16 ? pack_f16(d.x, d.y) : pack_bf16(d.x, d.y);
    hg[j] = H16 ? pack_f16(g.x, g.y) : pack_bf16(g.x, g.y);
  }
}
// same with nn.Dropout applied to the activation: gelu(pre) * m and gelu'(pre) * m (the saved derivative then carries the mask
// into the backward DGELU epilogue for free); blk0 = index of the first 8-element mask block of this row chunk
template <bool H16>
__device__ __forceinline__ void gelu_pack_drop(uint32_t (&hd)[16], uint32_t (&hg)[16], const float (&f)[32], const DropSpec& ds,
                                               unsigned long long dseed, unsigned long long blk0) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 bits = drop_bits8(dseed, ds.site, blk0 + c);
    const uint32_t w[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = 4 * c + q;
      float2 g, d;
      gelu_erf_both2(make_float2(f[2 * j], f[2 * j + 1]), g, d);
      const float2 m = drop_pair(w[q], ds.thresh, ds.inv_keep);
      g = __fmul2_rn(g, m);
      d = __fmul2_rn(d, m);
      hd[j] = H16 ? pack_f16(d.x, d.y) : pack_bf16(d.x, d.y);
      hg[j] = H16 ? pack_f16(g.x, g.y) : pack_bf16(g.x, g.y);
    }
  }
}
// f[0..8) *= mask factors of one 8-element block
__device__ __forceinline__ void drop_apply8(float* f, const DropSpec& ds, unsigned long long dseed, unsigned long long blk) {
  const uint4 bits = drop_bits8(dseed, ds.site, blk);
  const uint32_t w[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 m = drop_pair(w[q], ds.thresh, ds.inv_keep);
    f[2 * q] *= m.x;
    f[2 * q + 1] *= m.y;
  }
}

// ------------------------------------------------------------------ the kernel
template <int BN, int STAGES, bool A_MN, bool B_MN, bool CTA2, bool DROP>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                        const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
  constexpr int NCTA = CTA2 ? 2 : 1;
  constexpr int BN_LOCAL = BN / NCTA;             // rows of the B tile this CTA stages
  constexpr int TILE_M = BLOCK_M * NCTA;          // output rows per tile (per CTA pair in 2-CTA mode)
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_BYTES = BN_LOCAL * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int EPI_BYTES = EPI_WARPS * EPI_STAGE_FLOATS * 4;
  static_assert(!CTA2 || !B_MN || BN_LOCAL % 64 == 0, "MN-major B halves must be whole 64-column boxes");
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  const int unit_id = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // tile-walk position (per pair)
  const int unit_stride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  float* sEpi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  float* sBias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES + MAX_BIAS_SMEM * 4);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator stage ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator stage drained by the epilogue
  uint64_t* aux_bar = tempty_bar + 2;         // [EPI_WARPS] per-warp TMA-load barrier (dGELU pre-activation chunk)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) DBG_G(0);
  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  const int total_tiles = tiles_mn * p.num_splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPI_WARPS * NCTA);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&aux_bar[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CTA2) tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
    else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  // everything above is independent of the previous kernel's output: under programmatic dependent launch it overlaps that
  // kernel's tail.  From here on global memory is touched: wait for the previous grid, and let the next one start launching.
  pdl_trigger();
  pdl_wait();
  if (p.colsum != nullptr) {  // 16 k-rows x 128 B of ones (B operand of the column-sum MMA) in the tail of the bias region
    const uint32_t one2 = (p.idesc & (1u << 7)) ? 0x3f803f80u : 0x3c003c00u;  // bf16 / fp16 1.0 pairs
    for (int i = threadIdx.x; i < 2048 / 4; i += GEMM_THREADS) reinterpret_cast<uint32_t*>(sBias + ONES_OFFSET)[i] = one2;
    fence_proxy_async();
  }
  if (p.tma_epi) {  // bias of every column tile (zeros when there is none), zero-padded to the tile grid
    const int n_up = p.num_n_tiles * BN;
    for (int i = threadIdx.x; i < n_up; i += GEMM_THREADS) sBias[i] = (p.bias != nullptr && i < p.N) ? __ldg(p.bias + i) : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) DBG_G(1);

  // tile -> (m, n, split): n fastest (CTAs running together share the A rows through L2), split slowest.  Every role
  // walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...: the (split, mt, nt) counter advances in mixed radix by the
  // host-computed digits of gridDim.x, so no role pays integer divisions per tile.
  struct TileIter {
    int split, mt, nt;
  };
  auto tile_first = [&]() {
    TileIter t;
    t.split = unit_id / tiles_mn;
    const int r = unit_id - t.split * tiles_mn;
    t.mt = r / p.num_n_tiles;
    t.nt = r - t.mt * p.num_n_tiles;
    return t;
  };
  auto tile_next = [&](TileIter& t) {
    t.nt += p.step_nt;
    t.mt += p.step_mt;
    t.split += p.step_split;
    if (t.nt >= p.num_n_tiles) {
      t.nt -= p.num_n_tiles;
      ++t.mt;
    }
    if (t.mt >= p.num_m_tiles) {
      t.mt -= p.num_m_tiles;
      ++t.split;
    }
  };
  auto decode = [&](const TileIter& t, int& m0, int& n0, int& kb0, int& nk) {
    m0 = t.mt * TILE_M + (int)cta_rank * BLOCK_M;   // this CTA's 128 rows of the tile
    n0 = t.nt * BN;
    kb0 = t.split * p.kblocks_per_split;
    const int kb1 = min(p.num_kblocks, kb0 + p.kblocks_per_split);
    nk = kb1 - kb0;
  };

  if (warp == 0) {
    // ===================== TMA producer (one elected thread) =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      TileIter ti = tile_first();
      for (int tile = unit_id; tile < total_tiles; tile += unit_stride, tile_next(ti)) {
        int m0, n0, kb0, nk;
        decode(ti, m0, n0, kb0, nk);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 1);
          if (kb == 0) DBG_STAMP(0, (tile - unit_id) / unit_stride, 0);
          const int kc = (kb0 + kb) * BLOCK_K;
          uint8_t* a_dst = sA + s * A_BYTES;
          uint8_t* b_dst = sB + s * B_BYTES;
          if (!CTA2) {
            const bool skip_a = KNOB(8) && tile != unit_id, skip_b = KNOB(4) && tile != unit_id;
            mbar_expect_tx(&full_bar[s], (skip_a ? 0 : A_BYTES) + (skip_b ? 0 : B_BYTES));
            if (skip_a) {
            } else if (!A_MN) {
              tma_load_2d(a_dst, &tmA, &full_bar[s], kc, m0);
            } else {
#pragma unroll
              for (int i = 0; i < BLOCK_M / 64; ++i) tma_load_2d(a_dst + i * 8192, &tmA, &full_bar[s], m0 + 64 * i, kc);
            }
            if (skip_b) {
            } else if (!B_MN) {
              tma_load_2d(b_dst, &tmB, &full_bar[s], kc, n0);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i) tma_load_2d(b_dst + i * 8192, &tmB, &full_bar[s], n0 + 64 * i, kc);
            }
          } else {
            // both CTAs' loads are counted on the LEADER's full barrier (it expects the bytes of the whole pair)
            const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
            if (cta_rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
            const int nb0 = n0 + (int)cta_rank * BN_LOCAL;   // this CTA's half of the B tile
            if (!A_MN) {
              tma_load_2d_pair(a_dst, &tmA, bar, kc, m0);
            } else {
#pragma unroll
              for (int i = 0; i < BLOCK_M / 64; ++i) tma_load_2d_pair(a_dst + i * 8192, &tmA, bar, m0 + 64 * i, kc);
            }
            if (!B_MN) {
              tma_load_2d_pair(b_dst, &tmB, bar, kc, nb0);
            } else {
#pragma unroll
              for (int i = 0; i < BN_LOCAL / 64; ++i) tma_load_2d_pair(b_dst + i * 8192, &tmB, bar, nb0 + 64 * i, kc);
            }
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread; in 2-CTA mode the leader CTA issues for the pair) ==========
    if (elect_one() && cta_rank == 0) {
      const uint32_t idesc = p.idesc;
      constexpr uint32_t a_adv = A_MN ? (2048u >> 4) : (32u >> 4);  // desc.lo step per UMMA_K
      constexpr uint32_t b_adv = B_MN ? (2048u >> 4) : (32u >> 4);
      const uint64_t adesc0 = make_smem_desc<A_MN>(smem_u32(sA));
      const uint64_t bdesc0 = make_smem_desc<B_MN>(smem_u32(sB));
      int s = 0;
      uint32_t ph = 0, tcount = 0;
      TileIter ti = tile_first();
      for (int tile = unit_id; tile < total_tiles; tile += unit_stride, ++tcount, tile_next(ti)) {
        int m0, n0, kb0, nk;
        decode(ti, m0, n0, kb0, nk);
        const uint32_t acc = p.nacc == 2 ? (tcount & 1) : 0u;
        const uint32_t acc_ph = p.nacc == 2 ? ((tcount >> 1) & 1) : (tcount & 1);
        DBG_STAMP(1, tcount, 0);
        mbar_wait(&tempty_bar[acc], acc_ph ^ 1, 4);  // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        DBG_STAMP(1, tcount, 1);
        const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
        const bool do_colsum = p.colsum != nullptr && n0 == 0;
        const uint64_t ones_desc = make_smem_desc<true>(smem_u32(sBias + ONES_OFFSET));
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full_bar[s], ph, 2);
          tcgen05_fence_after();
          if (kb == 0) DBG_STAMP(1, tcount, 2);
          const uint64_t adesc = adesc0 + uint64_t(s * (A_BYTES >> 4));
          const uint64_t bdesc = bdesc0 + uint64_t(s * (B_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            if (!KNOB(32)) {
              if (CTA2) umma_f16_pair(d_tmem, adesc + uint64_t(k * a_adv), bdesc + uint64_t(k * b_adv), idesc, (kb > 0 || k > 0) ? 1u : 0u);
              else umma_f16(d_tmem, adesc + uint64_t(k * a_adv), bdesc + uint64_t(k * b_adv), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          if (do_colsum) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              if (CTA2) umma_f16_pair(d_tmem + BN, adesc + uint64_t(k * a_adv), ones_desc, p.idesc_ones, (kb > 0 || k > 0) ? 1u : 0u);
              else umma_f16(d_tmem + BN, adesc + uint64_t(k * a_adv), ones_desc, p.idesc_ones, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (CTA2) umma_commit_pair(&empty_bar[s]);
          else umma_commit(&empty_bar[s]);
          if (kb == nk - 1) {
            if (CTA2) umma_commit_pair(&tfull_bar[acc]);
            else umma_commit(&tfull_bar[acc]);
            DBG_STAMP(1, tcount, 3);
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..) =====================
    const int quad = warp & 3;                // TMEM lane quadrant this warp may read
    const int group = (warp - 2) >> 2;        // which of the EPI_WARPS/4 warps sharing that quadrant
    float* stage = sEpi + (warp - 2) * EPI_STAGE_FLOATS;
    const float alpha = p.alpha_dev != nullptr ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
    // DROP is a template flag (forward, K-major instantiations only) so the dropout code costs the other kernels no registers
    const unsigned long long dseed = (DROP && p.drop.seed != nullptr) ? __ldg(p.drop.seed) : 0ull;
    const int c4 = lane & 3;     // float4 column slot of this lane inside a 16-column half chunk
    const int rsub = lane >> 2;  // row (mod 8) this lane handles when reading the staged half chunk back
    uint32_t tcount = 0, aux_phase = 0, sbuf_idx = 0;
    TileIter ti = tile_first();
    for (int tile = unit_id; tile < total_tiles; tile += unit_stride, ++tcount, tile_next(ti)) {
      int m0, n0, kb0, nk;
      decode(ti, m0, n0, kb0, nk);
      const uint32_t acc = p.nacc == 2 ? (tcount & 1) : 0u;
      const uint32_t acc_ph = p.nacc == 2 ? ((tcount >> 1) & 1) : (tcount & 1);
      const uint32_t t_addr = tmem_base + acc * p.acc_stride + (uint32_t(quad * 32) << 16);
      const int row_base = m0 + quad * 32;
      if (p.tma_epi) {
        // ---- STORE / GELU / DGELU: units of 2 KB (32 rows x 32 16-bit columns, or 32 rows x 16 fp32 columns) ----
        // Math happens in the row-per-lane TMEM layout; the saved derivative / fp32 residual unit arrives by TMA into the
        // staging buffer the result will leave from (issued one unit ahead, the first one before the accumulator is even
        // ready), results leave by TMA store; two buffers per warp alternate so a store overlaps the next unit.
        const bool out32 = p.out_dtype == VITK_FP32, h16 = p.out_dtype == VITK_FP16;
        const int UC = out32 ? 16 : 32;
        const int ustride = UC * (EPI_WARPS / 4);
        const bool has_aux = p.epilogue == VITK_EPI_DGELU || p.residual != nullptr;
        uint8_t* const sbase = reinterpret_cast<uint8_t*>(stage);
        auto issue_aux = [&](int c0) {
          if (elect_one()) {
            tma_store_wait_read1();
            mbar_expect_tx(&aux_bar[warp - 2], 2048);
            tma_load_2d(sbase + (sbuf_idx & 1) * 2048, &tmAux, &aux_bar[warp - 2], n0 + c0, row_base);
          }
        };
        int c0 = group * UC;
        bool more = c0 < BN && n0 + c0 < p.N;
        if (has_aux && more) issue_aux(c0);
        if (warp == 2 && lane == 0) DBG_STAMP(2, tcount, 0);
        mbar_wait(&tfull_bar[acc], acc_ph, 3);
        tcgen05_fence_after();
        if (warp == 2 && lane == 0) DBG_STAMP(2, tcount, 1);
        int ucount = -1;
        while (more) {
          ++ucount;
          const int col0 = n0 + c0;
          uint8_t* sbuf = sbase + (sbuf_idx & 1) * 2048;
          ++sbuf_idx;
          if (KNOB(16)) {
          } else if (out32) {
            uint32_t v[16];
            tmem_ld16(t_addr + uint32_t(c0), v);
            float4 b4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) b4[c] = *reinterpret_cast<const float4*>(sBias + col0 + 4 * c);
            float f[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              f[4 * c + 0] = fmaf(__uint_as_float(v[4 * c + 0]), alpha, b4[c].x);
              f[4 * c + 1] = fmaf(__uint_as_float(v[4 * c + 1]), alpha, b4[c].y);
              f[4 * c + 2] = fmaf(__uint_as_float(v[4 * c + 2]), alpha, b4[c].z);
              f[4 * c + 3] = fmaf(__uint_as_float(v[4 * c + 3]), alpha, b4[c].w);
            }
            if (DROP && p.drop.seed != nullptr) {   // proj_drop / Mlp.drop on the branch output, before it joins the residual stream
              const unsigned long long blk = ((unsigned long long)(row_base + lane) * p.N + col0) >> 3;
              drop_apply8(f, p.drop, dseed, blk);
              drop_apply8(f + 8, p.drop, dseed, blk + 1);
            }
            if (p.row_scale != nullptr) {   // per-sample stochastic-depth factor of this residual branch
              const float rsc = (row_base + lane) < p.M ? __ldg(p.row_scale + row_base + lane) : 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] *= rsc;
            }
            if (has_aux) {
              mbar_wait(&aux_bar[warp - 2], aux_phase, 5);
              aux_phase ^= 1;
              float4 r[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) r[c] = *reinterpret_cast<const float4*>(sbuf + swz64(lane, c));
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                f[4 * c + 0] += r[c].x; f[4 * c + 1] += r[c].y; f[4 * c + 2] += r[c].z; f[4 * c + 3] += r[c].w;
              }
            } else {
              if (elect_one()) tma_store_wait_read1();
            }
            __syncwarp();  // residual rows read by every lane / buffer free: it may now be overwritten with the result
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<float4*>(sbuf + swz64(lane, c)) = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
            fence_proxy_async();
            __syncwarp();
            if (elect_one() && !KNOB(1)) {
              if (p.epilogue == VITK_EPI_ATOMIC_ADD) tma_reduce_add_2d(&tmOut, sbuf, col0, row_base);  // split-K partial tile
              else tma_store_2d(&tmOut, sbuf, col0, row_base);
            }
          } else {
            uint32_t v[32];
            DBG_UNIT(tcount, ucount, 0);
            tmem_ld32(t_addr + uint32_t(c0), v);
            DBG_UNIT(tcount, ucount, 1);
            float4 b4[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) b4[c] = *reinterpret_cast<const float4*>(sBias + col0 + 4 * c);
            float f[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              f[4 * c + 0] = fmaf(__uint_as_float(v[4 * c + 0]), alpha, b4[c].x);
              f[4 * c + 1] = fmaf(__uint_as_float(v[4 * c + 1]), alpha, b4[c].y);
              f[4 * c + 2] = fmaf(__uint_as_float(v[4 * c + 2]), alpha, b4[c].z);
              f[4 * c + 3] = fmaf(__uint_as_float(v[4 * c + 3]), alpha, b4[c].w);
            }
            uint32_t h[16];
            if (KNOB(2)) {
            } else if (p.epilogue == VITK_EPI_DGELU) {
              mbar_wait(&aux_bar[warp - 2], aux_phase, 5);
              aux_phase ^= 1;
              uint4 a[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) a[c] = *reinterpret_cast<const uint4*>(sbuf + swz64(lane, c));
              __syncwarp();  // every lane has read its derivative row before the buffer is reused for the output
              if (p.aux_dtype == VITK_FP16) mul_pack<true>(h, f, a, h16);
              else mul_pack<false>(h, f, a, h16);
              stage_and_store(sbuf, h, &tmOut, col0, row_base, lane, false, KNOB(1));
            } else if (p.epilogue == VITK_EPI_GELU) {
              // out = gelu'(pre) (all the backward needs), out2 = gelu(pre): one erf evaluation serves both
              uint32_t h2[16];
              if (DROP && p.drop.seed != nullptr) {
                const unsigned long long blk = ((unsigned long long)(row_base + lane) * p.N + col0) >> 3;
                if (h16) gelu_pack_drop<true>(h, h2, f, p.drop, dseed, blk);
                else gelu_pack_drop<false>(h, h2, f, p.drop, dseed, blk);
              } else if (h16) gelu_pack<true>(h, h2, f);
              else gelu_pack<false>(h, h2, f);
              stage_and_store(sbuf, h, &tmOut, col0, row_base, lane, true, KNOB(1));
              uint8_t* sbuf2 = sbase + (sbuf_idx & 1) * 2048;
              ++sbuf_idx;
              stage_and_store(sbuf2, h2, &tmOut2, col0, row_base, lane, true, KNOB(1));
            } else {
              if (h16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = pack_f16(f[2 * j], f[2 * j + 1]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
              }
              DBG_UNIT(tcount, ucount, 2);
              if (elect_one()) tma_store_wait_read1();
              __syncwarp();
              DBG_UNIT(tcount, ucount, 3);
#pragma unroll
              for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(sbuf + swz64(lane, c)) = make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
              fence_proxy_async();
              __syncwarp();
              DBG_UNIT(tcount, ucount, 4);
              if (elect_one() && !KNOB(1)) tma_store_2d(&tmOut, sbuf, col0, row_base);
              DBG_UNIT(tcount, ucount, 5);
            }
          }
          c0 += ustride;
          more = c0 < BN && n0 + c0 < p.N;
          if (has_aux && more) issue_aux(c0);
        }
        if (p.colsum != nullptr && n0 == 0 && group == 0) {  // row sums of A over this K range: one fp32 RED per row
          const float v = __uint_as_float(tmem_ld1(t_addr + uint32_t(BN)));
          if (row_base + lane < p.M) atomicAdd(p.colsum + row_base + lane, v * alpha);
        }
        tcgen05_fence_before();
        if (lane == 0) {
          if (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
          else mbar_arrive(&tempty_bar[acc]);
        }
        if (warp == 2 && lane == 0) DBG_STAMP(2, tcount, 2);
        continue;
      }
      if (warp == 2 && lane == 0) DBG_STAMP(2, tcount, 0);
      mbar_wait(&tfull_bar[acc], acc_ph, 3);
      tcgen05_fence_after();
      if (warp == 2 && lane == 0) DBG_STAMP(2, tcount, 1);
#pragma unroll 1
      for (int c0 = group * 32; c0 < BN; c0 += 32 * (EPI_WARPS / 4)) {
        const int col0 = n0 + c0;
        if (col0 >= p.N) break;  // warp-uniform
        if (KNOB(16)) continue;
        uint32_t v[32];
        tmem_ld32(t_addr + uint32_t(c0), v);
        // transpose through smem in two 16-column halves: lane = row on the way in,
        // lane = (row mod 8, float4 column) on the way out -> every global access covers whole 32/64-byte row segments
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          float* wr = stage + lane * EPI_PITCH;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(wr + j) =
                make_float4(__uint_as_float(v[hlf * 16 + j]), __uint_as_float(v[hlf * 16 + j + 1]),
                            __uint_as_float(v[hlf * 16 + j + 2]), __uint_as_float(v[hlf * 16 + j + 3]));
          __syncwarp();
          const int col = col0 + hlf * 16 + c4 * 4;
          if (col < p.N) {
            const float4 bias4 = p.bias != nullptr ? ldg_f4(p.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 ex[4];
            long long orow[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // phase 1: every global load of this half chunk in flight
              const int row = row_base + i * 8 + rsub;
              orow[i] = row < p.M ? epilogue_out_row(p, row) : 0;
              ex[i] = row < p.M ? epilogue_load(p, row, orow[i], col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // phase 2: arithmetic + stores
              const int r = i * 8 + rsub;
              if (row_base + r < p.M) {
                const float4 f = *reinterpret_cast<const float4*>(stage + r * EPI_PITCH + c4 * 4);
                epilogue_apply<DROP>(p, f, ex[i], orow[i], col, alpha, bias4, dseed);
              }
            }
          }
          __syncwarp();
        }
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld inside tmem_ld32): release the accumulator stage
      tcgen05_fence_before();
      if (lane == 0) {
        if (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (warp == 2 && lane == 0) DBG_STAMP(2, tcount, 2);
    }
    if (warp == 2 && lane == 0) DBG_G(2);
    if (elect_one()) tma_store_wait_all();  // outstanding TMA stores must complete before the CTA (and its smem) goes away
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer may still be reading this CTA's B half / arriving on its barriers
  if (warp == 1) {
    if (CTA2) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (threadIdx.x == 0) DBG_G(3);
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

// 2-D 16-bit tensor map over a row-major [outer, inner] matrix, 128B swizzle, zero OOB fill.
int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner,
                 uint32_t box_outer, bool fp16, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, bool fp32 = false) {
  auto fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VITK_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * (fp32 ? 4 : 2)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (base=%p inner=%llu outer=%llu pitch=%llu box=%ux%u)", (int)r, base,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner, box_outer);
    return VITK_ERR_CUDA;
  }
  return VITK_OK;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, bool CTA2, bool DROP = false>
int launch_gemm(const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st) {
  if constexpr (!DROP && !A_MN && !B_MN) {   // forward linears: the dropout-capable instantiation when a mask is requested
    if (p.drop.seed != nullptr) return launch_gemm<BN, STAGES, A_MN, B_MN, CTA2, true>(tm, p, grid, st);
  }
  constexpr int SMEM = STAGES * (BLOCK_M * BLOCK_K * 2 + (BN / (CTA2 ? 2 : 1)) * BLOCK_K * 2) + EPI_WARPS * EPI_STAGE_FLOATS * 4 +
                       MAX_BIAS_SMEM * 4 + (2 * STAGES + 4 + EPI_WARPS) * 8 + 16 + 1024;
  static_assert(SMEM <= 232448, "shared memory budget exceeded");
  static bool configured = false;
  auto kfn = gemm_tcgen05_kernel<BN, STAGES, A_MN, B_MN, CTA2, DROP>;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  VITK_CUDA(launch_pdl_cluster(kfn, dim3(grid), dim3(GEMM_THREADS), SMEM, st, CTA2 ? 2 : 1, tm[0], tm[1], tm[2], tm[3], tm[4], p));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

template <bool A_MN, bool B_MN>
int dispatch_gemm(int bn, bool cta2, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st) {
  if (cta2) {  // CTA pairs: 256 x BN tiles, half of B per CTA -> deeper rings in the same shared memory
    switch (bn) {
      case 128: return launch_gemm<128, 6, A_MN, B_MN, true>(tm, p, grid, st);
      case 256: return launch_gemm<256, 5, A_MN, B_MN, true>(tm, p, grid, st);
      default: break;
    }
    set_error("2-CTA mode supports BLOCK_N 128 / 256 (got %d)", bn);
    return VITK_ERR_UNSUPPORTED;
  }
  switch (bn) {
    case 64:  return launch_gemm<64, 6, A_MN, B_MN, false>(tm, p, grid, st);
    case 128: return launch_gemm<128, 5, A_MN, B_MN, false>(tm, p, grid, st);
    case 192: return launch_gemm<192, 4, A_MN, B_MN, false>(tm, p, grid, st);
    case 256: return launch_gemm<256, 3, A_MN, B_MN, false>(tm, p, grid, st);
    default:
      set_error("unsupported BLOCK_N %d", bn);
      return VITK_ERR_UNSUPPORTED;
  }
}

bool pair_mode_enabled() {
  static const bool on = [] {
    const char* e = getenv("VITK_GEMM_2CTA");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

// Pick the N tile.  A persistent grid of `units` CTAs walks m_tiles x n_tiles tiles in rounds, and the time of a tile is
// ~(50 + BN) column-units (fitted on B200 at K = 192: 1.44 / 2.12 / 3.07 us per round for BN = 64 / 128 / 192 -- a fixed
// hand-off cost plus an epilogue proportional to the width), so the candidate with the smallest rounds x (50 + BN) wins:
// e.g. N = 768 at M = 50688 takes 9 rounds of BN = 256 (1188 tiles on 148 SMs: the 9th round runs 4 tiles) but 11 fuller
// rounds of BN = 192 -- measured 56.4 -> 52.3 us (fc1 + GELU) and 52.3 -> 48.2 us (dGELU dgrad).  Ties go to the wider tile.
// Without M (split-K accumulate GEMMs pick their split afterwards) the widest tile with the least padding is taken.
int pick_bn(int N, int M, bool balance) {
  static const int forced = [] {            // VITK_GEMM_BN=64|128|192|256: experiments only
    const char* e = getenv("VITK_GEMM_BN");
    return e != nullptr ? atoi(e) : 0;
  }();
  if (forced == 64 || forced == 128 || forced == 192 || forced == 256) return forced;
  const int cands[4] = {256, 192, 128, 64};
  int best = 64;
  long best_cost = -1;
  const long units = num_sms();
  const long m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  for (int c : cands) {
    const long n_tiles = (N + c - 1) / c;
    long cost = n_tiles * c;  // padded width
    if (balance) {
      const long rounds = (m_tiles * n_tiles + units - 1) / units;
      cost = rounds * (50 + c);
    }
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

#ifdef VITK_GEMM_KNOBS
extern "C" int vitk_debug_read(long long* dst) {
  return (int)cudaMemcpyFromSymbol(dst, g_vitk_dbg, sizeof(g_vitk_dbg));
}
extern "C" int vitk_debug_read3(long long* dst) {
  return (int)cudaMemcpyFromSymbol(dst, g_vitk_dbg3, sizeof(g_vitk_dbg3));
}
extern "C" int vitk_debug_read2(long long* dst) {
  return (int)cudaMemcpyFromSymbol(dst, g_vitk_dbg2, sizeof(g_vitk_dbg2));
}
#endif

extern "C" int vitk_gemm(const vitk_gemm_args* a, void* stream) {
  VITK_CHECK_ARG(a != nullptr, "vitk_gemm: null args");
  VITK_CHECK_ARG(a->A && a->B && a->out, "vitk_gemm: null operand");
  VITK_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "vitk_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  VITK_CHECK_ARG(a->N % 8 == 0, "vitk_gemm: N=%d must be a multiple of 8", a->N);
  VITK_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0, "vitk_gemm: lda/ldb must be multiples of 8 (16-byte TMA pitch)");
  VITK_CHECK_ARG((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0,
                 "vitk_gemm: operands must be 16-byte aligned");
  VITK_CHECK_ARG(a->split_k >= 0, "vitk_gemm: split_k must be >= 1 (or 0: chosen by the library)");
  VITK_CHECK_ARG(a->split_k <= 1 || a->epilogue == VITK_EPI_ATOMIC_ADD, "vitk_gemm: split_k > 1 needs VITK_EPI_ATOMIC_ADD");
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
  VITK_CHECK_ARG(a->ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
                 "vitk_gemm: out rows must stay 8/16-byte aligned (ldo %% 4 == 0, 16-byte aligned base)");
  if (a->residual)
    VITK_CHECK_ARG(a->ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0, "vitk_gemm: residual alignment");
  if (a->bias) VITK_CHECK_ARG((reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "vitk_gemm: bias must be 16-byte aligned");
  if (a->colsum_out)
    VITK_CHECK_ARG(a->epilogue == VITK_EPI_ATOMIC_ADD && out_fp32 && a->bias == nullptr && a->residual == nullptr && a->ldo % 4 == 0,
                   "vitk_gemm: colsum_out needs the split-K accumulate epilogue (fp32 out, no bias / residual)");

  const int bn = pick_bn(a->N, a->M, a->epilogue != VITK_EPI_ATOMIC_ADD && a->M > 4 * BLOCK_M);
  const int num_kblocks = (a->K + BLOCK_K - 1) / BLOCK_K;
  int split_k = a->split_k;
  bool pair_ok = pair_mode_enabled() && (bn == 256 || bn == 128) && a->M > BLOCK_M && (num_sms() % 2 == 0);
  if (split_k == 0) {
    // split_k = 0: the library picks the K split of an accumulate GEMM -- about two work units per CTA (or CTA pair), at least
    // 8 k-blocks each, so that every SM is busy and the epilogue of one unit overlaps the MMAs of the next
    if (a->epilogue != VITK_EPI_ATOMIC_ADD) {
      split_k = 1;
    } else {
      const bool pair = pair_ok && num_kblocks >= 32;
      const int tile_rows = pair ? 2 * BLOCK_M : BLOCK_M;
      const long tiles = (long)((a->M + tile_rows - 1) / tile_rows) * ((a->N + bn - 1) / bn);
      const long units = pair ? num_sms() / 2 : num_sms();
      // CTA pairs (long K, ViT-B): about two work units per pair, so that the epilogue of one overlaps the MMAs of the next.
      // Single CTAs (DeiT-tiny's 192-wide gradients): ONE round -- the units are L2-bandwidth-bound, a second round only adds
      // its 1.9-round tail and twice the reduce-add traffic (measured, tools/wgrad_scan.py: qkv 31.7 -> 29.7 us, fc1 33.8 ->
      // 31.7, fc2 35.8 -> 33.8, proj 21.4 -> 19.5)
      long want = pair ? (2 * units) / tiles : units / tiles;   // floor: never a partial extra round of work units
      const long cap = num_kblocks / 8 > 0 ? num_kblocks / 8 : 1;
      split_k = (int)(want < 1 ? 1 : want > cap ? cap : want);
    }
  }
  const int kpb = (num_kblocks + split_k - 1) / split_k;
  const int splits = (num_kblocks + kpb - 1) / kpb;
  // CTA pairs (tcgen05 cta_group::2): each SM reads its A tile and only half of the B tile from shared memory per MMA
  // -- 64 instead of 96 B/clk of operand reads, and a third less TMA write traffic, through the 128 B/clk shared memory.  It
  // pays when the tile is MMA-bound (long K); short-K tiles are epilogue-bound and only suffer the cross-CTA handshakes.
  const bool long_k = kpb >= 24 || (kpb >= 8 && (!out_fp32 || a->epilogue == VITK_EPI_ATOMIC_ADD));
  const bool cta2 = pair_ok && long_k;
  const int tile_m = cta2 ? 2 * BLOCK_M : BLOCK_M;
  GemmParams p{};
#ifdef VITK_GEMM_KNOBS
  {
    const char* kn = getenv("VITK_GEMM_KNOBS");
    p.knobs = kn ? atoi(kn) : 0;
  }
#endif
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_kblocks = num_kblocks;
  p.kblocks_per_split = kpb;
  p.num_m_tiles = (a->M + tile_m - 1) / tile_m;
  p.num_n_tiles = (a->N + bn - 1) / bn;
  p.num_splits = splits;
  p.epilogue = a->epilogue; p.out_dtype = a->out_dtype; p.aux_dtype = a->aux_dtype;
  p.idesc = make_idesc(bn, a->a_mn_major != 0, a->b_mn_major != 0, a->a_dtype == VITK_FP16, a->b_dtype == VITK_FP16, tile_m);
  p.alpha = a->alpha; p.alpha_dev = a->alpha_dev;
  p.bias = a->bias; p.residual = a->residual; p.ldr = a->ldr;
  p.out = a->out; p.ldo = a->ldo; p.out2 = a->out2; p.ldo2 = a->ldo2;
  p.aux = a->aux; p.ldaux = a->ldaux;
  p.colsum = a->colsum_out;
  p.row_scale = a->row_scale;
  p.drop = make_drop_spec(a->drop_seed, a->drop_p, a->drop_site);
  p.acc_stride = bn + (a->colsum_out != nullptr ? 16 : 0);
  p.nacc = 2 * p.acc_stride <= 512 ? 2 : 1;
  {
    const int need = p.nacc * p.acc_stride;
    p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
  }
  p.idesc_ones = make_idesc(16, a->a_mn_major != 0, true, a->a_dtype == VITK_FP16, a->b_dtype == VITK_FP16, tile_m);
  p.rows_per_img = a->rows_per_img; p.tokens_per_img = a->tokens_per_img; p.prefix = a->prefix; p.pos = a->pos;

  CUtensorMap tm[5];  // A, B, out, out2, aux
  int rc;
  const bool ah = a->a_dtype == VITK_FP16, bh = a->b_dtype == VITK_FP16;
  if (!a->a_mn_major) rc = make_tmap_2d(&tm[0], a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BLOCK_K, BLOCK_M, ah);
  else                rc = make_tmap_2d(&tm[0], a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BLOCK_K, ah);
  if (rc != VITK_OK) return rc;
  if (!a->b_mn_major) rc = make_tmap_2d(&tm[1], a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BLOCK_K, (uint32_t)(cta2 ? bn / 2 : bn), bh);
  else                rc = make_tmap_2d(&tm[1], a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BLOCK_K, bh);
  if (rc != VITK_OK) return rc;
  // STORE / GELU / DGELU epilogues run on 2 KB TMA units (32x32 16-bit or 32x16 fp32 boxes, 64-byte rows, SWIZZLE_64B);
  // the only combination left on the per-thread path is a 16-bit output with an fp32 residual (its residual unit would be 4 KB)
  const bool store_like = a->epilogue == VITK_EPI_STORE || a->epilogue == VITK_EPI_GELU || a->epilogue == VITK_EPI_DGELU ||
                          (a->epilogue == VITK_EPI_ATOMIC_ADD && a->bias == nullptr && a->residual == nullptr);
  const int oalign = out_fp32 ? 4 : 8;
  p.tma_epi = (store_like && (a->residual == nullptr || out_fp32) && a->ldo % oalign == 0 &&
               (a->residual == nullptr || (a->ldr % 4 == 0)) &&
               (a->epilogue != VITK_EPI_GELU || (a->ldo2 % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out2) & 15) == 0)) &&
               (a->epilogue != VITK_EPI_DGELU || (a->ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0)))
                  ? 1 : 0;
  if (p.num_n_tiles * bn > MAX_BIAS_SMEM) p.tma_epi = 0;  // the TMA epilogue keeps the whole bias vector in shared memory
  if (a->row_scale != nullptr)
    VITK_CHECK_ARG(p.tma_epi && out_fp32 && a->epilogue == VITK_EPI_STORE, "vitk_gemm: row_scale needs the fp32 STORE epilogue");
  if (p.drop.seed != nullptr) {
    VITK_CHECK_ARG(a->drop_p < 1.f && a->N % 8 == 0 && !a->a_mn_major && !a->b_mn_major,
                   "vitk_gemm: dropout needs p < 1, N %% 8 == 0 and K-major operands (a forward linear)");
    VITK_CHECK_ARG(a->epilogue == VITK_EPI_TOKENS || (p.tma_epi && (a->epilogue == VITK_EPI_GELU ||
                                                                   (a->epilogue == VITK_EPI_STORE && out_fp32))),
                   "vitk_gemm: dropout is fused into the TOKENS, GELU and fp32 STORE epilogues only");
  }
  if (a->colsum_out != nullptr)
    VITK_CHECK_ARG(p.tma_epi && p.num_n_tiles * bn <= ONES_OFFSET, "vitk_gemm: colsum_out supports N <= %d", ONES_OFFSET);
  tm[2] = tm[0]; tm[3] = tm[0]; tm[4] = tm[0];
  if (p.tma_epi) {
    const bool oh = a->out_dtype == VITK_FP16;
    rc = make_tmap_2d(&tm[2], a->out, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldo, out_fp32 ? 16 : 32, 32, oh,
                      CU_TENSOR_MAP_SWIZZLE_64B, out_fp32);
    if (rc != VITK_OK) return rc;
    if (a->epilogue == VITK_EPI_GELU) {
      rc = make_tmap_2d(&tm[3], a->out2, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldo2, 32, 32, oh, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != VITK_OK) return rc;
    }
    if (a->epilogue == VITK_EPI_DGELU) {
      rc = make_tmap_2d(&tm[4], a->aux, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldaux, 32, 32, a->aux_dtype == VITK_FP16,
                        CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != VITK_OK) return rc;
    } else if (a->residual != nullptr) {
      rc = make_tmap_2d(&tm[4], a->residual, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldr, 16, 32, false,
                        CU_TENSOR_MAP_SWIZZLE_64B, true);
      if (rc != VITK_OK) return rc;
    }
  }

  const long long total_tiles = (long long)p.num_m_tiles * p.num_n_tiles * splits;
  const int max_units = cta2 ? num_sms() / 2 : num_sms();                 // CTA pairs, or CTAs
  const int units = (int)(total_tiles < max_units ? total_tiles : max_units);
  const int grid = cta2 ? 2 * units : units;
  {
    const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
    p.step_split = units / tiles_mn;
    const int r = units - p.step_split * tiles_mn;
    p.step_mt = r / p.num_n_tiles;
    p.step_nt = r - p.step_mt * p.num_n_tiles;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!a->a_mn_major && !a->b_mn_major) return dispatch_gemm<false, false>(bn, cta2, tm, p, grid, st);
  if (!a->a_mn_major && a->b_mn_major) return dispatch_gemm<false, true>(bn, cta2, tm, p, grid, st);
  if (a->a_mn_major && a->b_mn_major) return dispatch_gemm<true, true>(bn, cta2, tm, p, grid, st);
  return dispatch_gemm<true, false>(bn, cta2, tm, p, grid, st);
}
