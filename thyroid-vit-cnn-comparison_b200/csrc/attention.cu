// attention.cu -- C-ABI entry points of the fused softmax attention (head_dim 64: every ViT/DeiT variant of the reference,
// tiny 192/3, small 384/6, base 768/12) and its two CUDA-core helpers.
//
// Replaces Attention.forward's q@k^T*scale -> softmax -> attn@v (vision_transformer_base.py:182-191) and its autograd
// backward.  The [B,H,N,N] score tensor is never written in training: the forward saves only the log-sum-exp, the backward
// recomputes P from q, k and lse.  Layouts are the reference's own: qkv [B,N,3,H,64] is what nn.Linear(D,3D) emits (:178), the
// output [B,N,H,64] is (attn@v).transpose(1,2).reshape(B,N,C) (:191) -- no permute copies.
//
// Element type: fp16 or bf16 for ALL of qkv / out / dout / dqkv of one call; the engine uses fp16 (gradients are loss-scaled),
// bf16 stays available.  Softmax statistics and accumulators are fp32.
//
// Every matrix product runs on the tcgen05 / TMEM kernels of attention_tc.cu, for every sequence length and with or without
// attention-probability dropout (the round-1 mma.sync kernels that used to serve long sequences and dropout are gone).  What
// lives here: argument validation and delta = rowsum(dO o O) for the streaming long-sequence backward.
#include "vitk_common.cuh"

namespace vitk {
namespace {

constexpr int DH = 64;

typedef __nv_bfloat16 bf16;

template <bool H16>
__device__ __forceinline__ float2 upk(uint32_t u) {
  if (H16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
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

}  // namespace

// tcgen05 / TMEM kernels (attention_tc.cu)
constexpr int TC_MAX_TOKENS_BWD = 240;  // the single-CTA pipelined backward (shared-memory budget); streaming kernels beyond
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
  return attention_probs_tc(qkv, lse, probs, batch_stride, B, N, H, scale, H16, st);   // tensor-core S + the forward's lse
}

template <bool H16>
static int attention_fwd_impl(const void* qkv, void* out, float* lse, float* probs, int B, int N, int H, float scale, int q_rows,
                              cudaStream_t st) {
  const int rc = attention_fwd_tc(qkv, out, lse, B, N, H, scale, probs != nullptr ? 0 : q_rows, H16, nullptr, st);  // maps need every lse
  if (rc != VITK_OK) return rc;
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

template <bool H16, bool DROP>
static int attention_bwd_impl(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                              int B, int N, int H, float scale, DropSpec drop, cudaStream_t st, int q_rows = 0) {
  if (N <= TC_MAX_TOKENS_BWD) return attention_bwd_tc(qkv, out, dout, lse, delta, dqkv, B, N, H, scale, q_rows, H16, DROP ? &drop : nullptr, st);  // delta fused
  const long long rows = (long long)B * N * H;
  attn_delta_kernel<H16><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(reinterpret_cast<const bf16*>(out),
                                                                     reinterpret_cast<const bf16*>(dout), delta, B, N, H);
  VITK_LAUNCH_CHECK();
  return attention_bwd_tc_long(qkv, dout, lse, delta, dqkv, B, N, H, scale, q_rows, H16, DROP ? &drop : nullptr, st);
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
// ViT / DeiT configuration of the reference sets the rate to 0.  DROP instantiations of the same tcgen05 kernels.
extern "C" int vitk_attention_dropout_fwd(const void* qkv, void* out, int32_t dtype, float* lse, int32_t B, int32_t N, int32_t H,
                                          float scale, const vitk_dropout* attn_drop, void* stream) {
  VITK_CHECK_ARG(qkv && out && lse && attn_drop && attn_drop->seed, "vitk_attention_dropout_fwd: null pointer");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_attention_dropout_fwd: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "vitk_attention_dropout_fwd: bad shape B=%d N=%d H=%d", B, N, H);
  VITK_CHECK_ARG(attn_drop->p > 0.f && attn_drop->p < 1.f, "vitk_attention_dropout_fwd: need 0 < p < 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const DropSpec ds = make_drop_spec(attn_drop->seed, attn_drop->p, attn_drop->site);
  return attention_fwd_tc(qkv, out, lse, B, N, H, scale, 0, dtype == VITK_FP16, &ds, st);
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
