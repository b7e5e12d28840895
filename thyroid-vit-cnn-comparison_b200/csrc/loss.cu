// loss.cu -- fused classification / distillation loss with its gradient, one launch.
//
// Restates, for logits [B,C]:
//   plain ViT/DeiT step   lightning_modules.py:455-465   CE(cls,y)  or  0.5*CE(cls,y)+0.5*CE(dist,y)
//   distillation step     lightning_modules.py:959-974   (1-a)*CE(cls,y) + a*KL(log_softmax(d/T) || softmax(t/T))*T^2
//                                                        ('batchmean'), or hard: CE(d, argmax t)
//   DistillationLoss      deit_models.py:461-480          same formula
//   nn.CrossEntropyLoss(label_smoothing=eps)              lightning_modules.py:345-350
// plus the two counters the step logs (train_acc, teacher_agreement; :467,:976-979).
// The logits are a few KB: the kernel is latency-bound by design (single CTA, no HBM roofline).
#include "vitk_common.cuh"

namespace vitk {
namespace {

constexpr int LOSS_THREADS = 256;
constexpr int MAX_C = 1024;

// log-softmax statistics of row r scaled by inv_t: returns max and log-sum-exp
__device__ __forceinline__ void row_lse(const float* r, int C, float inv_t, float& mx, float& lse) {
  mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, r[c] * inv_t);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(r[c] * inv_t - mx);
  lse = mx + logf(s);
}
__device__ __forceinline__ int row_argmax(const float* r, int C) {
  int best = 0;
  float bv = r[0];
  for (int c = 1; c < C; ++c)
    if (r[c] > bv) {
      bv = r[c];
      best = c;
    }
  return best;
}
// CE with label smoothing on one row; writes grad (scaled by gscale) if g != nullptr
__device__ __forceinline__ float ce_row(const float* r, int C, int target, float eps, float gscale, float* g) {
  float mx, lse;
  row_lse(r, C, 1.f, mx, lse);
  float smooth = 0.f;
  for (int c = 0; c < C; ++c) smooth += lse - r[c];
  const float nll = lse - r[target];
  if (g != nullptr) {
    for (int c = 0; c < C; ++c) {
      const float p = expf(r[c] - lse);
      const float tgt = (c == target ? 1.f - eps : 0.f) + eps / float(C);
      g[c] = (p - tgt) * gscale;
    }
  }
  return (1.f - eps) * nll + eps * smooth / float(C);
}

__global__ void __launch_bounds__(LOSS_THREADS)
    loss_kernel(const float* __restrict__ cls, const float* __restrict__ dist, const float* __restrict__ teacher,
                const long long* __restrict__ labels, float* __restrict__ out, float* __restrict__ dcls,
                float* __restrict__ ddist, int B, int C, int mode, float w_cls, float w_dist, float T, float eps,
                float grad_div) {
  __shared__ float red[5][LOSS_THREADS / 32];
  float acc_cls = 0.f, acc_dist = 0.f, n_correct = 0.f, n_agree = 0.f, n_bad = 0.f;
  const float gscale_cls = w_cls / (float(B) * grad_div);
  const float gscale_dist = w_dist / (float(B) * grad_div);
  for (int b = threadIdx.x; b < B; b += LOSS_THREADS) {
    const float* rc = cls + (long long)b * C;
    // a label outside [0, C) must not index the row (torch's CE raises a device assert): it is counted in out[6], the
    // loss becomes NaN and so do this row's gradients, which makes the optimizer skip the step (non-finite norm)
    const long long yl = labels[b];
    const bool bad = yl < 0 || yl >= (long long)C;
    const int y = bad ? 0 : (int)yl;
    n_bad += bad ? 1.f : 0.f;
    acc_cls += ce_row(rc, C, y, eps, bad ? __int_as_float(0x7fc00000) : gscale_cls, dcls + (long long)b * C);
    const int pred = row_argmax(rc, C);
    n_correct += (pred == y) ? 1.f : 0.f;
    if (teacher != nullptr) n_agree += (pred == row_argmax(teacher + (long long)b * C, C)) ? 1.f : 0.f;
    if (dist != nullptr) {
      const float* rd = dist + (long long)b * C;
      float* gd = ddist + (long long)b * C;
      if (mode == 0) {
        acc_dist += ce_row(rd, C, y, eps, gscale_dist, gd);
      } else if (mode == 2) {
        acc_dist += ce_row(rd, C, row_argmax(teacher + (long long)b * C, C), eps, gscale_dist, gd);
      } else {
        // soft KL: sum_c p_t (log p_t - log q) * T^2, q = softmax(d/T), p_t = softmax(t/T)
        const float* rt = teacher + (long long)b * C;
        const float inv_t = 1.f / T;
        float mq, lq, mt, lt;
        row_lse(rd, C, inv_t, mq, lq);
        row_lse(rt, C, inv_t, mt, lt);
        float kl = 0.f;
        for (int c = 0; c < C; ++c) {
          const float logp = rt[c] * inv_t - lt;
          const float logq = rd[c] * inv_t - lq;
          const float p = expf(logp);
          if (p > 0.f) kl += p * (logp - logq);
          // d/dd_c [T^2 * KL] = T * (q_c - p_c)
          gd[c] = (expf(logq) - p) * T * gscale_dist;
        }
        acc_dist += kl * T * T;
      }
    }
  }
  float vals[5] = {acc_cls, acc_dist, n_correct, n_agree, n_bad};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float v = warp_sum(vals[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 5; ++k)
      for (int w = 0; w < LOSS_THREADS / 32; ++w) tot[k] += red[k][w];
    const float cls_loss = tot[0] / float(B);
    const float dist_loss = tot[1] / float(B);
    out[0] = tot[4] > 0.f ? __int_as_float(0x7fc00000) : w_cls * cls_loss + (dist != nullptr ? w_dist * dist_loss : 0.f);
    out[1] = cls_loss;
    out[2] = dist_loss;
    out[3] = tot[2];
    out[4] = tot[3];
    out[5] = float(B);
    out[6] = tot[4];   // number of labels outside [0, C)
    out[7] = 0.f;
  }
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_loss_fwd_bwd(const float* cls_logits, const float* dist_logits, const float* teacher_logits,
                                 const int64_t* labels, float* out_scalars, float* dcls, float* ddist, int32_t B, int32_t C,
                                 int32_t mode, float w_cls, float w_dist, float T, float label_smoothing, float grad_div,
                                 void* stream) {
  VITK_CHECK_ARG(cls_logits && labels && out_scalars && dcls, "vitk_loss_fwd_bwd: null pointer");
  VITK_CHECK_ARG(B > 0 && C > 1 && C <= MAX_C, "vitk_loss_fwd_bwd: bad shape B=%d C=%d", B, C);
  VITK_CHECK_ARG(mode >= 0 && mode <= 2, "vitk_loss_fwd_bwd: bad mode %d", mode);
  VITK_CHECK_ARG(dist_logits == nullptr || ddist != nullptr, "vitk_loss_fwd_bwd: ddist required with dist_logits");
  VITK_CHECK_ARG(mode == 0 || (dist_logits && teacher_logits), "vitk_loss_fwd_bwd: distillation modes need dist and teacher logits");
  VITK_CHECK_ARG(mode != 1 || T > 0.f, "vitk_loss_fwd_bwd: temperature must be > 0");
  VITK_CHECK_ARG(grad_div > 0.f, "vitk_loss_fwd_bwd: grad_div must be > 0");
  loss_kernel<<<1, LOSS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      cls_logits, dist_logits, teacher_logits, reinterpret_cast<const long long*>(labels), out_scalars, dcls, ddist, B, C, mode,
      w_cls, w_dist, T, label_smoothing, grad_div);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
