// metrics.cu -- on-device classification metrics for the validation / test steps.
//
// Replaces the torchmetrics objects of lightning_modules.py:358-374 and their per-step updates (:496-516 validation,
// :542-560 test, :902-920 distillation module), which in the reference run on the host side of every step:
//   Accuracy / F1Score / Specificity / Recall / Precision / StatScores are all functions of the confusion counters,
//   AUROC(task='binary', thresholds=None) is the area under the exact ROC curve, i.e. the Mann-Whitney statistic
//   (pairs with score_pos > score_neg, ties counted one half) / (P * N).
// Everything stays on the device in integer counters: one tiny latency-bound update launch per step (confusion matrix +
// append of the positive-class probability), one pairwise counting launch per epoch for the AUROC.  Integer counts make
// the result independent of summation order (bit-exact against the oracle).
#include "vitk_common.cuh"

namespace vitk {
namespace {

constexpr int MET_THREADS = 256;

// confusion[label * C + argmax] += 1 (slot C*C counts labels outside [0, C)); for C == 2 the softmax probability of
// class 1 (F.softmax(logits, 1)[:, 1], lightning_modules.py:493-495) and the label are appended at *count.
__global__ void __launch_bounds__(MET_THREADS)
    metrics_update_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int C,
                          unsigned long long* __restrict__ confusion, float* __restrict__ scores,
                          unsigned char* __restrict__ score_labels, unsigned long long* __restrict__ count, long long capacity) {
  const unsigned long long base = count != nullptr ? *count : 0ull;
  __syncthreads();   // every thread has read the old count before thread 0 advances it
  for (int b = threadIdx.x; b < B; b += MET_THREADS) {
    const float* r = logits + (long long)b * C;
    int pred = 0;
    float best = r[0];
    for (int c = 1; c < C; ++c)
      if (r[c] > best) {   // first maximum wins, as torch.argmax
        best = r[c];
        pred = c;
      }
    const long long y = labels[b];
    if (y < 0 || y >= C) {
      atomicAdd(confusion + (long long)C * C, 1ull);
      continue;
    }
    atomicAdd(confusion + y * C + pred, 1ull);
    if (scores != nullptr && base + b < (unsigned long long)capacity) {
      float sum = 0.f;
      for (int c = 0; c < C; ++c) sum += expf(r[c] - best);
      scores[base + b] = expf(r[1] - best) / sum;
      score_labels[base + b] = (unsigned char)y;
    }
  }
  if (threadIdx.x == 0 && count != nullptr) *count = base + (unsigned long long)B;
}

// acc[0] += #(pos, neg) pairs with s_pos > s_neg, acc[1] += # ties, acc[2] += P, acc[3] += N over the first n samples
constexpr int AUC_TILE = 2048;
__global__ void __launch_bounds__(MET_THREADS)
    auroc_pairs_kernel(const float* __restrict__ scores, const unsigned char* __restrict__ labels,
                       const unsigned long long* __restrict__ count, long long capacity, unsigned long long* __restrict__ acc) {
  __shared__ float s_sc[AUC_TILE];
  __shared__ unsigned char s_lb[AUC_TILE];
  __shared__ unsigned long long red[4][MET_THREADS / 32];
  long long n = (long long)*count;
  if (n > capacity) n = capacity;
  unsigned long long wins = 0, ties = 0, npos = 0, nneg = 0;
  for (long long i0 = (long long)blockIdx.x * MET_THREADS; i0 < n; i0 += (long long)gridDim.x * MET_THREADS) {
    const long long i = i0 + threadIdx.x;
    const bool valid = i < n;
    const float si = valid ? scores[i] : 0.f;
    const bool pos = valid && labels[i] != 0;
    npos += pos ? 1 : 0;
    nneg += (valid && !pos) ? 1 : 0;
    for (long long j0 = 0; j0 < n; j0 += AUC_TILE) {
      const int m = (int)min((long long)AUC_TILE, n - j0);
      __syncthreads();
      for (int t = threadIdx.x; t < m; t += MET_THREADS) {
        s_sc[t] = scores[j0 + t];
        s_lb[t] = labels[j0 + t];
      }
      __syncthreads();
      if (pos) {
        unsigned int w = 0, e = 0;   // per tile: < 2^32
        for (int t = 0; t < m; ++t) {
          const bool neg = s_lb[t] == 0;
          w += (neg && si > s_sc[t]) ? 1u : 0u;
          e += (neg && si == s_sc[t]) ? 1u : 0u;
        }
        wins += w;
        ties += e;
      }
    }
  }
  unsigned long long v[4] = {wins, ties, npos, nneg};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned long long x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned long long x = 0;
    for (int w = 0; w < MET_THREADS / 32; ++w) x += red[threadIdx.x][w];
    if (x != 0) atomicAdd(acc + threadIdx.x, x);
  }
}

// out[0] = AUROC (0 when either class is absent, as torchmetrics), out[1] = P, out[2] = N, out[3] = ties
__global__ void auroc_finalize_kernel(const unsigned long long* __restrict__ acc, double* __restrict__ out) {
  const double P = (double)acc[2], N = (double)acc[3];
  out[0] = (P > 0 && N > 0) ? ((double)acc[0] + 0.5 * (double)acc[1]) / (P * N) : 0.0;
  out[1] = P;
  out[2] = N;
  out[3] = (double)acc[1];
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_metrics_update(const float* logits, const int64_t* labels, int32_t B, int32_t C, int64_t* confusion,
                                   float* scores, uint8_t* score_labels, int64_t* count, int64_t capacity, void* stream) {
  VITK_CHECK_ARG(logits && labels && confusion && B > 0 && C >= 2 && C <= 1024, "vitk_metrics_update: bad args");
  VITK_CHECK_ARG(scores == nullptr || (C == 2 && score_labels && count && capacity > 0),
                 "vitk_metrics_update: the score buffer (binary AUROC) needs C == 2, score_labels, count and capacity");
  metrics_update_kernel<<<1, MET_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, reinterpret_cast<const long long*>(labels), B, C, reinterpret_cast<unsigned long long*>(confusion), scores,
      score_labels, reinterpret_cast<unsigned long long*>(count), capacity);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_binary_auroc(const float* scores, const uint8_t* score_labels, const int64_t* count, int64_t capacity,
                                 int64_t* scratch, double* out, void* stream) {
  VITK_CHECK_ARG(scores && score_labels && count && scratch && out && capacity > 0, "vitk_binary_auroc: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VITK_CUDA(cudaMemsetAsync(scratch, 0, 4 * sizeof(int64_t), st));
  long long blocks = (capacity + MET_THREADS - 1) / MET_THREADS;
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  auroc_pairs_kernel<<<(unsigned)blocks, MET_THREADS, 0, st>>>(scores, score_labels, reinterpret_cast<const unsigned long long*>(count),
                                                              capacity, reinterpret_cast<unsigned long long*>(scratch));
  VITK_LAUNCH_CHECK();
  auroc_finalize_kernel<<<1, 1, 0, st>>>(reinterpret_cast<const unsigned long long*>(scratch), out);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
