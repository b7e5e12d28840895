// teacher.cu -- the one memory-bound pass of the FROZEN DenseNet169 teacher forward (distillation step,
// lightning_modules.py:943-947: `self.teacher(images)` under no_grad, eval mode).  In eval mode every BatchNorm is a
// per-channel affine map, and DenseNet applies it, followed by ReLU, to the CONCATENATION of all earlier feature maps of a
// dense block (torchvision densenet.py _DenseLayer: norm1 -> relu1 -> conv1 -> norm2 -> relu2 -> conv2).  PyTorch eager runs
// torch.cat + batch_norm + relu = three read+write passes over the growing concatenation per layer (measured on B200,
// batch 256, bf16 channels_last: batch_norm 37 %, cat/copy 29 %, relu 11 % of the 29 ms forward; the convolutions are 16 %).
// Here the block's features live in ONE preallocated NHWC buffer and this kernel reads the first C channels of every pixel
// (row pitch = the buffer's full channel count), applies y = max(0, x * scale[c] + shift[c]) and writes the compact NHWC
// operand of the following cuDNN convolution: one read + one write, no concatenation copy.
#include "vitk_common.cuh"

namespace vitk {
namespace {

// 8 channels (16 bytes) per thread; x_ld / y_ld are the per-pixel pitches in elements (multiples of 8)
__global__ void __launch_bounds__(256)
    affine_relu_nhwc_kernel(const uint4* __restrict__ x, long long x_ld8, uint4* __restrict__ y, long long y_ld8,
                            const float* __restrict__ scale, const float* __restrict__ shift, long long pixels, int cv, int fp16,
                            int relu) {
  const long long total = pixels * cv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long p = i / cv;
    const int v = int(i - p * cv);
    const uint4 in = __ldg(x + p * x_ld8 + v);
    const float4 s0 = ldg_f4(scale + 8 * v), s1 = ldg_f4(scale + 8 * v + 4);
    const float4 h0 = ldg_f4(shift + 8 * v), h1 = ldg_f4(shift + 8 * v + 4);
    const float2 a = unpack16(in.x, fp16), b = unpack16(in.y, fp16), c = unpack16(in.z, fp16), d = unpack16(in.w, fp16);
    float o[8] = {fmaf(a.x, s0.x, h0.x), fmaf(a.y, s0.y, h0.y), fmaf(b.x, s0.z, h0.z), fmaf(b.y, s0.w, h0.w),
                  fmaf(c.x, s1.x, h1.x), fmaf(c.y, s1.y, h1.y), fmaf(d.x, s1.z, h1.z), fmaf(d.y, s1.w, h1.w)};
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaxf(o[k], 0.f);
    }
    y[p * y_ld8 + v] = make_uint4(pack16(o[0], o[1], fp16), pack16(o[2], o[3], fp16), pack16(o[4], o[5], fp16), pack16(o[6], o[7], fp16));
  }
}

// k x k / stride s / zero- or -inf-padded pooling of an NHWC tensor, 8 channels per thread, written with pixel pitch y_ld:
// the stem's MaxPool2d(3, 2, 1) and the transitions' AvgPool2d(2, 2) store straight into channels [0, C) of the NEXT dense
// block's concatenation buffer (no separate copy).  Sums in fp32, one rounding (what ATen does for 16-bit pooling).
__global__ void __launch_bounds__(256)
    pool_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long y_ld8, int B, int H, int W, int cv, int OH, int OW,
                     int k, int s, int pad, int is_max, int fp16) {
  const long long total = (long long)B * OH * OW * cv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float inv = 1.f / float(k * k);          // count_include_pad = True (AvgPool2d default)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = int(i % cv);
    long long r = i / cv;
    const int ow = int(r % OW);
    r /= OW;
    const int oh = int(r % OH);
    const int b = int(r / OH);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = is_max ? -INFINITY : 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oh * s - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ow * s - pad + kx;
        if (ix < 0 || ix >= W) continue;
        const uint4 in = __ldg(x + (((long long)b * H + iy) * W + ix) * cv + v);
        const float2 p0 = unpack16(in.x, fp16), p1 = unpack16(in.y, fp16), p2 = unpack16(in.z, fp16), p3 = unpack16(in.w, fp16);
        const float f[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = is_max ? fmaxf(acc[e], f[e]) : acc[e] + f[e];
      }
    }
    if (!is_max) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= inv;
    }
    y[(((long long)b * OH + oh) * OW + ow) * y_ld8 + v] =
        make_uint4(pack16(acc[0], acc[1], fp16), pack16(acc[2], acc[3], fp16), pack16(acc[4], acc[5], fp16), pack16(acc[6], acc[7], fp16));
  }
}


// Patch rows of a k x k / stride s / zero-padded convolution over an NHWC tensor with FEW channels (the 3-channel 7x7 stem):
// row (b, oy, ox) = [ky][kx][c] -> k*k*C elements, padded with zeros to `ld` columns (a multiple of 8).  One thread gathers
// the 8 elements of one 16-byte cell of a row (the input is small and cached; a per-block shared table turns the column
// index into (ky, kx, input offset)) and writes it with a single coalesced 16-byte store: the kernel is bound by the
// ld * 2 bytes per output pixel it writes.  Feeds vitk_gemm: cuDNN runs this 3-channel convolution on a pre-tensor-core
// implicit-GEMM kernel (1.8 ms at batch 256).
__global__ void __launch_bounds__(256)
    im2col_rows_kernel(const uint16_t* __restrict__ x, uint4* __restrict__ patches, int B, int H, int W, int C, int k, int s, int pad,
                       int OH, int OW, int ld) {
  extern __shared__ int im2col_tab[];            // [ld] input offset (ky*W + kx)*C + c, [ld] (ky << 8) | kx, -1 beyond k*k*C
  int* t_off = im2col_tab;
  int* t_kk = im2col_tab + ld;
  for (int e = threadIdx.x; e < ld; e += blockDim.x) {
    if (e < k * k * C) {
      const int ky = e / (k * C), rem = e - ky * (k * C), kx = rem / C, c = rem - kx * C;
      t_off[e] = (ky * W + kx) * C + c;
      t_kk[e] = (ky << 8) | kx;
    } else {
      t_off[e] = 0;
      t_kk[e] = -1;
    }
  }
  __syncthreads();
  const int cells = ld >> 3;
  const long long total = (long long)B * OH * OW * cells;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cell = int(i % cells);
    const long long r = i / cells;
    const int ox = int(r % OW);
    const long long r2 = r / OW;
    const int oy = int(r2 % OH);
    const int b = int(r2 / OH);
    const int iy0 = oy * s - pad, ix0 = ox * s - pad;
    const uint16_t* base = x + (((long long)b * H + iy0) * W + ix0) * C;       // may point before the image: only valid taps are read
    uint32_t w[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      uint32_t v2 = 0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int e = cell * 8 + 2 * h + q;
        const int kk = t_kk[e];
        uint32_t v = 0;
        if (kk >= 0) {
          const int iy = iy0 + (kk >> 8), ix = ix0 + (kk & 255);
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(base + t_off[e]);
        }
        v2 |= v << (16 * q);
      }
      w[h] = v2;
    }
    patches[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_pool_nhwc(const void* x, void* y, int64_t y_ld, int32_t B, int32_t H, int32_t W, int32_t C, int32_t kernel,
                              int32_t stride, int32_t pad, int32_t is_max, int32_t dtype, void* stream) {
  VITK_CHECK_ARG(x && y, "vitk_pool_nhwc: null pointer");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_pool_nhwc: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && y_ld >= C && y_ld % 8 == 0, "vitk_pool_nhwc: C=%d, pitch must be multiples of 8", C);
  VITK_CHECK_ARG(kernel > 0 && stride > 0 && pad >= 0 && 2 * pad <= kernel && H + 2 * pad >= kernel && W + 2 * pad >= kernel,
                 "vitk_pool_nhwc: bad window (kernel=%d stride=%d pad=%d)", kernel, stride, pad);
  VITK_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0), "vitk_pool_nhwc: 16-byte aligned pointers required");
  const int OH = (H + 2 * pad - kernel) / stride + 1, OW = (W + 2 * pad - kernel) / stride + 1;   // ceil_mode = False
  const long long total = (long long)B * OH * OW * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  pool_nhwc_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), y_ld / 8, B, H, W, C / 8, OH, OW, kernel, stride, pad, is_max,
      int(dtype == VITK_FP16));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_affine_relu_nhwc(const void* x, int64_t x_ld, void* y, int64_t y_ld, const float* scale, const float* shift,
                                     int64_t pixels, int32_t C, int32_t dtype, int32_t relu, void* stream) {
  VITK_CHECK_ARG(x && y && scale && shift, "vitk_affine_relu_nhwc: null pointer");
  VITK_CHECK_ARG(dtype == VITK_BF16 || dtype == VITK_FP16, "vitk_affine_relu_nhwc: dtype must be bf16 or fp16");
  VITK_CHECK_ARG(pixels > 0 && C > 0 && C % 8 == 0 && x_ld >= C && y_ld >= C && x_ld % 8 == 0 && y_ld % 8 == 0,
                 "vitk_affine_relu_nhwc: C=%d and the pixel pitches must be multiples of 8, pitches >= C", C);
  VITK_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0), "vitk_affine_relu_nhwc: 16-byte aligned pointers required");
  const int cv = C / 8;
  const long long total = (long long)pixels * cv;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  affine_relu_nhwc_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x), x_ld / 8, reinterpret_cast<uint4*>(y), y_ld / 8, scale, shift, (long long)pixels, cv,
      int(dtype == VITK_FP16), relu);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

extern "C" int vitk_im2col_rows(const void* x, void* patches, int32_t B, int32_t H, int32_t W, int32_t C, int32_t kernel,
                                int32_t stride, int32_t pad, int64_t ld, void* stream) {
  VITK_CHECK_ARG(x && patches && B > 0 && H > 0 && W > 0 && C > 0 && kernel > 0 && stride > 0 && pad >= 0, "vitk_im2col_rows: bad args");
  VITK_CHECK_ARG(ld >= (int64_t)kernel * kernel * C && ld % 8 == 0, "vitk_im2col_rows: ld must be a multiple of 8, >= k*k*C");
  const int OH = (H + 2 * pad - kernel) / stride + 1, OW = (W + 2 * pad - kernel) / stride + 1;
  VITK_CHECK_ARG(OH > 0 && OW > 0, "vitk_im2col_rows: empty output");
  const long long total = (long long)B * OH * OW * (ld / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  VITK_CHECK_ARG(kernel < 256 && ld <= 4096 && ((uintptr_t)patches % 16 == 0), "vitk_im2col_rows: kernel < 256, ld <= 4096, 16-byte aligned output");
  im2col_rows_kernel<<<(unsigned)blocks, 256, 2 * (size_t)ld * sizeof(int), reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint16_t*>(x), reinterpret_cast<uint4*>(patches), B, H, W, C, kernel, stride, pad, OH, OW, (int)ld);
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}
