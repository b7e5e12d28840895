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

}  // namespace
}  // namespace vitk

using namespace vitk;

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
