// mufu_probe.cu -- issue / pipe throughput of the softmax inner-loop instructions on one SM (sm_100a):
// MUFU.EX2, F2FP.F16.F32.PACK_AB, FFMA2, FADD2 and their mix, at 1 / 2 / 4 warps per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mufu_probe tools/probes/mufu_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, long long* cycles, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0.001f * (threadIdx.x + i);
  float2 acc = make_float2(0.f, 0.f);
  unsigned pk = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float a = x[i], b = x[i + 1];
      if (MODE == 0 || MODE == 2 || MODE == 4 || MODE == 5) {   // ex2
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      }
      if (MODE == 1 || MODE == 2 || MODE == 4) {                // pack to f16x2
        unsigned r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
        pk ^= r;
      }
      if (MODE == 3 || MODE == 4 || MODE == 5) {                // ffma2 + fadd2
        float2 v = make_float2(a, b);
        v = __ffma2_rn(v, make_float2(1.0001f, 1.0001f), make_float2(-0.5f, -0.5f));
        acc = __fadd2_rn(acc, v);
        a = v.x; b = v.y;
      }
      if (MODE == 6) {                                          // bf16 pack
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
        pk ^= r;
      }
      if (MODE == 7) {                                          // integer-ALU pack of two fp32 -> bf16x2 by truncation (PRMT)
        pk ^= __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632);
        a += 1.f;
      }
      x[i] = a * 0.5f; x[i + 1] = b * 0.5f;
    }
  }
  const long long t1 = clock64();
  float s = acc.x + acc.y + __uint_as_float(pk & 0xff);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_elem_pairs) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    probe<MODE><<<1, warps * 32, 0>>>(out, cyc, iters);
    probe<MODE><<<1, warps * 32, 0>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // per SM sub-partition: (warps/4) warps, each did iters*8 pairs
    const double pairs_per_smsp = double(warps / 4) * iters * 8;
    printf("%-38s warps/SMSP=%d  %.2f clk per element-pair per SMSP (=%.2f clk per warp-pair-instr group)\n", name, warps / 4,
           c / pairs_per_smsp, c / pairs_per_smsp);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("2x ex2", 1);
  run<1>("1x cvt.f16x2", 1);
  run<6>("1x cvt.bf16x2", 1);
  run<7>("1x prmt (+fadd)", 1);
  run<2>("2x ex2 + cvt.f16x2", 1);
  run<3>("ffma2 + fadd2 (+2 fmul)", 1);
  run<5>("2x ex2 + ffma2 + fadd2", 1);
  run<4>("2x ex2 + cvt + ffma2 + fadd2", 1);
  return 0;
}
