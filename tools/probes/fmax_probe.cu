// fmax_probe.cu -- cost of a 128-element row maximum held in registers (sm_100a): FMNMX (2-input) vs FMNMX3 (3-input),
// 4 / 8 / 16 independent chains, one warp per SM sub-partition and two.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/fmax_probe tools/probes/fmax_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int MODE, int CH>
__global__ void probe(const float* in, float* out, long long* cycles, int iters) {
  float x[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) x[i] = in[(threadIdx.x * 131 + i * 7) & 1023];
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float m[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) m[c] = -1e30f;
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 128; ++i) m[i % CH] = fmaxf(m[i % CH], x[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 128; i += 2) m[(i / 2) % CH] = max3(m[(i / 2) % CH], x[i], x[i + 1]);
    }
    float r = m[0];
#pragma unroll
    for (int c = 1; c < CH; ++c) r = fmaxf(r, m[c]);
    acc += r;
#pragma unroll
    for (int i = 0; i < 128; i += 16) x[i] += acc * 1e-9f;   // keep the loop from being hoisted
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE, int CH>
void run(const char* name) {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  cudaMemset(in, 0, 4096);
  const int iters = 1000;
  for (int warps : {4, 8}) {
    probe<MODE, CH><<<1, warps * 32>>>(in, out, cyc, iters);
    probe<MODE, CH><<<1, warps * 32>>>(in, out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-24s chains=%2d warps/SMSP=%d  %.0f clk per 128-element row max\n", name, CH, warps / 4, double(c) / iters);
  }
}

int main() {
  run<0, 4>("FMNMX"); run<0, 8>("FMNMX"); run<0, 16>("FMNMX");
  run<1, 4>("FMNMX3"); run<1, 8>("FMNMX3"); run<1, 16>("FMNMX3");
  return 0;
}
