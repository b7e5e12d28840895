// tmem_probe.cu -- how fast can warps read TMEM?  (tcgen05.ld 32x32b.x32: 4 KB per warp-instruction)
// For W warps (1..16) per CTA, one CTA per SM: each warp issues ITERS loads with DEPTH of them in flight before
// tcgen05.wait::ld.  Prints cycles per load and bytes/clk per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

template <int DEPTH>
__global__ void probe(long long* out, uint32_t* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i += DEPTH) {
    uint32_t v[DEPTH][32];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) ld32(tb + uint32_t(((i + d) * 32) & 255), v[d]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
      for (int j = 0; j < 32; ++j) acc ^= v[d][j];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 8);
  cudaMalloc(&sink, 4);
  const int iters = 240;
  printf("warps depth cycles/load(per warp) bytes/clk/SM\n");
  for (int w : {1, 2, 4, 8, 12, 16}) {
    for (int depth : {1, 2, 4}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (depth == 1) probe<1><<<148, 32 * w>>>(out, sink, iters);
        if (depth == 2) probe<2><<<148, 32 * w>>>(out, sink, iters);
        if (depth == 4) probe<4><<<148, 32 * w>>>(out, sink, iters);
        cudaDeviceSynchronize();
      }
      long long c;
      cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      printf("%5d %5d %10.1f %12.1f\n", w, depth, double(c) / iters, double(iters) * w * 4096.0 / double(c));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
