// TMEM read / write bandwidth per SM (sm_100a): w warps per sub-partition issue tcgen05.ld / tcgen05.st 32x32b.xN back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem.bin tmem.cu && ./tmem.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

template <int X, int MODE, int WAIT_EVERY>
__global__ void bench(long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) r[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it += WAIT_EVERY) {
#pragma unroll
    for (int u = 0; u < WAIT_EVERY; ++u) {
      const uint32_t a = base + uint32_t(((it + u) * X + (warp >> 2) * 64) & 511 & ~(X - 1));
      if (MODE == 0) {
        if (X == 32)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                       : "r"(a));
        else
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                       : "r"(a));
      } else {
        if (X == 32)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
                       ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
                       "r"(a) : "memory");
        else
          asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
                       ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                       "r"(a) : "memory");
      }
    }
    if (MODE == 0) {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] ^ r[X - 1];
    } else {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 32 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = float(acc);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

template <int X, int MODE, int WE>
void run(const char* name) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 32 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
  printf("%-40s", name);
  for (int w : {1, 2, 3, 6}) {
    const int threads = 128 * w;
    for (int rep = 0; rep < 2; ++rep) { bench<X, MODE, WE><<<148, threads>>>(cyc, sink); cudaDeviceSynchronize(); }
    long long h[148 * 32];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int b = 0; b < 148; ++b) { long long m = 0; for (int k = 0; k < 4 * w; ++k) m = h[b * 32 + k] > m ? h[b * 32 + k] : m; avg += m; }
    avg /= 148;
    const double bytes = double(ITERS) * X * 4 * 32 * 4 * w;     // per SM
    printf("  w=%d: %7.1f B/clk/SM (%5.1f clk/instr/SMSP)", w, bytes / avg, avg / (double(ITERS) * w));
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf(" ERR %s", cudaGetErrorString(e));
  }
  printf("\n");
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  run<32, 0, 1>("ld x32, wait every ld");
  run<32, 0, 2>("ld x32, wait every 2");
  run<16, 0, 1>("ld x16, wait every ld");
  run<16, 0, 4>("ld x16, wait every 4");
  run<32, 1, 1>("st x32, wait every st");
  run<32, 1, 2>("st x32, wait every 2");
  run<16, 1, 4>("st x16, wait every 4");
  return 0;
}
