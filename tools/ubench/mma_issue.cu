// tcgen05.mma issue / execution cost for small attention-shaped MMAs (sm_100a), one CTA per SM:
//   groups of NM MMAs (M=128, N, K=16, kind::f16, SS or TS) followed by one tcgen05.commit, G groups back to back,
//   the issuing thread never waits for completion until the end.  Reports cycles per group for the issue loop alone
//   (how long the thread is blocked) and for issue + drain (execution throughput).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../maest_b200/csrc -o mma_issue.bin mma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#define MAEST_F16 0
#define MAEST_BF16 1

#include "common.cuh"
using namespace mb;

#define G 512

// ALT = 1: consecutive MMAs of a group alternate between TWO accumulators (independent chains): separates the per-instruction
// issue cost from the latency of a chain of MMAs that accumulate into the same TMEM tile.
template <int N, int NM, int TS, int WARPS, int ALT = 0>
__global__ void bench(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  constexpr uint32_t idesc = make_idesc(DT_F16, 128, N, 0, 0);
  const uint64_t ad = make_sdesc(smem_u32(smem), 16, 1024), bd = make_sdesc(smem_u32(smem + 16384), 16, 1024);
  long long t_issue = 0, t_all = 0;
  if (warp < WARPS && lane == 0) {
    const uint32_t d = tb + warp * 128;
    const long long t0 = clock64();
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int k = 0; k < NM; ++k) {
        const uint32_t dk = ALT ? d + uint32_t((k & 1) * 64) : d;
        const uint32_t acc = ALT ? (k > 1 ? 1u : 0u) : (k ? 1u : 0u);
        if (TS) mma_ts(dk, tb + 256 + 8 * (k & 7), bd + 2 * (k & 3), idesc, acc);
        else mma_ss(dk, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, acc);
      }
      tc_commit(&bars[warp * 2 + (g & 1)]);
    }
    const long long t1 = clock64();
    // drain: wait for the last commit of each parity (G/2 completions each: parity of the last = ((G/2)-1)&1)
    mbar_wait(&bars[warp * 2 + 0], ((G / 2) - 1) & 1);
    mbar_wait(&bars[warp * 2 + 1], ((G / 2) - 1) & 1);
    const long long t2 = clock64();
    t_issue = t1 - t0; t_all = t2 - t0;
    out[(blockIdx.x * 4 + warp) * 2] = t_issue; out[(blockIdx.x * 4 + warp) * 2 + 1] = t_all;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

template <int N, int NM, int TS, int WARPS, int ALT = 0>
void run(const char* name) {
  long long* out; cudaMalloc(&out, 148 * 8 * 8); cudaMemset(out, 0, 148 * 8 * 8);
  cudaFuncSetAttribute(bench<N, NM, TS, WARPS, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int r = 0; r < 2; ++r) { bench<N, NM, TS, WARPS, ALT><<<148, 128, 65536>>>(out); cudaDeviceSynchronize(); }
  long long h[148 * 8]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double ti = 0, ta = 0; for (int b = 0; b < 148; ++b) { ti += h[b * 8]; ta += h[b * 8 + 1]; }
  cudaError_t e = cudaGetLastError();
  printf("%-52s issue %7.1f cyc/group, issue+drain %7.1f cyc/group (%5.1f per MMA; floor %d)%s\n", name, ti / 148 / G, ta / 148 / G,
         ta / 148 / G / NM, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  run<128, 4, 0, 1>("SS N=128 x4 + commit (QK^T tile of 128 keys)");
  run<64, 4, 0, 1>("SS N=64  x4 + commit (QK^T tile of 64 keys)");
  run<64, 8, 1, 1>("TS N=64  x8 + commit (PV, 128 keys)");
  run<64, 4, 1, 1>("TS N=64  x4 + commit (PV, 64 keys)");
  run<64, 1, 1, 1>("TS N=64  x1 + commit");
  run<64, 1, 0, 1>("SS N=64  x1 + commit");
  run<256, 1, 0, 1>("SS N=256 x1 + commit");
  run<128, 4, 0, 2>("SS N=128 x4 + commit, two issuing warps");
  run<64, 4, 1, 2>("TS N=64  x4 + commit, two issuing warps");
  run<64, 8, 1, 1, 1>("TS N=64  x8 + commit, alternating two accumulators");
  run<64, 8, 0, 1, 1>("SS N=64  x8 + commit, alternating two accumulators");
  run<32, 4, 0, 1>("SS N=32  x4 + commit");
  run<16, 4, 0, 1>("SS N=16  x4 + commit");
  return 0;
}
