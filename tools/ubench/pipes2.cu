// Round-2 pipe microbenchmarks for the attention softmax redesign (sm_100a):
//   * packed fp32x2 FMA / ADD (FFMA2 / FADD2) issue and pipe rate,
//   * MUFU.EX2 on f16 / bf16 inputs vs f32,
//   * the candidate per-pair softmax sequences (MUFU path, polynomial path, 5:3 mix) at 1..4 warps per sub-partition,
//   * L2 -> SM bandwidth of 16 KB TMA-less bulk loads (cp.async.bulk) that all hit in L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu && ./pipes2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITERS 256
#define UNROLL 16

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <int KIND>
__global__ void bench(float* out, long long* cycles, float seed) {
  u64 a[UNROLL];
  const u64 b = pk2(seed * 1.0001f, seed * 0.9999f), c = pk2(seed * 0.5f, seed * 0.25f);
  const u64 magic = pk2(12582912.f, 12582912.f), nmagic = pk2(-12582912.f, -12582912.f);
  const u64 c3 = pk2(0.0555f, 0.0555f), c2 = pk2(0.2402f, 0.2402f), c1 = pk2(0.6931f, 0.6931f), c0 = pk2(1.f, 1.f);
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) a[i] = pk2(seed + i * 0.001f + threadIdx.x * 1e-6f, seed - i * 0.001f);
  u64 l = pk2(0.f, 0.f);
  uint32_t acc = 0;
  float mx = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (KIND == 0) a[i] = fma2(a[i], b, c);                       // FFMA2
      if (KIND == 1) a[i] = add2(a[i], b);                          // FADD2
      if (KIND == 2) { float lo, hi; up2(a[i], lo, hi); lo = fmaf(lo, 1.0001f, 0.5f); hi = fmaf(hi, 1.0001f, 0.5f); a[i] = pk2(lo, hi); }   // 2 x FFMA
      if (KIND == 3) { uint32_t x = uint32_t(a[i]), y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); a[i] = (a[i] & 0xffffffff00000000ull) | y; }   // 2 x MUFU.EX2.F16 + PRMT
      if (KIND == 4) { uint32_t x = uint32_t(a[i]), y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); a[i] = (a[i] & 0xffffffff00000000ull) | y; }   // 2 x MUFU.EX2.BF16 + PRMT
      if (KIND == 5) { float lo, hi; up2(a[i], lo, hi); a[i] = pk2(ex2f(lo), ex2f(hi)); }   // 2 x MUFU.EX2 f32
      if (KIND == 6) {   // MUFU-path pair: FFMA2 + 2 MUFU + FADD2 + F2FP
        const u64 t = fma2(a[i], b, c); float lo, hi; up2(t, lo, hi); const float p0 = ex2f(lo), p1 = ex2f(hi);
        l = add2(l, pk2(p0, p1)); acc += cvt_h2(p0, p1);
      }
      if (KIND == 7 || (KIND == 8 && (i % 8) >= 5)) {   // polynomial-path pair (KIND 8: 3 of every 8 pairs)
        const u64 x = fma2(a[i], b, c);
        float lo, hi; up2(x, lo, hi);
        mx = fmaxf(mx, fmaxf(fabsf(lo), fabsf(hi)));
        const u64 t = add2(x, magic);
        const u64 f = add2(t, nmagic);
        const u64 r = sub2(x, f);
        u64 q = fma2(c3, r, c2); q = fma2(q, r, c1); q = fma2(q, r, c0);
        float q0, q1, t0f, t1f; up2(q, q0, q1); up2(t, t0f, t1f);
        const float p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0f) << 23));
        const float p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1f) << 23));
        l = add2(l, pk2(p0, p1)); acc += cvt_h2(p0, p1);
      } else if (KIND == 8) {
        const u64 t = fma2(a[i], b, c); float lo, hi; up2(t, lo, hi); const float p0 = ex2f(lo), p1 = ex2f(hi);
        l = add2(l, pk2(p0, p1)); acc += cvt_h2(p0, p1);
      }
      if (KIND == 9) { float lo, hi; up2(a[i], lo, hi); acc += cvt_h2(lo, hi); }   // F2FP + IADD
      if (KIND == 10) { float lo, hi; up2(a[i], lo, hi); int x = __float_as_int(lo) * 8388608 + __float_as_int(hi); a[i] = pk2(__int_as_float(x), hi); }  // IMAD
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) { float lo, hi; up2(a[i], lo, hi); s += lo + hi; }
  float l0, l1; up2(l, l0, l1);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + float(acc) + l0 + l1 + mx;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  printf("%-44s", name);
  for (int w : {1, 2, 3, 4}) {
    const int threads = 128 * w;
    bench<KIND><<<148, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    bench<KIND><<<148, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("  w=%d: %6.2f", w, avg / (double(ITERS) * UNROLL * w));
  }
  printf("   cycles per PAIR-group per warp-slot per SMSP\n");
  cudaFree(out); cudaFree(cyc);
}

// ---------------------------------------------------------------- L2 -> SM bandwidth (bulk async copies of 16 KB, all L2 hits)
__global__ void l2bw(const uint8_t* src, size_t span, int tiles_per_cta, long long* cycles, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4 * 16384);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    // 4 copies in flight; each CTA walks a window shared with 13 neighbours (like 14 q-tiles of one (clip, head))
    const size_t base = (size_t(blockIdx.x / 14) * 431360) % span;
    for (int t = 0; t < tiles_per_cta + 4; ++t) {
      const int s = t & 3;
      if (t >= 4) {
        const uint32_t ph = ((t - 4) >> 2) & 1;
        asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(bar0 + 8 * s), "r"(ph));
      }
      if (t < tiles_per_cta) {
        const size_t off = (base + size_t(t % 26) * 16384) % span;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(16384));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(sm + s * 16384)), "l"(src + off), "r"(16384), "r"(bar0 + 8 * s) : "memory");
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; sink[blockIdx.x] = float(sm[threadIdx.x]); }
}

int main() {
  run<0>("FFMA2");
  run<1>("FADD2");
  run<2>("2 x FFMA imm");
  run<5>("2 x MUFU.EX2 f32");
  run<3>("ex2.f16x2 (2 MUFU.EX2.F16 + PRMT)");
  run<4>("ex2.bf16x2 (2 MUFU.EX2.BF16 + PRMT)");
  run<9>("F2FP + IADD");
  run<10>("IMAD + (movs)");
  run<6>("pair, MUFU path (FFMA2 2MUFU FADD2 F2FP)");
  run<7>("pair, polynomial path");
  run<8>("pairs, 5 MUFU : 3 polynomial");
  {
    const size_t span = 64u << 20;
    uint8_t* src; long long* cyc; float* sink;
    cudaMalloc(&src, span + (1 << 20)); cudaMemset(src, 1, span + (1 << 20));
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 4);
    cudaFuncSetAttribute(l2bw, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 64);
    const int tiles = 2048;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      l2bw<<<148, 32, 4 * 16384 + 64>>>(src, span, tiles, cyc, sink);
      cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("L2->SM bulk copies: %.1f B/clk/SM (%.0f B/clk chip), %.2f TB/s (event time %.3f ms)\n", tiles * 16384.0 / avg,
             148 * tiles * 16384.0 / avg, 148.0 * tiles * 16384 / ms / 1e9, ms);
    }
  }
  return 0;
}
