// Pipe-throughput microbenchmark for sm_100a: cycles per warp-instruction per SM sub-partition for the instruction
// kinds the softmax / GELU inner loops are made of, at 1, 2 and 4 resident warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITERS 512
#define UNROLL 16

template <int KIND>
__global__ void bench(float* out, long long* cycles, float seed) {
  float a[UNROLL];
  float b = seed * 1.0001f, c = seed * 0.5f;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (KIND == 0) a[i] = fmaf(a[i], b, c);                       // FFMA reg,reg,reg
      if (KIND == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);              // FFMA imm
      if (KIND == 2) a[i] = a[i] + b;                               // FADD
      if (KIND == 3) a[i] = fmaxf(a[i], b);                         // FMNMX
      if (KIND == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }   // MUFU.EX2
      if (KIND == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[(i + 8) % UNROLL] = a[(i + 8) % UNROLL] + b; }  // MUFU + FADD
      if (KIND == 6) { __half2 h = __floats2half2_rn(a[i], a[(i + 1) % UNROLL]); acc ^= *reinterpret_cast<unsigned*>(&h); }            // F2FP (+LOP)
      if (KIND == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[(i + 8) % UNROLL] = fmaf(a[(i + 8) % UNROLL], b, c); a[(i + 4) % UNROLL] = a[(i + 4) % UNROLL] + c; }  // MUFU + FFMA + FADD
      if (KIND == 8) { a[i] = fmaf(a[i], b, c); a[(i + 8) % UNROLL] = fmaxf(a[(i + 8) % UNROLL], c); }   // FFMA + FMNMX (fma + alu pipes)
      if (KIND == 10) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); __half2 h = __floats2half2_rn(a[(i + 8) % UNROLL], a[(i + 9) % UNROLL]); acc += *reinterpret_cast<unsigned*>(&h); }   // MUFU + F2FP + IADD
      if (KIND == 11) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); float m; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a[(i + 8) % UNROLL]), "f"(a[(i + 9) % UNROLL]), "f"(b)); b = m; }   // MUFU + FMNMX3
      if (KIND == 12) { __half2 h = __floats2half2_rn(a[i], a[(i + 1) % UNROLL]); acc += *reinterpret_cast<unsigned*>(&h); a[i] = a[i] + c; }   // F2FP + IADD + FADD
      if (KIND == 13) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[(i + 8) % UNROLL] = fmaf(a[(i + 8) % UNROLL], b, c); a[(i + 4) % UNROLL] = a[(i + 4) % UNROLL] + c;
                        if (i & 1) { __half2 h = __floats2half2_rn(a[(i + 2) % UNROLL], a[(i + 3) % UNROLL]); acc += *reinterpret_cast<unsigned*>(&h); } }   // the softmax inner loop mix: MUFU + FFMA + FADD + 0.5 (F2FP + IADD)
      if (KIND == 9) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }   // MUFU.RCP
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + float(acc);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int per_inst) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  printf("%-28s", name);
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp;
    bench<KIND><<<148, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    bench<KIND><<<148, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // cycles per (group of per_inst instructions) per warp-slot: total instr groups per SMSP = ITERS*UNROLL*warps_per_smsp
    printf("  w=%d: %6.2f cyc/grp/SMSP", warps_per_smsp, avg / (double(ITERS) * UNROLL * warps_per_smsp));
  }
  printf("   (%d instr per group)\n", per_inst);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA r,r,r", 1);
  run<1>("FFMA r,imm,imm", 1);
  run<2>("FADD", 1);
  run<3>("FMNMX", 1);
  run<4>("MUFU.EX2", 1);
  run<9>("MUFU.RCP", 1);
  run<5>("MUFU.EX2 + FADD", 2);
  run<6>("F2FP.PACK + LOP", 2);
  run<7>("MUFU.EX2 + FFMA + FADD", 3);
  run<8>("FFMA + FMNMX", 2);
  run<10>("MUFU.EX2 + F2FP + IADD", 3);
  run<11>("MUFU.EX2 + FMNMX3", 2);
  run<12>("F2FP + IADD + FADD", 3);
  run<13>("softmax mix (4 instr)", 4);
  return 0;
}
