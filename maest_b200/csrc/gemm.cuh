// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue).
//
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, 4-stage mbarrier ring)
//   warp 1      MMA issuer     (one elected lane: tcgen05.mma cta_group::1 kind::f16, 128x256x16, fp32 accum in TMEM)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, double-buffered against the epilogue)
//   warps 4..7  epilogue       (tcgen05.ld 32x32b, one accumulator row per thread, fused bias/GELU/residual/pos-embed)
//
// Replaces the cuBLAS(Lt)/cuDNN calls behind models/maest.py:250 (patch-embed conv as GEMM), :361 (qkv),
// :376 (proj), :203-206 (fc1, GELU, fc2) and the separate bias / residual-add / pos-embed passes (:418-419, :670-675).
#pragma once
#include "common.cuh"

namespace mb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
constexpr int GEMM_B_BYTES = GEMM_BN * GEMM_BK * 2;  // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

enum : int {
  EPI_STORE16 = 0,  // out16[r,n] = acc + bias[n]
  EPI_GELU16 = 1,   // out16[r,n] = gelu_erf(acc + bias[n])
  EPI_RESID32 = 2,  // out32[r,n] = resid32[r,n] + acc + bias[n]        (out32 may alias resid32)
  EPI_STORE32 = 3,  // out32[r,n] = acc + bias[n] + addend[(m % rows_per_group), n]
};

struct GemmParams {
  int M, N, K;
  const float* bias;     // [N] or null
  void* out;             // 16-bit or fp32, row stride ld_out elements
  const float* resid;    // EPI_RESID32
  const float* addend;   // EPI_STORE32: optional [rows_per_group, N] table (pos-embed), else null
  int ld_out;
  // output row remap:  r = (m / rows_per_group) * group_stride + row_offset + (m % rows_per_group)
  int rows_per_group, group_stride, row_offset;
};

// exact-erf GELU (models/maest.py:500 nn.GELU default).  erf via Abramowitz-Stegun 7.1.26
// (|abs err| <= 1.5e-7, below fp32 epsilon of the 0.5*x*(1+erf) product for |x| < 1).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = exp2f(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p, e, 1.0f);
  const float erf_v = copysignf(erf_abs, x);
  const float hx = 0.5f * x;
  return fmaf(hx, erf_v, hx);
}

template <int DT, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + GEMM_STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * GEMM_STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * GEMM_STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (p.N + GEMM_BN - 1) / GEMM_BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < GEMM_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / num_n) * GEMM_BM;
        const int n0 = (t % num_n) * GEMM_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * GEMM_STAGE_BYTES;
          uint8_t* sb = sa + GEMM_A_BYTES;
          mbar_expect_tx(&full_bar[stage], GEMM_STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * GEMM_BK, m0);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * GEMM_BK, n0);
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(DT, GEMM_BM, GEMM_BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * GEMM_BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * GEMM_STAGE_BYTES);
          const uint64_t adesc = make_sdesc(sa, 16, 1024);
          const uint64_t bdesc = make_sdesc(sa + GEMM_A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 K-elements = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            mma_ss(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    using O = Op16<DT == DT_BF16 ? DT_BF16 : DT_F16>;
    const int q = warp & 3;  // TMEM lane quarter owned by this warp
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / num_n) * GEMM_BM;
      const int n0 = (t % num_n) * GEMM_BN;
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      long r = 0;
      int pr = 0;
      if (row_ok) {
        pr = m % p.rows_per_group;
        r = long(m / p.rows_per_group) * p.group_stride + p.row_offset + pr;
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * GEMM_BN);
#pragma unroll 1
      for (int c = 0; c < GEMM_BN / 32; ++c) {
        const int n = n0 + c * 32;
        if (n >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(taddr + uint32_t(c * 32), v);
        tc_wait_ld();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
          }
        }
        if (row_ok) {
          if constexpr (EPI == EPI_STORE16 || EPI == EPI_GELU16) {
            if constexpr (EPI == EPI_GELU16) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_erf_fast(f[j]);
            }
            typename O::T* dst = reinterpret_cast<typename O::T*>(p.out) + r * p.ld_out + n;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              st_global_v4(dst + j, O::pack(f[j], f[j + 1]), O::pack(f[j + 2], f[j + 3]),
                           O::pack(f[j + 4], f[j + 5]), O::pack(f[j + 6], f[j + 7]));
            }
          } else if constexpr (EPI == EPI_RESID32) {
            const float* src = p.resid + r * p.ld_out + n;
            float* dst = reinterpret_cast<float*>(p.out) + r * p.ld_out + n;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 x4 = *reinterpret_cast<const float4*>(src + j);
              float4 o4;
              o4.x = x4.x + f[j]; o4.y = x4.y + f[j + 1]; o4.z = x4.z + f[j + 2]; o4.w = x4.w + f[j + 3];
              *reinterpret_cast<float4*>(dst + j) = o4;
            }
          } else {
            float* dst = reinterpret_cast<float*>(p.out) + r * p.ld_out + n;
            const float* add = p.addend ? p.addend + long(pr) * p.N + n : nullptr;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o4 = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              if (add) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(add + j));
                o4.x += a4.x; o4.y += a4.y; o4.z += a4.z; o4.w += a4.w;
              }
              *reinterpret_cast<float4*>(dst + j) = o4;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mb
