// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue).
//
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, 4-stage mbarrier ring)
//   warp 1      MMA issuer     (one elected lane: tcgen05.mma cta_group::1 kind::f16, 128x256x16, fp32 accum in TMEM)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, double-buffered against the epilogue)
//   warps 4..11 epilogue       (8 warps: TMEM lane quarter = warp % 4, column half = (warp - 4) / 4; tcgen05.ld 32x32b gives
//                               one accumulator row per thread; rows are transposed through a per-warp XOR-swizzled smem
//                               tile so that bias/GELU/residual/pos-embed math and ALL global traffic are row-coalesced)
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
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_STG_BYTES = 32 * 32 * 4;        // per-epilogue-warp staging tile: 32 rows x 32 fp32
constexpr int GEMM_LN_PART = 128;                  // LayerNorm folding: features per row-statistics partial (= columns per epilogue warp)
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
constexpr int GEMM_B_BYTES = GEMM_BN * GEMM_BK * 2;  // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_EPI_WARPS * GEMM_STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

enum : int {
  EPI_STORE16 = 0,  // out16[r,n] = acc + bias[n]
  EPI_GELU16 = 1,   // out16[r,n] = gelu_erf(acc + bias[n])
  EPI_RESID32 = 2,  // out32[r,n] = resid32[r,n] + acc + bias[n]        (out32 may alias resid32)
  EPI_STORE32 = 3,  // out32[r,n] = acc + bias[n] + addend[(m % rows_per_group), n]
  EPI_GELUBWD16 = 4,  // out16[r,n] = acc * gelu'(aux16[r,n])                (fc2 dgrad fused with the GELU backward)
  EPI_ATOMIC32 = 5,   // out32[r,n] += acc  (red.global.add.f32; split-K weight gradients accumulate into .grad)
  EPI_GELU16_SAVE = 6,  // EPI_GELU16 that also stores the pre-activation to aux16 (training forward)
  // LayerNorm folded into the GEMMs around it (inference): the PRODUCER of the residual stream also emits the 16-bit operand
  // xg = x * gamma_next and the per-row sum / sum of squares of x; the CONSUMER multiplies xg by the un-normalised weight and
  // finishes the normalisation per output element:  LN(x) W^T + b = rstd (xg W^T) - rstd mean (W gamma) + (W beta + b).
  EPI_STORE16_LN = 7,   // consumer: out16[r,n] = rstd_r * acc - rstd_r * mean_r * ln_vec[n] + bias[n]            (qkv)
  EPI_GELU16_LN = 8,    // consumer: out16[r,n] = gelu_erf(same)                                                   (fc1)
  EPI_RESID32_LN = 9,   // producer: EPI_RESID32, plus out16b[r,n] = x * ln_vec[n] and, per row and 32-column chunk, the
                        // partial statistics about a pivot                                                  (proj, fc2)
};

struct GemmParams {
  int M, N, K;
  const float* bias;     // [N] or null
  void* out;             // 16-bit or fp32, row stride ld_out elements
  const float* resid;    // EPI_RESID32
  const float* addend;   // EPI_STORE32: optional [rows_per_group, N] table (pos-embed), else null
                         // EPI_GELUBWD16: optional fp32 [N] the column sums of the output are ACCUMULATED into (written through
                         // red.global: the bias gradient of the layer whose pre-activation gradient this GEMM produces)
  void* aux16;           // EPI_GELU16: optional second output, the pre-activation (saved for backward);
                         // EPI_GELUBWD16: input, the saved pre-activation.  Same shape / row stride as out.
  int ld_out;
  int reduce_out;        // EPI_RESID32 with out == resid (x += A W^T + b in place): acc + bias leaves as a TMA reduce-add into out
                         // (tmap_c = fp32 map of out, [32 rows x 32 floats] boxes); the residual never enters the SM
  int k_splits;          // >1: the K loop is split across CTAs (use with EPI_ATOMIC32)
  // LayerNorm folding (EPI_*_LN)
  float* ln_stats;       // producer: partials [N/32][ln_rows][4] = (pivot, sum (x - pivot), sum (x - pivot)^2, -) of every 32-column
                         // chunk (plain stores, no atomics: results are bit-reproducible);  consumer: [rows][2] = (rstd, -mean * rstd)
                         // as written by ln_finalize_kernel
  long ln_rows;          // producer: row pitch of the partials array
  const float* ln_vec;   // consumer: W gamma [N];  producer: gamma of the LayerNorm that follows [N]
  void* out16b;          // producer: second output, x * gamma as op16, same shape / row stride as out

  // output row remap:  r = (m / rows_per_group) * group_stride + row_offset + (m % rows_per_group)
  int rows_per_group, group_stride, row_offset;
};

// erf-form GELU (models/maest.py:500 nn.GELU default; NOT the tanh approximation).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 12 instructions, ONE MUFU: erfc(|x|/sqrt 2) = 2^q(|x|) with q a degree-7 polynomial without constant term (weighted
// least-squares fit of log2 erfc on |x| <= 4 sqrt 2, tools/fit_gelu_poly.py: |gelu error| <= 6e-7 absolute and <= 6e-5 relative
// wherever |gelu| > 1e-4, i.e. well inside the 2.4e-4 rounding of the 16-bit output), then
// 0.5 x (1 + sign(x) erf|.|) = fma(|0.5 x|, 1 - erfc, 0.5 x).  The epilogue of the fc1 GEMM is MUFU- and issue-limited
// (two MUFU per element kept the XU pipe 67 % busy relative to the tile's MMA time), hence one exponential and no reciprocal.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float ax = fminf(fabsf(x), 5.65685425f);
  float q = fmaf(-5.775989393e-07f, ax, 3.963761264e-05f);
  q = fmaf(q, ax, -7.848305395e-04f);
  q = fmaf(q, ax, 8.050128818e-03f);
  q = fmaf(q, ax, -5.326407775e-02f);
  q = fmaf(q, ax, -4.589391351e-01f);
  q = fmaf(q, ax, -1.151135445e+00f);
  const float e = ex2_approx_ftz(q * ax);          // erfc(|x| / sqrt 2)
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), 1.0f - e, hx);
}

// Two elements per call on packed fp32x2 instructions: the same polynomial, 8.5 issue slots per element instead of 13
// (FMNMX x2, seven FFMA2, MUFU x2, FADD2, FSEL x2, FMUL2).  gelu = x Phi(x) with Phi = 1 - erfc/2 (x >= 0) or erfc/2 (x < 0); the
// factor 1/2 rides in the exponent (2^(q |x| - 1)).  The fc1 epilogue is bound by its issue slots and dependency chains on two
// epilogue warps per sub-partition, not by the FMA pipe (r01c_prof_fc1_*: issue 54 %, FMA 35 %).
__device__ __forceinline__ void gelu_erf_fast2(float x0, float x1, float& y0, float& y1) {
  const u64 ax = f2_packf(fminf(fabsf(x0), 5.65685425f), fminf(fabsf(x1), 5.65685425f));
  u64 q = f2_fma(f2_packf(-5.775989393e-07f, -5.775989393e-07f), ax, f2_packf(3.963761264e-05f, 3.963761264e-05f));
  q = f2_fma(q, ax, f2_packf(-7.848305395e-04f, -7.848305395e-04f));
  q = f2_fma(q, ax, f2_packf(8.050128818e-03f, 8.050128818e-03f));
  q = f2_fma(q, ax, f2_packf(-5.326407775e-02f, -5.326407775e-02f));
  q = f2_fma(q, ax, f2_packf(-4.589391351e-01f, -4.589391351e-01f));
  q = f2_fma(q, ax, f2_packf(-1.151135445e+00f, -1.151135445e+00f));
  q = f2_fma(q, ax, f2_packf(-1.0f, -1.0f));                   // log2(erfc(|x| / sqrt 2) / 2)
  float q0, q1;
  f2_unpack(q, q0, q1);
  const float t0 = ex2_approx_ftz(q0), t1 = ex2_approx_ftz(q1);     // erfc / 2 = Phi(-|x|)
  float u0, u1;
  f2_unpack(f2_sub(f2_packf(1.0f, 1.0f), f2_packf(t0, t1)), u0, u1);   // Phi(|x|)
  f2_unpack(f2_mul(f2_packf(x0, x1), f2_packf(x0 >= 0.f ? u0 : t0, x1 >= 0.f ? u1 : t1)), y0, y1);
}

// d/dx [x Phi(x)] = Phi(x) + x phi(x), same erf approximation as the forward
__device__ __forceinline__ float gelu_erf_grad_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx_ftz(-1.4426950408889634f * z * z);   // exp(-x^2 / 2)
  const float erf_v = copysignf(fmaf(-p, e, 1.0f), x);
  return fmaf(0.5f, erf_v, 0.5f) + x * e * 0.3989422804014327f;
}

// Two elements per call, packed fp32x2: Phi(-|x|) from the forward's polynomial (one MUFU), phi(x) from a second one;
// 10.5 issue slots per element instead of ~20 with two MUFU each way (the fc2 input-gradient GEMM was epilogue-bound at 0.65 ms
// against a 0.18 ms mainloop: 13 % of the training step).
__device__ __forceinline__ void gelu_erf_grad_fast2(float x0, float x1, float& g0, float& g1) {
  const u64 x = f2_packf(x0, x1);
  const u64 ax = f2_packf(fminf(fabsf(x0), 5.65685425f), fminf(fabsf(x1), 5.65685425f));
  u64 q = f2_fma(f2_packf(-5.775989393e-07f, -5.775989393e-07f), ax, f2_packf(3.963761264e-05f, 3.963761264e-05f));
  q = f2_fma(q, ax, f2_packf(-7.848305395e-04f, -7.848305395e-04f));
  q = f2_fma(q, ax, f2_packf(8.050128818e-03f, 8.050128818e-03f));
  q = f2_fma(q, ax, f2_packf(-5.326407775e-02f, -5.326407775e-02f));
  q = f2_fma(q, ax, f2_packf(-4.589391351e-01f, -4.589391351e-01f));
  q = f2_fma(q, ax, f2_packf(-1.151135445e+00f, -1.151135445e+00f));
  q = f2_fma(q, ax, f2_packf(-1.0f, -1.0f));                        // log2 Phi(-|x|)
  const u64 z = f2_mul(f2_mul(x, x), f2_packf(-0.7213475204444817f, -0.7213475204444817f));   // -x^2 / 2 in log2 units
  float q0, q1, z0, z1;
  f2_unpack(q, q0, q1);
  f2_unpack(z, z0, z1);
  const float t0 = ex2_approx_ftz(q0), t1 = ex2_approx_ftz(q1);
  const float e0 = ex2_approx_ftz(z0), e1 = ex2_approx_ftz(z1);
  // Phi(x) = 0.5 + copysign(0.5 - t, x) with t = Phi(-|x|) <= 0.5: one LOP3 per element instead of a compare and a select
  float h0, h1;
  f2_unpack(f2_sub(f2_packf(0.5f, 0.5f), f2_packf(t0, t1)), h0, h1);
  h0 = __uint_as_float(__float_as_uint(h0) | (__float_as_uint(x0) & 0x80000000u));
  h1 = __uint_as_float(__float_as_uint(h1) | (__float_as_uint(x1) & 0x80000000u));
  const u64 w = f2_mul(x, f2_packf(0.3989422804014327f, 0.3989422804014327f));
  f2_unpack(f2_add(f2_fma(w, f2_packf(e0, e1), f2_packf(h0, h1)), f2_packf(0.5f, 0.5f)), g0, g1);
}

// 32 contiguous, 32-byte aligned bytes of read-once global data (sm_100: LDG.E.256), not allocated in L1
__device__ __forceinline__ void ldg256_stream(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

// The staging tile is addressed in the shared window explicitly: through the generic pointer ptxas emitted generic LD.E / ST.E
// (address-space resolution on every access) for what are plain ld.shared / st.shared.
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
  return r;
}

// Drain one accumulator half-tile (32 TMEM lanes x up to 128 columns) of the calling epilogue warp: TMEM -> registers ->
// swizzled per-warp smem tile -> row-coalesced fused epilogue.  m0 / n0: first output row / column of this warp's
// sub-tile; `release_tmem()` is invoked once, as soon as the last accumulator column has been read.
template <int DT, int EPI, typename ReleaseFn>
__device__ __forceinline__ void gemm_epilogue_subtile(const GemmParams& p, const CUtensorMap* tmap_c, const CUtensorMap* tmap_c2, uint8_t* stg, uint32_t taddr, int m0, int n0, int lane,
                                                      uint64_t* tfull, uint32_t aphase, ReleaseFn release_tmem) {
  using O = Op16<DT == DT_BF16 ? DT_BF16 : DT_F16>;
  const bool identity_rows = p.rows_per_group == 0x7fffffff;   // set by the host when no remap is requested
  const int sub_row = lane >> 3;  // coalesced phase: 4 rows per instruction, 8 lanes x 16 B per row
  const int c4 = lane & 7;
  int nchunks = (p.N - n0 + 31) / 32;
  nchunks = nchunks < 0 ? 0 : (nchunks > 4 ? 4 : nchunks);
  // Two epilogue organisations.  16-bit outputs (STORE16, GELU16, GELU16_SAVE, LN consumers): math in the TMEM-load layout, packed
  // staging, TMA bulk store (kPacked16 below).  Everything else (fp32 outputs with residual / table reads, GELU backward,
  // atomics): fp32 staging transpose so that all global traffic is row-coalesced.
  // RESID32: the residual tile does not depend on the accumulator, so its global loads are issued one chunk
  // ahead (and, for chunk 0, before waiting for the accumulator) to keep HBM requests in flight.
  // EPI_STORE32 is the only epilogue with an output-row remap (patch tokens -> packed token buffer).  The remap needs an
  // integer division per row, so it is done ONCE per sub-tile (doing it per 32-column chunk made the patch-embed GEMM epilogue
  // 3x slower than its HBM bound); every other epilogue keeps the cheap r = m form and no extra live registers.
  if constexpr (EPI == EPI_RESID32) {
    if (p.reduce_out) {
      // In-place residual update.  The load-add-store form below is bound by the latency of its own residual loads (8 warps x
      // two 4 KB chunks in flight per SM ~ 43 GB/s per SM, just the SM's share of HBM, and nothing left for the stores): proj ran
      // at 0.20 ms against an HBM bound of 0.10 ms.  Here the add happens in L2: every warp stages acc + bias as a
      // [32 rows x 128 B] SWIZZLE_128B tile and fires cp.reduce.async.bulk.tensor (.add.f32); exactly one add per element, so the
      // result is deterministic (and differs from the other form by the order of the two additions only).
      const uint32_t stg_r = smem_u32(stg);
      mbar_wait(tfull, aphase);
      tc_fence_after();
      if (nchunks == 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_tmem();
        return;
      }
      uint32_t va[32], vb[32];
      auto emit = [&](uint32_t (&v)[32], const int cc) {
        const int n = n0 + cc * 32;
        if (lane == 0) tma_store_wait_read<0>();      // staging tile free again (the previous reduce has read it)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias != nullptr) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4 * j));   // warp-uniform (lane = row)
          sts_v4(stg_r + lane * 128 + ((uint32_t(j) ^ uint32_t(lane & 7)) << 4), __float_as_uint(__uint_as_float(v[4 * j]) + bb.x),
                 __float_as_uint(__uint_as_float(v[4 * j + 1]) + bb.y), __float_as_uint(__uint_as_float(v[4 * j + 2]) + bb.z),
                 __float_as_uint(__uint_as_float(v[4 * j + 3]) + bb.w));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(tmap_c, stg, n, m0);      // rows beyond M / columns beyond N are clipped by the tensor map
          tma_store_commit();
        }
      };
      auto release_after_last = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_tmem();
      };
      tmem_ld32(taddr, va);
#pragma unroll 1
      for (int cc = 0; cc < nchunks; cc += 2) {
        tc_wait_ld();
        reg_fence32(va);
        const bool has_b = cc + 1 < nchunks;
        if (has_b) tmem_ld32(taddr + uint32_t((cc + 1) * 32), vb);
        else release_after_last();
        emit(va, cc);
        if (has_b) {
          tc_wait_ld();
          reg_fence32(vb);
          const bool has_a = cc + 2 < nchunks;
          if (has_a) tmem_ld32(taddr + uint32_t((cc + 2) * 32), va);
          else release_after_last();
          emit(vb, cc + 1);
        }
      }
      return;
    }
  }
  constexpr bool kRemap = (EPI == EPI_STORE32);
  long orow_[kRemap ? 8 : 1];
  int prow_[kRemap ? 8 : 1];
  if constexpr (kRemap) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + i * 4 + sub_row;
      prow_[i] = 0;
      if (m >= p.M) orow_[i] = -1;
      else if (identity_rows) orow_[i] = m;
      else { prow_[i] = m % p.rows_per_group; orow_[i] = long(m / p.rows_per_group) * p.group_stride + p.row_offset + prow_[i]; }
    }
  }
  auto row_ok = [&](int i) -> bool { return m0 + i * 4 + sub_row < p.M; };
  auto out_row = [&](int i) -> long {
    if constexpr (kRemap) return orow_[i];
    return long(m0 + i * 4 + sub_row);
  };
  float4 xr[8], xn[8];   // residual / table rows of the current chunk and of the next one (two chunks of loads in flight)
  uint2 ar[8];   // (unused since GELUBWD16 moved to the TMEM-layout path)
  auto load_resid_to = [&](float4 (&xr)[8], int cc) {
    if constexpr (EPI == EPI_RESID32 || EPI == EPI_RESID32_LN) {
      const int col = n0 + cc * 32 + c4 * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xr[i] = (row_ok(i) && cc < nchunks) ? *reinterpret_cast<const float4*>(p.resid + out_row(i) * p.ld_out + col)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if constexpr (EPI == EPI_STORE32) {   // pos-embed table rows (L2-resident): same one-chunk-ahead prefetch
      const int col = n0 + cc * 32 + c4 * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xr[i] = (p.addend != nullptr && row_ok(i) && cc < nchunks) ? __ldg(reinterpret_cast<const float4*>(p.addend + long(prow_[kRemap ? i : 0]) * p.N + col))
                                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if constexpr (EPI == EPI_GELUBWD16) {
      const int col = n0 + cc * 32 + c4 * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ar[i] = (row_ok(i) && cc < nchunks) ? *reinterpret_cast<const uint2*>(reinterpret_cast<const typename O::T*>(p.aux16) + out_row(i) * p.ld_out + col)
                                               : make_uint2(0u, 0u);
      }
    }
  };
  auto load_resid = [&](int cc) { load_resid_to(xr, cc); };
  constexpr bool kLnConsumer = (EPI == EPI_STORE16_LN || EPI == EPI_GELU16_LN);
  constexpr bool kLnProducer = (EPI == EPI_RESID32_LN);
  // 16-bit-output epilogues do their math in the TMEM-load layout (lane = row) and transpose the PACKED 16-bit tile (half the
  // shared-memory traffic of staging fp32).  Measured with experiment builds (tools/gemm_diag.py): with no epilogue at all the
  // pair mainloops run at 1580-1600 TFLOP/s; reading the accumulator out of TMEM is free; for qkv the LSU global stores were
  // 0.048 of the 0.062 ms the epilogue added (hence the TMA store), for fc1 the GELU math is 0.09 ms and the stores 0.045 ms.
  constexpr bool kPacked16 = (EPI == EPI_STORE16 || EPI == EPI_GELU16 || EPI == EPI_GELU16_SAVE || kLnConsumer || EPI == EPI_GELUBWD16);
  // 16-bit output(s) without row remap leave through the TMA store engine.  GELU16_SAVE has two (activation and saved
  // pre-activation): both go through the ONE staging tile of the warp, one after the other.
  constexpr bool kTmaStore16 = (EPI == EPI_STORE16 || EPI == EPI_GELU16 || kLnConsumer || EPI == EPI_GELUBWD16 || EPI == EPI_GELU16_SAVE);
  float lane_r = 0.f, lane_mr = 0.f;   // LN consumer: rstd and -mean * rstd of row m0 + lane
  if constexpr (kLnConsumer) {
    if (m0 + lane < p.M) {
      const float2 st = *reinterpret_cast<const float2*>(p.ln_stats + 2 * long(m0 + lane));
      lane_r = st.x;
      lane_mr = st.y;
    }
  }
  if constexpr (EPI != EPI_GELUBWD16) load_resid(0);      // (GELUBWD16 fetches its pre-activation rows in the TMEM-load layout, below)
#ifdef GEMM_PRE_ALL
  uint4 u_all[EPI == EPI_GELUBWD16 ? 16 : 1];
  if constexpr (EPI == EPI_GELUBWD16) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int n = n0 + cc * 32;
      const bool ok = (m0 + lane < p.M) && (n + 32 <= p.N);
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const typename O::T*>(p.aux16) + long(m0 + lane) * p.ld_out + n);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (ok) ldg256_stream(src + 2 * h, u_all[4 * cc + 2 * h], u_all[4 * cc + 2 * h + 1]);
        else u_all[4 * cc + 2 * h] = u_all[4 * cc + 2 * h + 1] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
#endif
  const uint32_t stg_s = smem_u32(stg);
  // bias of the next chunk is fetched while the current one is processed (its L2 / L1 latency sat on the first FADD of every chunk)
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias != nullptr && nchunks > 0) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c4 * 4));
  mbar_wait(tfull, aphase);
  tc_fence_after();
#ifdef GEMM_DIAG_NOEPI   // timing diagnostic: mainloop only (outputs are not written)
  nchunks = 0;
#endif
  if (nchunks == 0) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) release_tmem();     // one arrival per warp (the barrier counts warps, not threads)
  }
  if constexpr (kPacked16) {
    // 16-bit outputs: the accumulator chunks are read out of TMEM one chunk AHEAD (two register tiles, the tcgen05.ld of chunk
    // cc + 1 is in flight while chunk cc is processed), and the TMEM stage goes back to the MMA warp as soon as the last
    // chunk's load has completed.
    uint32_t va[32], vb[32];
    uint32_t pre_keep[EPI == EPI_GELU16_SAVE ? 16 : 1];   // GELU16_SAVE: packed pre-activations of the even chunk, until its pair is complete
    // GELUBWD16: this thread's row of the saved pre-activation, 32 columns (64 contiguous bytes) per chunk, fetched one chunk
    // ahead like the accumulator
    uint4 ua[EPI == EPI_GELUBWD16 ? 4 : 1], ub[EPI == EPI_GELUBWD16 ? 4 : 1];
    auto load_pre = [&](uint4 (&u)[EPI == EPI_GELUBWD16 ? 4 : 1], const int cc) {
      if constexpr (EPI == EPI_GELUBWD16) {
        const int n = n0 + cc * 32;
        const bool ok = (m0 + lane < p.M) && (n + 32 <= p.N);
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const typename O::T*>(p.aux16) + long(m0 + lane) * p.ld_out + n);
#pragma unroll
#ifdef GEMM_DIAG_NOPRE
        for (int q4 = 0; q4 < 4; ++q4) u[q4] = make_uint4(uint32_t(n), 0u, uint32_t(cc), 0u);
#elif defined(GEMM_PRE_LDG128)
        for (int q4 = 0; q4 < 4; ++q4) u[q4] = ok ? src[q4] : make_uint4(0u, 0u, 0u, 0u);
#else
        // lane = row, so a warp-wide load touches 32 different rows: with 16-byte loads every request used half of each 32-byte
        // sector and relied on L1 (a few KB beside 224 KB of shared memory) for the other half -- measured 0.15 ms of a 0.38 ms
        // GEMM.  One 256-bit load per lane = one whole sector per lane, streamed past L1.
        for (int h = 0; h < 2; ++h) {
          if (ok) ldg256_stream(src + 2 * h, u[2 * h], u[2 * h + 1]);
          else u[2 * h] = u[2 * h + 1] = make_uint4(0u, 0u, 0u, 0u);
        }
#endif
      }
    };
    auto release_after_last_load = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_tmem();   // one arrival per warp: 8 (16 for a CTA pair, half of them remote) instead of 256 (512)
    };
    auto process = [&](uint32_t (&v)[32], const uint4* u, const int cc) {
      const int n = n0 + cc * 32;

      // lane = row: bias / fold vectors are warp-uniform (broadcast) loads; results are packed to 16 bits, written to a
      // [32 rows x 64 B] tile (16-byte slots XOR-swizzled by (row >> 1) & 3: conflict-free both ways) and read back so that
      // 4 lanes cover one row's 64 contiguous bytes (a warp instruction stores 8 full row segments).
      uint32_t o16[16], pre16[16];
      float cs[EPI == EPI_GELUBWD16 ? 32 : 1];     // GELUBWD16: this row's 32 output values (fp32) for the column sums below
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (EPI != EPI_GELUBWD16) {     // (an input gradient has no bias term)
          if (p.bias != nullptr) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4 * j));
        }
        float4 a = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        if constexpr (kLnConsumer) {   // finish the LayerNorm: rstd * acc - rstd * mean * (W gamma) + (W beta + b)
          const float4 gg = __ldg(reinterpret_cast<const float4*>(p.ln_vec + n + 4 * j));
          a.x = fmaf(a.x, lane_r, fmaf(lane_mr, gg.x, bb.x)); a.y = fmaf(a.y, lane_r, fmaf(lane_mr, gg.y, bb.y));
          a.z = fmaf(a.z, lane_r, fmaf(lane_mr, gg.z, bb.z)); a.w = fmaf(a.w, lane_r, fmaf(lane_mr, gg.w, bb.w));
        } else if constexpr (EPI != EPI_GELUBWD16) {
          a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
        }
        if constexpr (EPI == EPI_GELU16_SAVE) { pre16[2 * j] = O::pack(a.x, a.y); pre16[2 * j + 1] = O::pack(a.z, a.w); }
        if constexpr (EPI == EPI_GELU16 || EPI == EPI_GELU16_SAVE || EPI == EPI_GELU16_LN) {
          gelu_erf_fast2(a.x, a.y, a.x, a.y);
          gelu_erf_fast2(a.z, a.w, a.z, a.w);
        }
        if constexpr (EPI == EPI_GELUBWD16) {     // d(pre-activation) = d(activation) * gelu'(pre-activation)
          const uint32_t w0 = (j & 1) ? u[j >> 1].z : u[j >> 1].x, w1 = (j & 1) ? u[j >> 1].w : u[j >> 1].y;
          const float2 u01 = O::unpack(w0), u23 = O::unpack(w1);
          float g0, g1, g2, g3;
#ifdef GEMM_DIAG_NOGRAD
          g0 = u01.x; g1 = u01.y; g2 = u23.x; g3 = u23.y;
#else
          gelu_erf_grad_fast2(u01.x, u01.y, g0, g1);
          gelu_erf_grad_fast2(u23.x, u23.y, g2, g3);
#endif
          a.x *= g0; a.y *= g1; a.z *= g2; a.w *= g3;
          cs[4 * j] = a.x; cs[4 * j + 1] = a.y; cs[4 * j + 2] = a.z; cs[4 * j + 3] = a.w;
        }
        o16[2 * j] = O::pack(a.x, a.y);
        o16[2 * j + 1] = O::pack(a.z, a.w);
      }
      if constexpr (EPI == EPI_GELUBWD16) {
        // bias gradient = column sums over the rows: thread = row here, so a transposing butterfly over the warp -- at distance
        // 16, 8, .. 1 every lane keeps the half of its values whose column bit matches its lane bit and receives the partner's
        // sums for that half: 31 shuffles + 31 adds, after which lane L holds the 32-row sum of column n + L.  Rows beyond M
        // contribute exact zeros (their accumulator rows are zero).
#ifdef GEMM_DIAG_NOCS
        if (false) {
#else
        if (p.addend != nullptr) {
#endif
#pragma unroll
          for (int w = 16; w >= 1; w >>= 1) {
            const bool up = (lane & w) != 0;
#pragma unroll
            for (int k = 0; k < w; ++k) {
              const float keep = up ? cs[k + w] : cs[k];
              const float send = up ? cs[k] : cs[k + w];
              cs[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
            }
          }
          if (n + lane < p.N) atomicAdd(const_cast<float*>(p.addend) + n + lane, cs[0]);
        }
      }
      if constexpr (kTmaStore16) {
        // The tile leaves through the TMA store engine: two 32-column chunks form a [32 rows x 128 B] SWIZZLE_128B tile in the
        // staging buffer, one elected lane issues cp.async.bulk.tensor (full 128-byte lines, rows beyond M clipped by the
        // tensor map).  Measured before: of the 0.062 ms the qkv epilogue added to a 0.242 ms mainloop, 0.048 ms were the
        // LSU global stores themselves (0.043 of 0.132 ms for fc1).
        const int half_sel = cc & 1;
        if (half_sel == 0) {                       // staging tile free again? (reads of the previous TMA store done)
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          sts_v4(stg_s + lane * 128 + ((uint32_t(half_sel * 4 + q4) ^ uint32_t(lane & 7)) << 4), o16[4 * q4], o16[4 * q4 + 1], o16[4 * q4 + 2], o16[4 * q4 + 3]);
        if constexpr (EPI == EPI_GELU16_SAVE) {
          if (half_sel == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) pre_keep[q] = pre16[q];
          }
        }
        if (half_sel == 1 || cc == nchunks - 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tmap_c, stg, n - 32 * half_sel, m0);   // columns beyond N / rows beyond M are clipped
            tma_store_commit();
          }
          if constexpr (EPI == EPI_GELU16_SAVE) {
            // second output through the same staging tile: wait until the store above has read it, refill with the pre-activations
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            if (half_sel == 1) {
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                sts_v4(stg_s + lane * 128 + ((uint32_t(q4) ^ uint32_t(lane & 7)) << 4), pre_keep[4 * q4], pre_keep[4 * q4 + 1], pre_keep[4 * q4 + 2], pre_keep[4 * q4 + 3]);
            }
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              sts_v4(stg_s + lane * 128 + ((uint32_t(half_sel * 4 + q4) ^ uint32_t(lane & 7)) << 4), pre16[4 * q4], pre16[4 * q4 + 1], pre16[4 * q4 + 2], pre16[4 * q4 + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(tmap_c2, stg, n - 32 * half_sel, m0);
              tma_store_commit();
            }
          }
        }
        return;
      }
      __syncwarp();            // previous chunk's staging reads are complete
      const uint32_t wsw = uint32_t((lane >> 1) & 3);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        sts_v4(stg_s + lane * 64 + ((uint32_t(q4) ^ wsw) << 4), o16[4 * q4], o16[4 * q4 + 1], o16[4 * q4 + 2], o16[4 * q4 + 3]);
        if constexpr (EPI == EPI_GELU16_SAVE)
          sts_v4(stg_s + 2048 + lane * 64 + ((uint32_t(q4) ^ wsw) << 4), pre16[4 * q4], pre16[4 * q4 + 1], pre16[4 * q4 + 2], pre16[4 * q4 + 3]);
      }
      __syncwarp();
      const int slot = lane & 3;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rl = 8 * k + (lane >> 2);
        const int m = m0 + rl;
        const uint32_t off = uint32_t(rl * 64) + ((uint32_t(slot) ^ uint32_t((rl >> 1) & 3)) << 4);
        const float4 w = lds_f4(stg_s + off);
        if (m < p.M)
          st_global_v4(reinterpret_cast<typename O::T*>(p.out) + long(m) * p.ld_out + n + 8 * slot, __float_as_uint(w.x), __float_as_uint(w.y),
                       __float_as_uint(w.z), __float_as_uint(w.w));
        if constexpr (EPI == EPI_GELU16_SAVE) {
          const float4 wp = lds_f4(stg_s + 2048 + off);
          if (m < p.M)
            st_global_v4(reinterpret_cast<typename O::T*>(p.aux16) + long(m) * p.ld_out + n + 8 * slot, __float_as_uint(wp.x), __float_as_uint(wp.y),
                         __float_as_uint(wp.z), __float_as_uint(wp.w));
        }
      }
    };
#ifdef GEMM_PRE_ALL
    if constexpr (EPI == EPI_GELUBWD16) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (cc < nchunks) {
          tmem_ld32(taddr + uint32_t(cc * 32), va);
          tc_wait_ld();
          reg_fence32(va);
          if (cc == nchunks - 1) release_after_last_load();
          process(va, &u_all[4 * cc], cc);
        }
      }
      return;
    }
#endif
    if (nchunks > 0) { tmem_ld32(taddr, va); load_pre(ua, 0); }
#pragma unroll 1
    for (int cc = 0; cc < nchunks; cc += 2) {
      tc_wait_ld();
      reg_fence32(va);
      const bool has_b = cc + 1 < nchunks;
      if (has_b) { tmem_ld32(taddr + uint32_t((cc + 1) * 32), vb); load_pre(ub, cc + 1); }
      else release_after_last_load();
      process(va, ua, cc);
      if (has_b) {
        tc_wait_ld();
        reg_fence32(vb);
        const bool has_a = cc + 2 < nchunks;
        if (has_a) { tmem_ld32(taddr + uint32_t((cc + 2) * 32), va); load_pre(ua, cc + 2); }
        else release_after_last_load();
        process(vb, ub, cc + 1);
      }
    }
    return;
  }
  float ln_pv[kLnProducer ? 8 : 1], ln_s1[kLnProducer ? 8 : 1], ln_s2[kLnProducer ? 8 : 1];
#pragma unroll 1
  for (int cc = 0; cc < nchunks; ++cc) {
    const int n = n0 + cc * 32;
    uint32_t v[32];
    tmem_ld32(taddr + uint32_t(cc * 32), v);
    tc_wait_ld();
    if (cc == nchunks - 1) {   // accumulator fully drained by this warp: hand the TMEM stage back early
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_tmem();   // one arrival per warp: 8 (16 for a CTA pair, half of them remote) instead of 256 (512)
    }
    // the NEXT chunk's residual / table rows are requested before this chunk is touched: two chunks of global loads in flight per
    // thread (the HBM-bound proj GEMM ran at 4.2 TB/s with one: the loads issued after chunk cc's arithmetic were not back
    // when chunk cc + 1 needed them)
    if constexpr (EPI == EPI_RESID32 || EPI == EPI_STORE32 || kLnProducer) load_resid_to(xn, cc + 1);
    __syncwarp();              // previous chunk's staging reads are complete
#pragma unroll
    for (int j = 0; j < 8; ++j) sts_v4(stg_s + lane * 128 + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int col = n + c4 * 4;
    const float4 b4_cur = b4;
    if (p.bias != nullptr && cc + 1 < nchunks) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 32));
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (kLnProducer) g4 = __ldg(reinterpret_cast<const float4*>(p.ln_vec + col));
    float4 acc4[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = i * 4 + sub_row;
      float4 a = lds_f4(stg_s + rl * 128 + ((c4 ^ (rl & 7)) << 4));
      a.x += b4_cur.x; a.y += b4_cur.y; a.z += b4_cur.z; a.w += b4_cur.w;
      if constexpr (EPI == EPI_RESID32 || EPI == EPI_STORE32 || kLnProducer) { a.x += xr[i].x; a.y += xr[i].y; a.z += xr[i].z; a.w += xr[i].w; }
      acc4[i] = a;
    }
    if constexpr (EPI == EPI_RESID32 || EPI == EPI_STORE32 || kLnProducer) {
#pragma unroll
      for (int i = 0; i < 8; ++i) xr[i] = xn[i];
    }
    if constexpr (kLnProducer) {
      // Row statistics of the new residual stream over THIS WARP'S 128 columns (GEMM_LN_PART), about a pivot (the first element of
      // the row in the warp's first chunk, broadcast inside the 8 lanes that share a row) so that a large common offset of the
      // row does not cancel.  Every lane accumulates (sum d, sum d^2), d = x - pivot, over its 4 columns of all four chunks in
      // registers; ONE 3-step butterfly per row after the last chunk (it was one per 32-column chunk: 224 shuffles per sub-tile
      // and thread instead of 56 -- the folded proj epilogue took 0.40 ms against 0.20 ms for the plain one).
      // ln_finalize_kernel merges the N / 128 partials of a row.  Partials are laid out [part][row]: coalesced merge reads.
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = acc4[i];
        if (cc == 0) {
          ln_pv[i] = __shfl_sync(0xffffffffu, a.x, lane & ~7);
          ln_s1[i] = 0.f;
          ln_s2[i] = 0.f;
        }
        const float pv = ln_pv[i];
        const float dx = a.x - pv, dy = a.y - pv, dz = a.z - pv, dw = a.w - pv;
        ln_s1[i] += (dx + dy) + (dz + dw);
        ln_s2[i] = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, ln_s2[i]))));
      }
      if (cc == nchunks - 1) {
        const int part = n0 / GEMM_LN_PART;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float s1 = ln_s1[i], s2 = ln_s2[i];
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          if (row_ok(i) && c4 == 0)
            *reinterpret_cast<float4*>(p.ln_stats + (long(part) * p.ln_rows + out_row(i)) * 4) = make_float4(ln_pv[i], s1, s2, float(32 * nchunks));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 a = acc4[i];
      if (!row_ok(i)) continue;
      const long r = out_row(i);
      if constexpr (EPI == EPI_GELUBWD16) {
        const uint2 pre = *reinterpret_cast<const uint2*>(reinterpret_cast<const typename O::T*>(p.aux16) + r * p.ld_out + col);
        const float2 u01 = O::unpack(pre.x), u23 = O::unpack(pre.y);
        a.x *= gelu_erf_grad_fast(u01.x); a.y *= gelu_erf_grad_fast(u01.y);
        a.z *= gelu_erf_grad_fast(u23.x); a.w *= gelu_erf_grad_fast(u23.y);
        uint2 o;
        o.x = O::pack(a.x, a.y);
        o.y = O::pack(a.z, a.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<typename O::T*>(p.out) + r * p.ld_out + col) = o;
      } else if constexpr (EPI == EPI_ATOMIC32) {
        float* dst = reinterpret_cast<float*>(p.out) + r * p.ld_out + col;
        atomicAdd(dst, a.x); atomicAdd(dst + 1, a.y); atomicAdd(dst + 2, a.z); atomicAdd(dst + 3, a.w);
      } else if constexpr (EPI == EPI_RESID32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + r * p.ld_out + col) = a;
      } else if constexpr (kLnProducer) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + r * p.ld_out + col) = a;
        uint2 o;
        o.x = O::pack(a.x * g4.x, a.y * g4.y);
        o.y = O::pack(a.z * g4.z, a.w * g4.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<typename O::T*>(p.out16b) + r * p.ld_out + col) = o;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + r * p.ld_out + col) = a;
      }
    }
  }
}

// A_MN / B_MN: the operand is stored with its M (resp. N) index contiguous and the reduction index as the row
// ("MN-major": natural layout of activations / gradients / weights in the backward GEMMs, so no transposes are
// materialised).  Such a tile is fetched as [64 k-rows x 64 elements] TMA boxes, one per 64-wide M/N group.
template <int DT, int EPI, bool A_MN = false, bool B_MN = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + GEMM_STAGES * GEMM_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + GEMM_EPI_WARPS * GEMM_STG_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + GEMM_STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * GEMM_STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * GEMM_STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (p.N + GEMM_BN - 1) / GEMM_BN;
  const int splits = p.k_splits > 1 ? p.k_splits : 1;
  const int num_tiles = num_m * num_n * splits;       // work items: (output tile, K split)
  const int num_kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  auto kb_range = [&](int t, int& kb0, int& kb1) {
    const int sp = t % splits;
    kb0 = int(long(num_kb_total) * sp / splits);
    kb1 = int(long(num_kb_total) * (sp + 1) / splits);
  };

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
      mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);
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
        const int tile = t / splits;
        const int m0 = (tile / num_n) * GEMM_BM;
        const int n0 = (tile % num_n) * GEMM_BN;
        int kb0, kb1;
        kb_range(t, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * GEMM_STAGE_BYTES;
          uint8_t* sb = sa + GEMM_A_BYTES;
          mbar_expect_tx(&full_bar[stage], GEMM_STAGE_BYTES);
          if constexpr (A_MN) {
#pragma unroll
            for (int g = 0; g < GEMM_BM / 64; ++g) tma_load_2d(sa + g * 8192, &tmap_a, &full_bar[stage], m0 + g * 64, kb * GEMM_BK);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * GEMM_BK, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int g = 0; g < GEMM_BN / 64; ++g) tma_load_2d(sb + g * 8192, &tmap_b, &full_bar[stage], n0 + g * 64, kb * GEMM_BK);
          } else {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * GEMM_BK, n0);
          }
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(DT, GEMM_BM, GEMM_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * GEMM_BN);
        int kb0, kb1;
        kb_range(t, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * GEMM_STAGE_BYTES);
          const uint32_t sb = sa + GEMM_A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // K-major: advance 16 K-elements = 32 bytes inside the 128-byte swizzle row.
            // MN-major: advance 16 k-rows = 2048 bytes; LBO = 8192 (next 64-wide M/N group), SBO = 1024 (next 8 k-rows).
            const uint64_t adesc = A_MN ? make_sdesc(sa + uint32_t(k * 2048), 8192, 1024) : make_sdesc(sa, 16, 1024) + uint64_t(2 * k);
            const uint64_t bdesc = B_MN ? make_sdesc(sb + uint32_t(k * 2048), 8192, 1024) : make_sdesc(sb, 16, 1024) + uint64_t(2 * k);
            mma_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    const int q = warp & 3;        // TMEM lane quarter this warp may access
    const int half = e >> 2;       // which 128 accumulator columns this warp drains
    uint8_t* stg = stg_base + e * GEMM_STG_BYTES;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tile = t / splits;
      const int m0 = (tile / num_n) * GEMM_BM + q * 32;
      const int n0 = (tile % num_n) * GEMM_BN + half * 128;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * GEMM_BN + half * 128);
      gemm_epilogue_subtile<DT, EPI>(p, &tmap_c, &tmap_c2, stg, taddr, m0, n0, lane, &tfull_bar[as], aphase, [&]() { mbar_arrive(&tempty_bar[as]); });
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();   // bulk stores of this warp (TMA-store epilogues) have landed before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mb
