// Flash-style attention forward for sm_100a (tcgen05 + TMEM + TMA), d_head = 64, no N x N matrix in HBM.
//
// Replaces models/maest.py:362-375 (reshape/permute, q@k^T*scale, softmax, attn@v, transpose) — the
// reference materialises [B,12,N,N] several times per block.
//
// Input is the packed qkv activation [B*N, 3*768] (16-bit) exactly as the qkv GEMM writes it; q/k/v
// tiles of head h are fetched by TMA as 64-column boxes at column offsets {0,768,1536} + 64 h, so no
// head-major re-layout pass exists.  Output o[B*N, 768] (16-bit), column block 64 h.
//
// CTA = one 128-query tile of one (clip, head); 2 CTAs co-reside per SM (256 TMEM columns, <=112 KB smem
// each) so one CTA's softmax overlaps the other's MMAs.
//   warps 0..3  softmax: thread = one query row (tcgen05.ld 32x32b: TMEM lane == row), online softmax in
//               fp32 with exp2 and lazy rescaling of the O accumulator, P written as 16-bit
//   warp 4      TMA producer (Q once, K/V double-buffered) + TMEM allocator
//   warp 5      MMA issuer: S = Q K^T (128x128x64), O += P V (128x64x128); V is consumed in its natural
//               [key][d] layout as an MN-major B operand
// P_IN_TMEM: P is stored back to TMEM (tcgen05.st) and fed as the A operand from TMEM (no smem round trip).
//
// Measured dead ends on B200 (config 3, per launch; baseline 0.889 ms): skipping the padded 32-key chunks of the last KV
// tile 1.07 ms (per-chunk predicates defeat the scheduling of the hot loop); letting warps whose query rows are all
// padding idle 0.889 ms (no change); polynomial exp2 for 25 % / 50 % of the scores 0.901 / 0.991 ms; two threads per
// query row 1.33 ms.  The kernel is bound by the per-tile S read from TMEM plus the MUFU exponentials, not by issue slots.
#pragma once
#include "common.cuh"

namespace mb {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: any [128 x 64] 16-bit tile
constexpr int ATT_KV_STAGES = 2;

template <bool P_IN_TMEM>
__host__ __device__ constexpr int att_smem_bytes() {
  return ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES + (P_IN_TMEM ? 0 : 2)) + 128;
}

struct AttnParams {
  int B, N;         // clips, tokens per clip (rows of clip b are [b*N, (b+1)*N))
  int H;            // heads
  int ld_qkv;       // 3*H*64
  int ld_out;       // H*64
  void* out;        // [B*N, H*64] 16-bit
  float scale_log2; // d^-0.5 * log2(e)
  float* lse;       // optional [B, H, N] fp32: m + log2(l) per query row in the scaled log2 domain (saved for backward)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest range reduction with the 1.5*2^23 trick, degree-4 polynomial on
// [-0.5, 0.5] (max relative error 8.4e-6, far below the 2.4e-4 rounding of the 16-bit P it feeds), exponent re-inserted
// with an integer add.  At d_head = 64 the softmax is MUFU-bound (one exp2 per score vs. 2 x 64 MACs), so a quarter of the
// exponentials can be moved off the MUFU unit (ATT_POLY_EVERY).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float r = x - (t - 12582912.0f);
  float q = fmaf(0.0092549640f, r, 0.0558560291f);
  q = fmaf(q, r, 0.2403125134f);
  q = fmaf(q, r, 0.6931306379f);
  q = fmaf(q, r, 0.9999977904f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}
#ifndef ATT_POLY_EVERY
#define ATT_POLY_EVERY 0   // N > 0: every N-th exponential uses ex2_poly.  Measured on B200 (config 3): 0 -> 0.888 ms, 4 -> 0.901 ms, 2 -> 0.991 ms per launch: the kernel is latency-, not MUFU-throughput-bound, so it stays off.
#endif

template <int DT, bool P_IN_TMEM>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using O16 = Op16<DT>;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;                              // [stages]
  uint8_t* sV = smem + ATT_TILE_BYTES * (1 + ATT_KV_STAGES);        // [stages]
  uint8_t* sP = smem + ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES);    // SS mode only: two [128 x 64] halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + att_smem_bytes<P_IN_TMEM>() - 128);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]
  uint64_t* s_full = bars + 9;        // S_j written by the tensor pipe
  uint64_t* s_free = bars + 10;       // S_j copied to registers by all 128 softmax threads (count 128)
  uint64_t* p_full = bars + 11;       // P_j written (count 128)
  uint64_t* o_done = bars + 12;       // O += P_j V_j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row_base = b * p.N;
  const int nkv = (p.N + ATT_BKV - 1) / ATT_BKV;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("attention: dynamic smem base not 1024-aligned\n");
    __trap();
  }
  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    if (lane == 0) tma_prefetch_desc(&tmap_qkv);
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;          // 128 fp32 columns
  const uint32_t tO = tmem_base + 128;    // 64 fp32 columns
  const uint32_t tP = tmem_base + 192;    // 64 columns of packed 16-bit pairs (TS mode)

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(sQ, &tmap_qkv, q_full, h * ATT_D, row_base + q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        const int r = row_base + j * ATT_BKV;
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_expect_tx(&k_full[stage], ATT_TILE_BYTES);
        tma_load_2d(sK + stage * ATT_TILE_BYTES, &tmap_qkv, &k_full[stage], p.H * ATT_D + h * ATT_D, r);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_expect_tx(&v_full[stage], ATT_TILE_BYTES);
        tma_load_2d(sV + stage * ATT_TILE_BYTES, &tmap_qkv, &v_full[stage], 2 * p.H * ATT_D + h * ATT_D, r);
        if (++stage == ATT_KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc(DT, 128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(DT, 128, 64, 0, 1);  // B = V, MN-major
      const uint64_t qdesc = make_sdesc(smem_u32(sQ), 16, 1024);
      auto issue_qk = [&](int stage) {
        const uint64_t kdesc = make_sdesc(smem_u32(sK + stage * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          mma_ss(tS, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_qk, k ? 1u : 0u);
        tc_commit(&k_empty[stage]);   // K stage reusable once S = Q K^T has retired
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == ATT_KV_STAGES) { nstage = 0; nphase ^= 1; }
        if (j + 1 < nkv) {
          // S_j lives in the softmax threads' registers now: overwrite it with S_{j+1} while they exponentiate
          mbar_wait(s_free, j & 1);
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          issue_qk(nstage);
        }
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[stage], phase);
        tc_fence_after();
        // O (+)= P_j V_j : 8 K-steps of 16 keys
        const uint32_t vbase = smem_u32(sV + stage * ATT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k) {
          const uint64_t vdesc = make_sdesc(vbase + uint32_t(k * 16 * 128), 8192, 1024);
          if constexpr (P_IN_TMEM) {
            mma_ts(tO, tP + uint32_t(8 * k), vdesc, idesc_pv, (j | k) ? 1u : 0u);
          } else {
            const uint64_t pdesc = make_sdesc(smem_u32(sP + (k >> 2) * ATT_TILE_BYTES), 16, 1024) + uint64_t(2 * (k & 3));
            mma_ss(tO, pdesc, vdesc, idesc_pv, (j | k) ? 1u : 0u);
          }
        }
        tc_commit(&v_empty[stage]);
        tc_commit(o_done);
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ---------------- softmax warps: thread <-> query row (TMEM lane) ----------------
    const int row = warp * 32 + lane;
    const uint32_t lane_off = uint32_t(warp * 32) << 16;
    float m_run = -INFINITY;  // running max of s * scale_log2
    float l_run = 0.f;
    const float sc = p.scale_log2;
    for (int j = 0; j < nkv; ++j) {
      const int kv0 = j * ATT_BKV;
      const int valid = p.N - kv0;  // keys [0, valid) of this tile are real (>=128 when not the last tile)
      const bool full_tile = valid >= ATT_BKV;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // the whole score row goes to registers in one shot; the TMEM columns are handed back immediately
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tS + lane_off + uint32_t(c * 32), *reinterpret_cast<uint32_t(*)[32]>(s + 32 * c));
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(s_free);
      float mt = -INFINITY;
      if (!full_tile) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) s[i] = 0xff800000u;   // -inf: masked keys contribute exp2(-inf) = 0
      }
      {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;   // four independent chains
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(s[i]));
          m1 = fmaxf(m1, __uint_as_float(s[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(s[i + 2]));
          m3 = fmaxf(m3, __uint_as_float(s[i + 3]));
        }
        mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      const float mt_sc = mt * sc;
      bool waited_pv = (j == 0);   // PV_{j-1} must retire before O is corrected or P is overwritten
      if (j > 0) {
        const bool need = mt_sc > m_run + 8.0f;   // lazy rescale: p stays <= 2^8 against a stale max
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(o_done, (j - 1) & 1);
          tc_fence_after();
          waited_pv = true;
          float f = 1.0f;
          if (need) {
            f = ex2_approx(m_run - mt_sc);
            m_run = mt_sc;
            l_run *= f;
          }
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(tO + lane_off + uint32_t(c * 32), v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st32(tO + lane_off + uint32_t(c * 32), v);
          }
          tc_wait_st();
        }
      } else {
        m_run = mt_sc;
      }
      // p = exp2(s*sc - m), row sum, P -> 16-bit.  The first 32 exponentials are computed BEFORE waiting for
      // PV_{j-1}, so the tensor-pipe latency of the previous tile hides behind MUFU work.
      const float neg_m = -m_run;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a0 = fmaf(__uint_as_float(s[c * 32 + i]), sc, neg_m);
          const float a1 = fmaf(__uint_as_float(s[c * 32 + i + 1]), sc, neg_m);
          const float p0 = ex2_approx(a0);
          #if ATT_POLY_EVERY > 0
          const float p1 = (((i + 1) % ATT_POLY_EVERY) == ATT_POLY_EVERY - 1) ? ex2_poly(a1) : ex2_approx(a1);
#else
          const float p1 = ex2_approx(a1);
#endif
          l0 += p0;
          l1 += p1;
          pk[i >> 1] = O16::pack(p0, p1);
        }
        if (c == 0 && !waited_pv) {
          mbar_wait(o_done, (j - 1) & 1);
          tc_fence_after();
        }
        if constexpr (P_IN_TMEM) {
          tmem_st16(tP + lane_off + uint32_t(c * 16), pk);
        } else {
          // K-major SWIZZLE_128B tile [128 rows x 64 keys]: 16-byte chunk index XOR (row & 7)
          uint8_t* base = sP + (c >> 1) * ATT_TILE_BYTES + row * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = ((c & 1) * 4 + q4) ^ (row & 7);
            *reinterpret_cast<uint4*>(base + chunk * 16) =
                make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
        }
      }
      l_run += l0 + l1;
      if constexpr (P_IN_TMEM) tc_wait_st();
      else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---------------- epilogue: O / l -> 16-bit ----------------
    mbar_wait(o_done, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int qrow = q0 + row;
    if (p.lse != nullptr && qrow < p.N) p.lse[(long(b) * p.H + h) * p.N + qrow] = m_run + log2f(l_run);
    typename O16::T* dst = reinterpret_cast<typename O16::T*>(p.out) + long(row_base + qrow) * p.ld_out + h * ATT_D;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tO + lane_off + uint32_t(c * 32), v);
      tc_wait_ld();
      if (qrow < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          st_global_v4(dst + c * 32 + i,
                       O16::pack(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l),
                       O16::pack(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l),
                       O16::pack(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l),
                       O16::pack(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Variant 2 ("split rows", experimental — measured SLOWER on B200: 1.33 ms vs 0.89 ms per launch at config 3, the per-tile
// 256-thread barrier costs more than the extra warps hide; kept for reference, not used by default): same pipeline as attention_fwd_kernel<DT, true>, but every query row is shared by TWO softmax
// threads (64 score columns each; warps w and w+4 own the same TMEM lane quarter), i.e. 8 softmax warps per CTA and
// 16 per SM.  Twice the warps hide the TMEM / mbarrier / MUFU latencies that bound the 4-warp version; the only
// cross-thread traffic is one float (the partial row max) per row per KV tile through shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int ATT2_THREADS = 320;
constexpr int ATT2_SMEM_BYTES = ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES) + 4096 + 128;

template <int DT>
__global__ void __launch_bounds__(ATT2_THREADS, 2)
attention_fwd_split_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using O16 = Op16<DT>;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;
  uint8_t* sV = smem + ATT_TILE_BYTES * (1 + ATT_KV_STAGES);
  float* xmax = reinterpret_cast<float*>(smem + ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES));   // [2 parity][2 half][128]
  float* xsum = xmax + 512;                                                                 // [2 half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT2_SMEM_BYTES - 128);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;   // count 256
  uint64_t* p_full = bars + 11;   // count 256
  uint64_t* o_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row_base = b * p.N;
  const int nkv = (p.N + ATT_BKV - 1) / ATT_BKV;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("attention: dynamic smem base not 1024-aligned\n");
    __trap();
  }
  if (warp == 9 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(p_full, 256);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    if (lane == 0) tma_prefetch_desc(&tmap_qkv);
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 192;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(sQ, &tmap_qkv, q_full, h * ATT_D, row_base + q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        const int r = row_base + j * ATT_BKV;
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_expect_tx(&k_full[stage], ATT_TILE_BYTES);
        tma_load_2d(sK + stage * ATT_TILE_BYTES, &tmap_qkv, &k_full[stage], p.H * ATT_D + h * ATT_D, r);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_expect_tx(&v_full[stage], ATT_TILE_BYTES);
        tma_load_2d(sV + stage * ATT_TILE_BYTES, &tmap_qkv, &v_full[stage], 2 * p.H * ATT_D + h * ATT_D, r);
        if (++stage == ATT_KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc(DT, 128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(DT, 128, 64, 0, 1);
      const uint64_t qdesc = make_sdesc(smem_u32(sQ), 16, 1024);
      auto issue_qk = [&](int stage) {
        const uint64_t kdesc = make_sdesc(smem_u32(sK + stage * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          mma_ss(tS, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_qk, k ? 1u : 0u);
        tc_commit(&k_empty[stage]);
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == ATT_KV_STAGES) { nstage = 0; nphase ^= 1; }
        if (j + 1 < nkv) {
          mbar_wait(s_free, j & 1);
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          issue_qk(nstage);
        }
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[stage], phase);
        tc_fence_after();
        const uint32_t vbase = smem_u32(sV + stage * ATT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k)
          mma_ts(tO, tP + uint32_t(8 * k), make_sdesc(vbase + uint32_t(k * 16 * 128), 8192, 1024), idesc_pv, (j | k) ? 1u : 0u);
        tc_commit(&v_empty[stage]);
        tc_commit(o_done);
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    auto softmax_bar = [&]() { asm volatile("bar.sync 2, 256;" ::: "memory"); };
    float m_run = -INFINITY, l_run = 0.f;
    const float sc = p.scale_log2;
    for (int j = 0; j < nkv; ++j) {
      const int valid = p.N - j * ATT_BKV - 64 * half;   // my columns [0, valid) are real keys
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s[64];
      tmem_ld32(tS + lane_off + uint32_t(64 * half), *reinterpret_cast<uint32_t(*)[32]>(s));
      tmem_ld32(tS + lane_off + uint32_t(64 * half + 32), *reinterpret_cast<uint32_t(*)[32]>(s + 32));
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(s_free);
      if (valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) s[i] = 0xff800000u;
      }
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(s[i]));
        m1 = fmaxf(m1, __uint_as_float(s[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(s[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(s[i + 3]));
      }
      float mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      float* xm = xmax + (j & 1) * 256;
      xm[half * 128 + row] = mt;
      softmax_bar();
      mt = fmaxf(mt, xm[(half ^ 1) * 128 + row]);
      const float mt_sc = mt * sc;
      bool waited_pv = (j == 0);
      if (j > 0) {
        const bool need = mt_sc > m_run + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(o_done, (j - 1) & 1);
          tc_fence_after();
          waited_pv = true;
          float f = 1.0f;
          if (need) {
            f = ex2_approx(m_run - mt_sc);
            m_run = mt_sc;
            l_run *= f;
          }
          uint32_t v[32];
          tmem_ld32(tO + lane_off + uint32_t(32 * half), v);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          tmem_st32(tO + lane_off + uint32_t(32 * half), v);
          tc_wait_st();
        }
      } else {
        m_run = mt_sc;
      }
      const float neg_m = -m_run;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(s[c * 32 + i]), sc, neg_m));
          const float p1 = ex2_approx(fmaf(__uint_as_float(s[c * 32 + i + 1]), sc, neg_m));
          l0 += p0;
          l1 += p1;
          pk[i >> 1] = O16::pack(p0, p1);
        }
        if (c == 0 && !waited_pv) {
          mbar_wait(o_done, (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st16(tP + lane_off + uint32_t(32 * half + 16 * c), pk);
      }
      l_run += l0 + l1;
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue: combine the two partial row sums, O / l -> 16-bit (32 columns per thread)
    xsum[half * 128 + row] = l_run;
    softmax_bar();
    const float l_tot = l_run + xsum[(half ^ 1) * 128 + row];
    mbar_wait(o_done, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_tot;
    const int qrow = q0 + row;
    if (half == 0 && p.lse != nullptr && qrow < p.N) p.lse[(long(b) * p.H + h) * p.N + qrow] = m_run + log2f(l_tot);
    typename O16::T* dst = reinterpret_cast<typename O16::T*>(p.out) + long(row_base + qrow) * p.ld_out + h * ATT_D + 32 * half;
    uint32_t v[32];
    tmem_ld32(tO + lane_off + uint32_t(32 * half), v);
    tc_wait_ld();
    if (qrow < p.N) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        st_global_v4(dst + i, O16::pack(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l),
                     O16::pack(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l),
                     O16::pack(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l),
                     O16::pack(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace mb
