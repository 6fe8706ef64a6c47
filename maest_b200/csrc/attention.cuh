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

#ifdef ATT_DIAG_CLOCKS   // timing diagnostic: per-phase clock64 deltas of softmax thread 0, written over lse[b,h,q0..q0+7]
#define ATT_CLK(i) do { const long long t_ = clock64(); dg[i] += t_ - tprev; tprev = t_; } while (0)
#define ATT_CLK_ARGS , dg, tprev
#else
#define ATT_CLK(i) do { } while (0)
#define ATT_CLK_ARGS
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
#ifdef ATT_DIAG_CLOCKS
    long long dg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
    const long long tstart = tprev;
#endif
    for (int j = 0; j < nkv; ++j) {
      const int kv0 = j * ATT_BKV;
      const int valid = p.N - kv0;  // keys [0, valid) of this tile are real (>=128 when not the last tile)
      const bool full_tile = valid >= ATT_BKV;
      mbar_wait(s_full, j & 1);
      ATT_CLK(0);
      tc_fence_after();
      // the whole score row goes to registers in one shot; the TMEM columns are handed back immediately
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tS + lane_off + uint32_t(c * 32), *reinterpret_cast<uint32_t(*)[32]>(s + 32 * c));
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(s_free);
      ATT_CLK(1);
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
      ATT_CLK(2);
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
      ATT_CLK(3);
      const float neg_m = -m_run;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a0 = fmaf(__uint_as_float(s[c * 32 + i]), sc, neg_m);
          const float a1 = fmaf(__uint_as_float(s[c * 32 + i + 1]), sc, neg_m);
#ifdef ATT_DIAG_NOEXP
          const float p0 = a0 * 1e-3f;   // timing diagnostic only: no MUFU in the hot loop (results are garbage)
          const float p1 = a1 * 1e-3f;
#else
          const float p0 = ex2_approx(a0);
          #if ATT_POLY_EVERY > 0
          const float p1 = (((i + 1) % ATT_POLY_EVERY) == ATT_POLY_EVERY - 1) ? ex2_poly(a1) : ex2_approx(a1);
#else
          const float p1 = ex2_approx(a1);
#endif
#endif
          l0 += p0;
          l1 += p1;
          pk[i >> 1] = O16::pack(p0, p1);
        }
        if (c == 0 && !waited_pv) {
          ATT_CLK(4);
          mbar_wait(o_done, (j - 1) & 1);
          ATT_CLK(5);
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
      ATT_CLK(4);
      if constexpr (P_IN_TMEM) tc_wait_st();
      else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      ATT_CLK(6);
    }
    // ---------------- epilogue: O / l -> 16-bit ----------------
    mbar_wait(o_done, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int qrow = q0 + row;
#ifdef ATT_DIAG_CLOCKS
    if (p.lse != nullptr && row == 0 && q0 + 8 < p.N) {
      dg[7] = clock64() - tstart;
      for (int i = 0; i < 8; ++i) p.lse[(long(b) * p.H + h) * p.N + q0 + i] = float(dg[i]);
    }
#else
    if (p.lse != nullptr && qrow < p.N) p.lse[(long(b) * p.H + h) * p.N + qrow] = m_run + log2f(l_run);
#endif
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
// Default kernel ("speculative max", variant 0).  Same pipeline and TMEM/smem plan as attention_fwd_kernel<DT, true>, with
// the two changes the per-phase clock measurements asked for (profiles/r01_attention_phase_clocks.md):
//  * the row max of a KV tile is no longer computed BEFORE the exponentials.  Tiles j >= 1 exponentiate against the
//    running max of the previous tiles and fold the max of the shifted scores into the same loop (FMNMX on the ALU pipe
//    while the MUFU pipe is the busy one); only if some row's tile max exceeds the running max by more than 2^8 (the
//    same bound the lazy rescaling uses) the warp re-does the tile against the new max.  The max pass + vote + branch
//    (~370 of ~2360 cycles per tile on the critical path of a softmax warp) disappear from the common path.
//  * the last KV tile is peeled: only ceil32(valid keys) score columns are computed (QK^T with N = 32..128), read,
//    exponentiated and multiplied (PV with K = 32..128), instead of masking a full 128-wide tile.
// ------------------------------------------------------------------------------------------------------------------
template <int DT, int NC>
__device__ __forceinline__ void att_spec_tile(const int j, const int valid, const float sc, const uint32_t tS, const uint32_t tO,
                                              const uint32_t tP, const uint32_t lane_off, uint64_t* s_full, uint64_t* s_free,
                                              uint64_t* p_full, uint64_t* o_done, float& m_run, float& l_run
#ifdef ATT_DIAG_CLOCKS
                                              , long long (&dg)[8], long long& tprev
#endif
                                              ) {
  using O16 = Op16<DT>;
  mbar_wait(s_full, j & 1);
  ATT_CLK(0);
  tc_fence_after();
  uint32_t s[NC];
#pragma unroll
  for (int c = 0; c < NC / 32; ++c) tmem_ld32(tS + lane_off + uint32_t(c * 32), *reinterpret_cast<uint32_t(*)[32]>(s + 32 * c));
  tc_wait_ld();
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(s_free);   // S_j is in registers (one arrival per warp): the tensor pipe may overwrite it with S_{j+1}
  ATT_CLK(1);
  if (valid < NC) {
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (i >= valid) s[i] = 0xff800000u;   // -inf: masked keys contribute exp2(-inf) = 0
  }
  if (j == 0) {                 // first tile: there is no running max to speculate against
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < NC; i += 4) {
      m0 = fmaxf(m0, __uint_as_float(s[i]));
      m1 = fmaxf(m1, __uint_as_float(s[i + 1]));
      m2 = fmaxf(m2, __uint_as_float(s[i + 2]));
      m3 = fmaxf(m3, __uint_as_float(s[i + 3]));
    }
    m_run = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sc;
  }
  ATT_CLK(2);
  float l0 = 0.f, l1 = 0.f;
  float x0 = -INFINITY, x1 = -INFINITY;     // max of the shifted scores a = s*sc - m_run
  {
    const float neg_m = -m_run;
#pragma unroll
    for (int c = 0; c < NC / 32; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float a0 = fmaf(__uint_as_float(s[c * 32 + i]), sc, neg_m);
        const float a1 = fmaf(__uint_as_float(s[c * 32 + i + 1]), sc, neg_m);
        x0 = fmaxf(x0, a0);
        x1 = fmaxf(x1, a1);
        const float p0 = ex2_approx(a0);
        const float p1 = ex2_approx(a1);
        l0 += p0;
        l1 += p1;
        pk[i >> 1] = O16::pack(p0, p1);
      }
      if (c == 0 && j > 0) {    // PV_{j-1} must have consumed P_{j-1} before P_j overwrites it
        ATT_CLK(4);
        mbar_wait(o_done, (j - 1) & 1);
        ATT_CLK(5);
        tc_fence_after();
      }
      tmem_st16(tP + lane_off + uint32_t(c * 16), pk);
    }
  }
  ATT_CLK(4);
  const float amax = fmaxf(x0, x1);
  const bool exceed = amax > 8.0f;           // never on the first tile (amax == 0 there)
  if (__any_sync(0xffffffffu, exceed)) {
    // slow path (warp-uniform): raise the running max of the rows that need it, rescale their O and l, redo P_j.
    float f = 1.0f;
    if (exceed) {
      f = ex2_approx(-amax);
      m_run += amax;
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
    const float neg_m = -m_run;
    l0 = 0.f;
    l1 = 0.f;
#pragma unroll
    for (int c = 0; c < NC / 32; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[c * 32 + i]), sc, neg_m));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[c * 32 + i + 1]), sc, neg_m));
        l0 += p0;
        l1 += p1;
        pk[i >> 1] = O16::pack(p0, p1);
      }
      tmem_st16(tP + lane_off + uint32_t(c * 16), pk);
    }
  }
  l_run += l0 + l1;
  ATT_CLK(3);
  tc_wait_st();
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(p_full);   // one arrival per warp: 4 instead of 128 serialised mbarrier updates per tile
  ATT_CLK(6);
}

// Measured dead end (B200, config 3): 64-key tiles with 128 + 32 TMEM columns per CTA so that THREE CTAs co-reside per SM
// (three softmax warps per sub-partition): 0.97 ms vs 0.866 ms -- the per-tile handshakes double and eat the extra overlap.
template <int BKV> struct AttSpecCfg;
template <> struct AttSpecCfg<128> { static constexpr int kCtas = 2, kStages = 2, kSmemLaunch = ATT_TILE_BYTES * 5 + 128; };

template <int DT, int BKV>
__global__ void __launch_bounds__(ATT_THREADS, AttSpecCfg<BKV>::kCtas)
attention_fwd_spec_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using O16 = Op16<DT>;
  constexpr int STAGES = AttSpecCfg<BKV>::kStages;
  constexpr int KV_BYTES = BKV * ATT_D * 2;                         // one K or V tile
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;                              // [stages]
  uint8_t* sV = sK + STAGES * KV_BYTES;                             // [stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * KV_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [STAGES]
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* s_full = v_empty + STAGES;   // S_j written by the tensor pipe
  uint64_t* s_free = s_full + 1;         // S_j copied to registers by all 128 softmax threads (count 128)
  uint64_t* p_full = s_full + 2;         // P_j written (count 128)
  uint64_t* o_done = s_full + 3;         // O += P_j V_j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);   // [2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row_base = b * p.N;
  const int nkv = (p.N + BKV - 1) / BKV;
  const int valid_last = p.N - (nkv - 1) * BKV;           // 1..BKV real keys in the last tile
  const int nc_last = (valid_last + 31) & ~31;            // score columns computed for it

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("attention: dynamic smem base not 1024-aligned\n");
    __trap();
  }
  // Prologue off the critical path: the TMA thread initialises the load barriers itself and has Q, K_0 and V_0 in flight
  // before the CTA-wide sync; TMEM allocation and the compute barriers are set up by warp 5 meanwhile.
  if (warp == 4 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    fence_mbar_init();
    mbar_expect_tx(q_full, ATT_TILE_BYTES);
    tma_load_2d(sQ, &tmap_q, q_full, h * ATT_D, row_base + q0);
    mbar_expect_tx(&k_full[0], KV_BYTES);
    tma_load_2d(sK, &tmap_kv, &k_full[0], p.H * ATT_D + h * ATT_D, row_base);
    mbar_expect_tx(&v_full[0], KV_BYTES);
    tma_load_2d(sV, &tmap_kv, &v_full[0], 2 * p.H * ATT_D + h * ATT_D, row_base);
  }
  if (warp == 5) {
    if (lane == 0) {
      mbar_init(s_full, 1);
      mbar_init(s_free, 4);
      mbar_init(p_full, 4);
      mbar_init(o_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot[0];
  const uint32_t tS = tmem_base;                                   // BKV fp32 columns
  const uint32_t tO = tmem_base + BKV;                             // 64 fp32 columns
  const uint32_t tP = tmem_base + 192;                             // BKV/2 columns of packed 16-bit pairs

  if (warp == 4) {
    if (lane == 0) {
      int stage = 1 % STAGES;      // tile 0 was issued in the prologue
      uint32_t phase = STAGES == 1 ? 1 : 0;
      for (int j = 1; j < nkv; ++j) {
        const int r = row_base + j * BKV;
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_expect_tx(&k_full[stage], KV_BYTES);
        tma_load_2d(sK + stage * KV_BYTES, &tmap_kv, &k_full[stage], p.H * ATT_D + h * ATT_D, r);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_expect_tx(&v_full[stage], KV_BYTES);
        tma_load_2d(sV + stage * KV_BYTES, &tmap_kv, &v_full[stage], 2 * p.H * ATT_D + h * ATT_D, r);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc(DT, 128, BKV, 0, 0);
      const uint32_t idesc_qk_last = make_idesc(DT, 128, nc_last, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(DT, 128, 64, 0, 1);  // B = V, MN-major
      const uint64_t qdesc = make_sdesc(smem_u32(sQ), 16, 1024);
      auto issue_qk = [&](int stage, uint32_t idesc) {
        const uint64_t kdesc = make_sdesc(smem_u32(sK + stage * KV_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          mma_ss(tS, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc, k ? 1u : 0u);
        tc_commit(&k_empty[stage]);   // K stage reusable once S = Q K^T has retired
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, nkv == 1 ? idesc_qk_last : idesc_qk);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
        if (j + 1 < nkv) {
          // S_j lives in the softmax threads' registers now: overwrite it with S_{j+1} while they exponentiate
          mbar_wait(s_free, j & 1);
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          issue_qk(nstage, j + 2 == nkv ? idesc_qk_last : idesc_qk);
        }
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[stage], phase);
        tc_fence_after();
        // O (+)= P_j V_j : K-steps of 16 keys (only the computed columns of the last tile)
        const uint32_t vbase = smem_u32(sV + stage * KV_BYTES);
        const int ksteps = (j + 1 == nkv ? nc_last : BKV) / 16;
#pragma unroll 2
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t vdesc = make_sdesc(vbase + uint32_t(k * 16 * 128), 8192, 1024);
          mma_ts(tO, tP + uint32_t(8 * k), vdesc, idesc_pv, (j | k) ? 1u : 0u);
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
#ifdef ATT_DIAG_CLOCKS
    long long dg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
    const long long tstart = tprev;
#endif
    for (int j = 0; j + 1 < nkv; ++j)
      att_spec_tile<DT, BKV>(j, BKV, sc, tS, tO, tP, lane_off, s_full, s_free, p_full, o_done, m_run, l_run ATT_CLK_ARGS);
    switch (nc_last) {
      case 32: att_spec_tile<DT, 32>(nkv - 1, valid_last, sc, tS, tO, tP, lane_off, s_full, s_free, p_full, o_done, m_run, l_run ATT_CLK_ARGS); break;
      case 64: att_spec_tile<DT, 64>(nkv - 1, valid_last, sc, tS, tO, tP, lane_off, s_full, s_free, p_full, o_done, m_run, l_run ATT_CLK_ARGS); break;
      case 96: att_spec_tile<DT, 96>(nkv - 1, valid_last, sc, tS, tO, tP, lane_off, s_full, s_free, p_full, o_done, m_run, l_run ATT_CLK_ARGS); break;
      default: att_spec_tile<DT, 128>(nkv - 1, valid_last, sc, tS, tO, tP, lane_off, s_full, s_free, p_full, o_done, m_run, l_run ATT_CLK_ARGS); break;
    }
    // ---------------- epilogue: O / l -> 16-bit ----------------
    mbar_wait(o_done, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int qrow = q0 + row;
#ifdef ATT_DIAG_CLOCKS
    if (p.lse != nullptr && row == 0 && q0 + 8 < p.N) {
      dg[7] = clock64() - tstart;
      for (int i = 0; i < 8; ++i) p.lse[(long(b) * p.H + h) * p.N + q0 + i] = float(dg[i]);
    }
#else
    if (p.lse != nullptr && qrow < p.N) p.lse[(long(b) * p.H + h) * p.N + qrow] = m_run + log2f(l_run);
#endif
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
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace mb
