// Flash-style attention forward, persistent and software-pipelined: ONE CTA per SM works on TWO 128-query tiles of the same
// (clip, head) at a time.  Replaces models/maest.py:362-375 like attention.cuh does; same inputs/outputs.
//
// Why (profiles/r01_attention_phase_clocks.md): the 1-tile kernel keeps the MUFU pipe 66 % busy -- per KV tile a softmax warp
// spends ~1000 cycles on  wait S -> tcgen05.ld -> row max -> vote  with no exponential in flight, and the two warps that share
// an SM sub-partition tend to do so at the same time.  Here each softmax thread overlaps that work with its OWN exponentials:
// the scores of tile t+1 are loaded into the registers the scores of tile t free chunk by chunk, and masked / max-reduced in
// the issue slots the MUFU-bound loop leaves empty (att_pl_tile).  K/V tiles are fetched once per PAIR of query tiles (half
// the L2 -> smem traffic), and the kernel is persistent over (clip, head, tile-pair) work items so that the Q/K/V loads of the
// next item and the output epilogue of the previous one overlap the running one.
// Measured dead end: handing the MUFU pipe back and forth between the two warpgroups with named barriers (exponential phases
// mutually exclusive) -- 0.91 ms vs 0.86 ms: a single warp cannot saturate the pipe (11 cycles per exponential alone, 8.35 with
// two warps, tools/ubench/pipes.cu), so exclusivity wastes more than the interleaving wins.
//
//   warps 0..3   softmax warpgroup A (query tile 0 of the pair): thread = one row, TMEM lane = row
//   warps 4..7   softmax warpgroup B (query tile 1)
//   warp  8      TMA producer: Q_A, Q_B per item; K/V ring (3 stages) shared by both tiles
//   warp  9/10   MMA issuers for tile A / tile B: S = Q K^T (128 x 128 x 64), O += P V (128 x 64 x 128), P read from TMEM
//   TMEM (512 columns): tile X at 256 X: S [0,128) | O [128,192) | P [192,256)
#pragma once
#include "attention.cuh"

namespace mb {

constexpr int ATTP_THREADS = 352;
constexpr int ATTP_STAGES = 3;
constexpr int ATTP_SMEM_BYTES = ATT_TILE_BYTES * (2 + 2 * ATTP_STAGES) + 512;

__device__ __forceinline__ void att_mask_chunk(uint32_t* s32, int first_col, int valid) {
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (first_col + i >= valid) s32[i] = 0xff800000u;   // -inf: masked keys contribute exp2(-inf) = 0
}
__device__ __forceinline__ void att_max_chunk(const uint32_t* s32, float (&n)[4]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    n[0] = fmaxf(n[0], __uint_as_float(s32[i]));
    n[1] = fmaxf(n[1], __uint_as_float(s32[i + 1]));
    n[2] = fmaxf(n[2], __uint_as_float(s32[i + 2]));
    n[3] = fmaxf(n[3], __uint_as_float(s32[i + 3]));
  }
}

struct AttPlCtx {
  float sc;
  uint32_t tS, tO, tP, lane_off;
  uint64_t *s_full, *s_free, *p_full, *o_done;
};

// First KV tile of an item: nothing to overlap it with.  Loads S_t, releases the S columns, masks, takes the row max.
__device__ __forceinline__ void att_pl_first(uint32_t (&cur)[128], const uint32_t t, const int valid, const AttPlCtx& x, float& m_run) {
  mbar_wait(x.s_full, t & 1);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_ld32(x.tS + x.lane_off + uint32_t(c * 32), *reinterpret_cast<uint32_t(*)[32]>(cur + 32 * c));
  tc_wait_ld();
  tc_fence_before();
  mbar_arrive(x.s_free);
  float n[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (valid < 128) att_mask_chunk(cur + 32 * c, 32 * c, valid);
    att_max_chunk(cur + 32 * c, n);
  }
  m_run = fmaxf(fmaxf(n[0], n[1]), fmaxf(n[2], n[3])) * x.sc;
}

// One KV tile, software-pipelined inside the thread: while the 128 exponentials of tile t (scores already in `cur`, running
// max already covering them) keep the MUFU pipe busy, the scores of tile t+1 are pulled from TMEM into the registers that
// `cur` frees chunk by chunk, masked, and reduced to the row max in the issue slots the MUFU-bound loop leaves empty.
// The per-tile critical path of a softmax warp shrinks from  wait S -> tcgen05.ld -> max -> vote -> exp -> st  to  exp -> vote.
template <int DT, bool HAS_NEXT, bool MASK_NEXT>
__device__ __forceinline__ void att_pl_tile(uint32_t (&cur)[128], uint32_t (&nxt)[128], const uint32_t t, bool& pv_waited,
                                            const int valid_nxt, const AttPlCtx& x, float& m_run, float& l_run) {
  using O16 = Op16<DT>;
  const float neg_m = -m_run;
  float l0 = 0.f, l1 = 0.f;
  float n[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  uint32_t pk0[16];               // P chunk 0 is held back: P_{t-1} may still be feeding the tensor pipe
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float p0 = ex2_approx(fmaf(__uint_as_float(cur[c * 32 + i]), x.sc, neg_m));
      const float p1 = ex2_approx(fmaf(__uint_as_float(cur[c * 32 + i + 1]), x.sc, neg_m));
      l0 += p0;
      l1 += p1;
      if (c == 0) pk0[i >> 1] = O16::pack(p0, p1);
      else pk[i >> 1] = O16::pack(p0, p1);
    }
    if (c == 1) {
      // PV of tile t-1 must have consumed P_{t-1} before P_t overwrites it: it was issued when this thread finished tile
      // t-1 and has had 64 exponentials (~550-1100 cycles) to retire -- waiting after the first 32 stalled every tile.
      if (!pv_waited) {
        mbar_wait(x.o_done, (t - 1) & 1);
        tc_fence_after();
      }
      tmem_st16(x.tP + x.lane_off, pk0);
    }
    if (c >= 1) tmem_st16(x.tP + x.lane_off + uint32_t(c * 16), pk);
    if constexpr (HAS_NEXT) {
      if (c == 2) {               // cur[0..63] are dead: their registers receive the first half of S_{t+1}
        mbar_wait(x.s_full, (t + 1) & 1);
        tc_fence_after();
        tmem_ld32(x.tS + x.lane_off, *reinterpret_cast<uint32_t(*)[32]>(nxt));
        tmem_ld32(x.tS + x.lane_off + 32u, *reinterpret_cast<uint32_t(*)[32]>(nxt + 32));
      }
      if (c == 3) {
        tc_wait_ld();
        if constexpr (MASK_NEXT) { att_mask_chunk(nxt, 0, valid_nxt); att_mask_chunk(nxt + 32, 32, valid_nxt); }
        att_max_chunk(nxt, n);
        att_max_chunk(nxt + 32, n);
        tmem_ld32(x.tS + x.lane_off + 64u, *reinterpret_cast<uint32_t(*)[32]>(nxt + 64));
        tmem_ld32(x.tS + x.lane_off + 96u, *reinterpret_cast<uint32_t(*)[32]>(nxt + 96));
      }
    }
  }
  pv_waited = false;
  l_run += l0 + l1;
  tc_wait_st();
  tc_fence_before();
  mbar_arrive(x.p_full);          // P_t complete: the tensor pipe may run O += P_t V_t
  if constexpr (HAS_NEXT) {
    tc_wait_ld();
    tc_fence_before();
    mbar_arrive(x.s_free);        // S_{t+1} is in registers: the tensor pipe may overwrite it with S_{t+2}
    if constexpr (MASK_NEXT) { att_mask_chunk(nxt + 64, 64, valid_nxt); att_mask_chunk(nxt + 96, 96, valid_nxt); }
    att_max_chunk(nxt + 64, n);
    att_max_chunk(nxt + 96, n);
    const float mt_sc = fmaxf(fmaxf(n[0], n[1]), fmaxf(n[2], n[3])) * x.sc;
    const bool need = mt_sc > m_run + 8.0f;   // lazy rescale: p stays <= 2^8 against a stale max
    if (__any_sync(0xffffffffu, need)) {
      mbar_wait(x.o_done, t & 1);             // O must hold P_t V_t before it is corrected
      tc_fence_after();
      pv_waited = true;
      float f = 1.0f;
      if (need) {
        f = ex2_approx(m_run - mt_sc);
        m_run = mt_sc;
        l_run *= f;
      }
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(x.tO + x.lane_off + uint32_t(c * 32), v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
        tmem_st32(x.tO + x.lane_off + uint32_t(c * 32), v);
      }
      tc_wait_st();
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(ATTP_THREADS, 1)
attention_fwd_pp_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using O16 = Op16<DT>;
  uint8_t* sQ = smem;                                      // [2] query tiles of the pair
  uint8_t* sK = smem + 2 * ATT_TILE_BYTES;                 // [stages]
  uint8_t* sV = sK + ATTP_STAGES * ATT_TILE_BYTES;         // [stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATTP_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                  // TMA tx, both Q tiles
  uint64_t* q_empty = bars + 1;             // count 2: last QK of the item retired (one commit per MMA warp)
  uint64_t* k_full = bars + 2;              // [stages]
  uint64_t* k_empty = k_full + ATTP_STAGES; // [stages] count 2
  uint64_t* v_full = k_empty + ATTP_STAGES;
  uint64_t* v_empty = v_full + ATTP_STAGES; // count 2
  uint64_t* xbar = v_empty + ATTP_STAGES;   // per query tile X: s_full, s_free, p_full, o_done, o_free  at xbar[5 X + i]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkv = (p.N + ATT_BKV - 1) / ATT_BKV;
  const int valid_last = p.N - (nkv - 1) * ATT_BKV;       // 1..128 real keys in the last tile
  const int nc_last = (valid_last + 31) & ~31;
  const int npairs = (p.N + 2 * ATT_BQ - 1) / (2 * ATT_BQ);
  const int total = p.B * p.H * npairs;                   // work items, pair index fastest: concurrent CTAs share K/V in L2

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("attention_pp: dynamic smem base not 1024-aligned\n");
    __trap();
  }
  if (warp == 8 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);
    for (int i = 0; i < ATTP_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&xbar[5 * x + 0], 1);     // s_full
      mbar_init(&xbar[5 * x + 1], 128);   // s_free
      mbar_init(&xbar[5 * x + 2], 128);   // p_full
      mbar_init(&xbar[5 * x + 3], 1);     // o_done
      mbar_init(&xbar[5 * x + 4], 128);   // o_free
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_qkv);
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
        const int qp = item % npairs;
        const int bh = item / npairs;
        const int h = bh % p.H, b = bh / p.H;
        const int row_base = b * p.N;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
        tma_load_2d(sQ, &tmap_qkv, q_full, h * ATT_D, row_base + qp * 2 * ATT_BQ);
        tma_load_2d(sQ + ATT_TILE_BYTES, &tmap_qkv, q_full, h * ATT_D, row_base + qp * 2 * ATT_BQ + ATT_BQ);
        for (int j = 0; j < nkv; ++j) {
          const int r = row_base + j * ATT_BKV;
          mbar_wait(&k_empty[stage], phase ^ 1);
          mbar_expect_tx(&k_full[stage], ATT_TILE_BYTES);
          tma_load_2d(sK + stage * ATT_TILE_BYTES, &tmap_qkv, &k_full[stage], p.H * ATT_D + h * ATT_D, r);
          mbar_wait(&v_empty[stage], phase ^ 1);
          mbar_expect_tx(&v_full[stage], ATT_TILE_BYTES);
          tma_load_2d(sV + stage * ATT_TILE_BYTES, &tmap_qkv, &v_full[stage], 2 * p.H * ATT_D + h * ATT_D, r);
          if (++stage == ATTP_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ------------------------------------------------ MMA issuer of query tile X
    if (lane == 0) {
      const int X = warp - 9;
      uint64_t* s_full = &xbar[5 * X + 0];
      uint64_t* s_free = &xbar[5 * X + 1];
      uint64_t* p_full = &xbar[5 * X + 2];
      uint64_t* o_done = &xbar[5 * X + 3];
      uint64_t* o_free = &xbar[5 * X + 4];
      const uint32_t tS = tmem_base + uint32_t(256 * X);
      const uint32_t tO = tS + 128;
      const uint32_t tP = tS + 192;
      constexpr uint32_t idesc_qk = make_idesc(DT, 128, 128, 0, 0);
      const uint32_t idesc_qk_last = make_idesc(DT, 128, nc_last, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(DT, 128, 64, 0, 1);  // B = V, MN-major
      const uint64_t qdesc = make_sdesc(smem_u32(sQ + X * ATT_TILE_BYTES), 16, 1024);
      auto issue_qk = [&](int stage, uint32_t idesc) {
        const uint64_t kdesc = make_sdesc(smem_u32(sK + stage * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          mma_ss(tS, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc, k ? 1u : 0u);
        tc_commit(&k_empty[stage]);   // K stage reusable once both tiles' S = Q K^T have retired (count 2)
        tc_commit(s_full);
      };
      int stage = 0;
      uint32_t phase = 0;
      uint32_t t = 0;                 // running KV-tile counter of this query tile (barrier parities)
      uint32_t it = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
        mbar_wait(q_full, it & 1);
        if (t > 0) mbar_wait(s_free, (t - 1) & 1);   // last S of the previous item has been read
        mbar_wait(&k_full[stage], phase);
        tc_fence_after();
        issue_qk(stage, nkv == 1 ? idesc_qk_last : idesc_qk);
        for (int j = 0; j < nkv; ++j, ++t) {
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == ATTP_STAGES) { nstage = 0; nphase ^= 1; }
          if (j + 1 < nkv) {
            mbar_wait(s_free, t & 1);
            mbar_wait(&k_full[nstage], nphase);
            tc_fence_after();
            issue_qk(nstage, j + 2 == nkv ? idesc_qk_last : idesc_qk);
          } else {
            tc_commit(q_empty);       // every QK of this item is issued: Q is reusable once they retire
          }
          mbar_wait(p_full, t & 1);
          mbar_wait(&v_full[stage], phase);
          if (j == 0) mbar_wait(o_free, (it & 1) ^ 1);   // the previous item's O has been read by its epilogue
          tc_fence_after();
          const uint32_t vbase = smem_u32(sV + stage * ATT_TILE_BYTES);
          const int ksteps = (j + 1 == nkv ? nc_last : ATT_BKV) / 16;
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
    }
  } else {
    // ------------------------------------------------ softmax warpgroups: thread <-> query row (TMEM lane)
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    uint64_t* s_full = &xbar[5 * wg + 0];
    uint64_t* s_free = &xbar[5 * wg + 1];
    uint64_t* p_full = &xbar[5 * wg + 2];
    uint64_t* o_done = &xbar[5 * wg + 3];
    uint64_t* o_free = &xbar[5 * wg + 4];
    const uint32_t tS = tmem_base + uint32_t(256 * wg);
    const uint32_t tO = tS + 128;
    const uint32_t tP = tS + 192;
    AttPlCtx x;
    x.sc = p.scale_log2; x.tS = tS; x.tO = tO; x.tP = tP; x.lane_off = lane_off;
    x.s_full = s_full; x.s_free = s_free; x.p_full = p_full; x.o_done = o_done;
    uint32_t t = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int qp = item % npairs;
      const int bh = item / npairs;
      const int h = bh % p.H, b = bh / p.H;
      const int row_base = b * p.N;
      const int q0 = qp * 2 * ATT_BQ + wg * ATT_BQ;
      float m_run = -INFINITY;  // running max of s * scale_log2
      float l_run = 0.f;
      uint32_t sA[128], sB[128];
      bool pv_waited = true;    // the previous item's epilogue has waited for its last PV
      att_pl_first(sA, t, nkv == 1 ? valid_last : 128, x, m_run);
      // tiles 0 .. nkv-3 prefetch a full next tile; tile nkv-2 prefetches the (possibly partial, masked) last one
      for (int j = 0;;) {
        if (j + 1 >= nkv) { att_pl_tile<DT, false, false>(sA, sB, t, pv_waited, 128, x, m_run, l_run); ++t; break; }
        if (j + 2 == nkv && valid_last < 128) att_pl_tile<DT, true, true>(sA, sB, t, pv_waited, valid_last, x, m_run, l_run);
        else att_pl_tile<DT, true, false>(sA, sB, t, pv_waited, 128, x, m_run, l_run);
        ++t; ++j;
        if (j + 1 >= nkv) { att_pl_tile<DT, false, false>(sB, sA, t, pv_waited, 128, x, m_run, l_run); ++t; break; }
        if (j + 2 == nkv && valid_last < 128) att_pl_tile<DT, true, true>(sB, sA, t, pv_waited, valid_last, x, m_run, l_run);
        else att_pl_tile<DT, true, false>(sB, sA, t, pv_waited, 128, x, m_run, l_run);
        ++t; ++j;
      }
      // ---------------- epilogue: O / l -> 16-bit ----------------
      mbar_wait(o_done, (t - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const int qrow = q0 + row;
      if (p.lse != nullptr && qrow < p.N) p.lse[(long(b) * p.H + h) * p.N + qrow] = m_run + log2f(l_run);
      typename O16::T* dst = reinterpret_cast<typename O16::T*>(p.out) + long(row_base + qrow) * p.ld_out + h * ATT_D;
      uint32_t v0[32], v1[32];
      tmem_ld32(tO + lane_off, v0);
      tmem_ld32(tO + lane_off + 32u, v1);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(o_free);            // O is in registers: the next item's first PV may overwrite it
      if (qrow < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          st_global_v4(dst + i,
                       O16::pack(__uint_as_float(v0[i]) * inv_l, __uint_as_float(v0[i + 1]) * inv_l),
                       O16::pack(__uint_as_float(v0[i + 2]) * inv_l, __uint_as_float(v0[i + 3]) * inv_l),
                       O16::pack(__uint_as_float(v0[i + 4]) * inv_l, __uint_as_float(v0[i + 5]) * inv_l),
                       O16::pack(__uint_as_float(v0[i + 6]) * inv_l, __uint_as_float(v0[i + 7]) * inv_l));
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          st_global_v4(dst + 32 + i,
                       O16::pack(__uint_as_float(v1[i]) * inv_l, __uint_as_float(v1[i + 1]) * inv_l),
                       O16::pack(__uint_as_float(v1[i + 2]) * inv_l, __uint_as_float(v1[i + 3]) * inv_l),
                       O16::pack(__uint_as_float(v1[i + 4]) * inv_l, __uint_as_float(v1[i + 5]) * inv_l),
                       O16::pack(__uint_as_float(v1[i + 6]) * inv_l, __uint_as_float(v1[i + 7]) * inv_l));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mb
