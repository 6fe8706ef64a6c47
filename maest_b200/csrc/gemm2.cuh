// CTA-pair (cta_group::2) version of the forward GEMM:  C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue), K-major operands.
//
// Two CTAs on the two SMs of a TPC form a cluster and compute one 256 x 256 output tile: each CTA TMA-loads its own 128 rows
// of A and HALF of the W tile (128 of the 256 N-rows), the leader's single MMA thread issues tcgen05.mma.cta_group::2
// (M = 256, N = 256, K = 16) which reads both CTAs' shared memory and writes 128 accumulator rows into each CTA's TMEM.
// Versus the 1-CTA kernel the W operand is read from shared memory once per pair instead of once per CTA (smem operand
// traffic per SM per MMA: 8 KB instead of 12 KB) and a stage shrinks to 32 KB, so the TMA ring is 6 deep instead of 4.
// Epilogue: identical code (gemm_epilogue_subtile), each CTA drains its own 128 rows.
#pragma once
#include "gemm.cuh"

namespace mb {

constexpr int GEMM2_STAGES = 6;
constexpr int GEMM2_STAGE_BYTES = 2 * GEMM_A_BYTES;   // A: 128 x 64, W half: 128 x 64
constexpr int GEMM2_SMEM_BYTES = GEMM2_STAGES * GEMM2_STAGE_BYTES + GEMM_EPI_WARPS * GEMM_STG_BYTES + 1024 + 256;

template <int DT, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + GEMM2_STAGES * GEMM2_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + GEMM_EPI_WARPS * GEMM_STG_BYTES);
  uint64_t* full_bar = bars;                            // [STAGES]  (only the leader's copies are waited on)
  uint64_t* empty_bar = bars + GEMM2_STAGES;            // [STAGES]  (each CTA waits on its own; signalled by multicast commit)
  uint64_t* tfull_bar = bars + 2 * GEMM2_STAGES;        // [2]       (multicast commit)
  uint64_t* tempty_bar = bars + 2 * GEMM2_STAGES + 2;   // [2]       (leader's copy: one arrival per epilogue warp of both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GEMM2_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_pairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;
  const int num_m = (p.M + 255) / 256;
  const int num_n = (p.N + GEMM_BN - 1) / GEMM_BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < GEMM2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2cta<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();          // both CTAs' barriers initialised and TMEM allocated before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int m0 = (t / num_n) * 256 + int(rank) * 128;
        const int n0 = (t % num_n) * GEMM_BN + int(rank) * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * GEMM2_STAGE_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * GEMM2_STAGE_BYTES);   // bytes of BOTH CTAs land on the leader's barrier
          tma_load_2d_2cta(sa, &tmap_a, &full_bar[stage], kb * GEMM_BK, m0);
          tma_load_2d_2cta(sa + GEMM_A_BYTES, &tmap_b, &full_bar[stage], kb * GEMM_BK, n0);
          if (++stage == GEMM2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(DT, 256, GEMM_BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * GEMM_BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * GEMM2_STAGE_BYTES);
          const uint64_t adesc = make_sdesc(sa, 16, 1024);
          const uint64_t bdesc = make_sdesc(sa + GEMM_A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            mma_ss_2cta(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) ? 1u : 0u);
          tc_commit_2cta(&empty_bar[stage]);
          if (++stage == GEMM2_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_2cta(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    const int q = warp & 3;
    const int half = e >> 2;
    uint8_t* stg = stg_base + e * GEMM_STG_BYTES;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int m0 = (t / num_n) * 256 + int(rank) * 128 + q * 32;
      const int n0 = (t % num_n) * GEMM_BN + half * 128;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * GEMM_BN + half * 128);
      gemm_epilogue_subtile<DT, EPI>(p, &tmap_c, &tmap_c2, stg, taddr, m0, n0, lane, &tfull_bar[as], aphase,
                                     [&]() { mbar_arrive_cluster(&tempty_bar[as], 0); });
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();   // bulk stores of this warp have landed before the CTA exits
  }

  tc_fence_before();
  cluster_sync_all();          // no CTA may exit (or free TMEM) while its pair can still touch its smem / TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

}  // namespace mb
