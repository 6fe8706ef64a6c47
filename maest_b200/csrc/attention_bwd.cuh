// Flash-style attention BACKWARD for sm_100a (autograd of models/maest.py:362-375), d_head = 64.
//
// CTA = one 128-key tile of one (clip, head); it keeps K_j, V_j resident, streams the query tiles (Q_i, dO_i) and
// accumulates dK_j, dV_j in TMEM.  Per (i, j) five tcgen05 GEMMs:
//     S  = Q_i K_j^T            dP = dO_i V_j^T                       (128x128x64 each, fp32 in TMEM)
//     dV += P^T dO_i            dK += dS^T Q_i        dQ_i = dS K_j    (128x64x128 each)
// with  P = exp2(S*c - lse),  dS = P * (dP - delta) * scale  computed by 128 threads (one query row each) and
// written as 16-bit tiles to shared memory in the SWIZZLE_128B layout, where the same bytes serve as a K-major
// A operand (dQ) and as an MN-major A operand (dV, dK).  Q, K, dO are consumed in their natural [row][d] layout
// (K-major for S / dP, MN-major B for dK / dQ / dV) — no transposes are materialised anywhere.
// TWO compute warpgroups (round 2): warps w and w + 4 own the same 32 query rows (TMEM lanes) and split every row's 128 keys in
// halves -- which are also the two [128 x 64] halves of the P / dS shared-memory tiles, the two 32-column halves of the dQ
// staging tile and (at the end) dK vs dV.  With one warp per sub-partition the P / dS arithmetic (2 tcgen05.ld, one exponential,
// ~12 instructions and two packs per score) was a 2500-cycle serial phase per (query tile, key tile) pair next to ~1800 cycles
// of MMA issue; the kernel ran at 7300 cycles per pair.
// dQ tiles are reduced across key tiles by the TMA engine (cp.reduce.async.bulk.tensor .add on fp32, staged through a
// swizzled smem tile) into dq32; dK/dV are written once as 16-bit.
// PERSISTENT (round 2, late): one CTA per SM walks the (clip, head, key tile) work items.  The one-item CTA spent ~11 900 of its
// ~44 000 cycles outside the steady state (launch, barrier init, TMEM allocation, the first K / V / Q / dO round trip, the dK / dV
// read-out with nothing else running); now K / V are double-buffered so the next item's tiles arrive during the current one, its
// first S / dP GEMMs are issued while the compute warps drain dK / dV, and every barrier keeps its phase across items (parities
// come from a per-CTA count of query-tile iterations).
#pragma once
#include "attention.cuh"

namespace mb {

#ifndef ATTB_COMPUTE_WARPS
#define ATTB_COMPUTE_WARPS 16
#endif
constexpr int ATTB_CWARPS = ATTB_COMPUTE_WARPS;      // compute warps: 8 (two warpgroups, two 32-key chunks each) or 16 (four, one each)
constexpr int ATTB_NWG = ATTB_CWARPS / 4;
constexpr int ATTB_CPW = 4 / ATTB_NWG;               // 32-key chunks of a row per warpgroup
static_assert(ATTB_CWARPS == 8 || ATTB_CWARPS == 16, "two or four compute warpgroups");
// + one warpgroup of TMA warp, two MMA-issuing warps and an idle warp, + one warpgroup that reads dQ_i out of TMEM and hands it to
// the TMA reduce (that read-out was 730 of the compute warps' ~4900 cycles per iteration, and they are the critical path)
constexpr int ATTB_THREADS = (ATTB_CWARPS + 8) * 32;
// setmaxnreg moves registers inside the CTA's allocation AT LAUNCH (threads x the kernel's register count: 768 x 80 here), not
// inside the whole register file: what the compute warps take must have been released by the two auxiliary warpgroups -- a
// setmaxnreg.inc that asks for more blocks forever (bring-up: 104 / 40 hung with the compute warps parked on it).
constexpr int ATTB_REGS_LAUNCH = ATTB_CWARPS == 16 ? 80 : 128;      // what ptxas settles on under __launch_bounds__(ATTB_THREADS, 1)
constexpr int ATTB_REGS_COMPUTE = ATTB_CWARPS == 16 ? 104 : 168, ATTB_REGS_AUX = ATTB_CWARPS == 16 ? 32 : 48;
static_assert(ATTB_CWARPS * 32 * (ATTB_REGS_COMPUTE - ATTB_REGS_LAUNCH) <= 256 * (ATTB_REGS_LAUNCH - ATTB_REGS_AUX), "setmaxnreg budget");
constexpr int ATTB_SMEM_BYTES = ATT_TILE_BYTES * 14 + 128;   // (K, V)[2], Q[2], dO[2], P (2 halves), dS (2 halves), dQ staging (2 x [128 x 32] fp32)
static_assert(ATTB_SMEM_BYTES <= 227 * 1024, "shared memory budget");

// Of every 8 score pairs of a full tile, ATTB_NPOLY take their exponential on the FMA pipe (Cody-Waite + minimax cubic, relative
// error 7.7e-5, as in the forward chains kernel): 16 compute warps x 32 exponentials are 1024 MUFU cycles per sub-partition and
// iteration, the longest single item of the P / dS phase.
#ifndef ATTB_NPOLY
#define ATTB_NPOLY 0      // measured (same box, 64 x 866): 0/8 0.820 ms, 2/8 0.835 ms, 3/8 and 4/8 spill under the register cap (1.15 ms):
                          // the phase is as much issue-bound as MUFU-bound, moving exponentials to the FMA pipe buys nothing
#endif
// 2^a for a packed pair, a <= 0 (clamped at -126, where the result is 1.2e-38 ~ 0)
__device__ __forceinline__ void attb_ex2_poly2(const u64 a_in, float& p0, float& p1) {
  float a0, a1;
  f2_unpack(a_in, a0, a1);
  const u64 a = f2_packf(fmaxf(a0, -126.0f), fmaxf(a1, -126.0f));
  const u64 t = f2_add(a, f2_packf(12582912.f, 12582912.f));
  const u64 f = f2_add(t, f2_packf(-12582912.f, -12582912.f));
  const u64 r = f2_sub(a, f);
  u64 q = f2_fma(r, f2_packf(0.05519810691475868f, 0.05519810691475868f), f2_packf(0.24267712235450745f, 0.24267712235450745f));
  q = f2_fma(q, r, f2_packf(0.6932618021965027f, 0.6932618021965027f));
  q = f2_fma(q, r, f2_packf(0.9999227523803711f, 0.9999227523803711f));
  float q0, q1, t0, t1;
  f2_unpack(q, q0, q1);
  f2_unpack(t, t0, t1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

struct AttnBwdParams {
  int B, N, H;
  const float* lse;     // [B, H, N]
  const float* delta;   // [B, H, N]
  float* dq32;          // [B*N, H*64] fp32, zero-initialised by the caller (atomically accumulated)
  void* dqkv16;         // [B*N, 3*H*64] 16-bit: the k and v column blocks are written here
  float scale_log2, scale;
};

template <int DT>
__global__ void __launch_bounds__(ATTB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                     const __grid_constant__ CUtensorMap tmap_dq, const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using O16 = Op16<DT>;
  uint8_t* sKV = smem;                        // [2 buffers] x (K, V)
  uint8_t* sQ = smem + 4 * ATT_TILE_BYTES;    // [2]
  uint8_t* sdO = smem + 6 * ATT_TILE_BYTES;   // [2]
  uint8_t* sP = smem + 8 * ATT_TILE_BYTES;    // two [128 q x 64 keys] halves
  uint8_t* sdS = smem + 10 * ATT_TILE_BYTES;  // two halves
  uint8_t* sdQ = smem + 12 * ATT_TILE_BYTES;  // two [128 rows x 32 fp32] SWIZZLE_128B halves (TMA reduce source)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATTB_SMEM_BYTES - 128);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_empty = bars + 2;     // [2] every GEMM of the item that used this K / V buffer has retired (one commit per issuing warp)
  uint64_t* qdo_full = bars + 4;     // [2]
  uint64_t* qdo_empty = bars + 6;    // [2]
  uint64_t* sdp_full = bars + 8;     // S_i, dP_i in TMEM
  uint64_t* pds_full = bars + 9;     // P_i, dS_i in smem (count 128)
  uint64_t* mma2_done = bars + 10;   // dV/dK/dQ GEMMs of iteration i retired
  uint64_t* sdp_free = bars + 11;    // every compute warp has S_i / dP_i in registers: the TMEM buffers may take S_{i+1} / dP_{i+1}
  uint64_t* dkv_free = bars + 12;    // every compute warp has read the item's dK / dV out of TMEM: the next item may overwrite them
  uint64_t* dq_free = bars + 13;     // the drain warpgroup has dQ_i in registers: the accumulator may take dQ_{i+1}
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (p.N + 127) / 128;           // query tiles = key tiles per (clip, head)
  const int n_items = p.B * p.H * nq;
  // item -> (clip, head, key tile): key tile fastest, so the CTAs running side by side share Q / dO of a (clip, head) in L2
  auto item_kv0 = [&](int it) { return (it % nq) * 128; };
  auto item_h = [&](int it) { return (it / nq) % p.H; };
  auto item_b = [&](int it) { return it / (nq * p.H); };

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("attention_bwd: dynamic smem base not 1024-aligned\n");
    __trap();
  }
  if (warp == ATTB_CWARPS + 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 2);    // one commit per issuing warp
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, ATTB_CWARPS * 32);
    mbar_init(mma2_done, 2);
    mbar_init(sdp_free, ATTB_CWARPS);
    mbar_init(dkv_free, ATTB_CWARPS);
    mbar_init(dq_free, 4);
    fence_mbar_init();
  }
  if (warp == ATTB_CWARPS) {
    if (lane == 0) { tma_prefetch_desc(&tmap_qkv); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_dq); }
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320, tdQ = tmem_base + 384;

  // Parities.  g counts this CTA's query-tile iterations over all of its items: Q / dO stage = g & 1 with phase (g >> 1) & 1;
  // sdp_full, pds_full, mma2_done and sdp_free complete once per iteration (parity g & 1).  n counts its items: K / V buffer
  // n & 1 with phase (n >> 1) & 1; dkv_free completes once per item (parity n & 1).
  // (each setmaxnreg sits at the top of its role's branch: ptxas budgets registers per branch only then)
  if (warp >= ATTB_CWARPS + 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATTB_REGS_AUX));
    // dQ_i: TMEM -> swizzled smem -> global fp32 reduce-add by the TMA engine.  Rows of the tile that lie beyond this
    // clip carry dS = 0, hence dQ = 0, so adding them to the next clip's rows is harmless; rows beyond the tensor are clipped.
    const int dw = warp & 3;                       // TMEM lane quarter (ATTB_CWARPS + 4 is a multiple of 4)
    const int row = dw * 32 + lane;
    const uint32_t lane_off = uint32_t(dw * 32) << 16;
    const bool elected = (warp == ATTB_CWARPS + 4) && lane == 0;     // issues (and waits for) every bulk reduce of the CTA
    auto drain_bar = [&]() { asm volatile("bar.sync 2, 128;" ::: "memory"); };
    int g = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int h = item_h(it), row_base = item_b(it) * p.N;
      for (int i = 0; i < nq; ++i, ++g) {
        mbar_wait(mma2_done, g & 1);
        tc_fence_after();
        if (elected) tma_store_wait_read<0>();     // the previous reduce has finished reading the staging tiles
        drain_bar();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {              // 16 of the 64 dQ columns = half of a [128 x 32] fp32 staging tile
          uint32_t v[16];
          tmem_ld16(tdQ + lane_off + uint32_t(c * 16), v);
          tc_wait_ld();
          if (c == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_free);
          }
          uint8_t* base = sdQ + (c >> 1) * ATT_TILE_BYTES + row * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<uint4*>(base + (((4 * (c & 1) + q4) ^ (row & 7)) << 4)) = make_uint4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
        }
        fence_proxy_async_smem();
        drain_bar();
#ifndef ATTB_DIAG_NO_REDUCE     // timing diagnostic: dQ is not accumulated (results wrong)
        if (elected) {
          tma_reduce_add_2d(&tmap_dq, sdQ, h * 64, row_base + i * 128);
          tma_reduce_add_2d(&tmap_dq, sdQ + ATT_TILE_BYTES, h * 64 + 32, row_base + i * 128);
          tma_store_commit();
        }
#endif
      }
    }
    // the last reduce-add must have finished READING its staging tile before the CTA (and its shared memory) goes away; its
    // global writes complete on their own before the grid does (waiting for them here cost ~2 us per CTA)
    if (elected) tma_store_wait_read<0>();
  } else if (warp >= ATTB_CWARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATTB_REGS_AUX));
    if (warp == ATTB_CWARPS) {
    if (lane == 0) {
      int g = 0, n = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
        const int kv0 = item_kv0(it), h = item_h(it), row_base = item_b(it) * p.N;
        const int ks = n & 1;
        mbar_wait(&kv_empty[ks], ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[ks], 2 * ATT_TILE_BYTES);
        tma_load_2d(sKV + (2 * ks) * ATT_TILE_BYTES, &tmap_qkv, &kv_full[ks], p.H * 64 + h * 64, row_base + kv0);
        tma_load_2d(sKV + (2 * ks + 1) * ATT_TILE_BYTES, &tmap_qkv, &kv_full[ks], 2 * p.H * 64 + h * 64, row_base + kv0);
        for (int i = 0; i < nq; ++i, ++g) {
          const int stage = g & 1;
          mbar_wait(&qdo_empty[stage], ((g >> 1) & 1) ^ 1);
          mbar_expect_tx(&qdo_full[stage], 2 * ATT_TILE_BYTES);
          tma_load_2d(sQ + stage * ATT_TILE_BYTES, &tmap_qkv, &qdo_full[stage], h * 64, row_base + i * 128);
          tma_load_2d(sdO + stage * ATT_TILE_BYTES, &tmap_do, &qdo_full[stage], h * 64, row_base + i * 128);
        }
      }
    }
    } else if (warp == ATTB_CWARPS + 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(DT, 128, 128, 0, 0);    // S, dP: A K-major, B K-major
      constexpr uint32_t idesc_t = make_idesc(DT, 128, 64, 1, 1);     // dV, dK: A = P^T / dS^T (MN-major), B MN-major
      const uint32_t aP = smem_u32(sP);
      // S and dP of iteration g (Q / dO stage g & 1) against the K / V buffer of item n
      auto issue_s_dp = [&](int g, int n) {
        const int stage = g & 1, ks = n & 1;
        const uint64_t qd = make_sdesc(smem_u32(sQ + stage * ATT_TILE_BYTES), 16, 1024);
        const uint64_t od = make_sdesc(smem_u32(sdO + stage * ATT_TILE_BYTES), 16, 1024);
        const uint64_t kd = make_sdesc(smem_u32(sKV + (2 * ks) * ATT_TILE_BYTES), 16, 1024);
        const uint64_t vd = make_sdesc(smem_u32(sKV + (2 * ks + 1) * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(tS, qd + uint64_t(2 * k), kd + uint64_t(2 * k), idesc_s, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(tdP, od + uint64_t(2 * k), vd + uint64_t(2 * k), idesc_s, k ? 1u : 0u);
        tc_commit(sdp_full);
      };
      const int my_items = blockIdx.x < n_items ? (n_items - 1 - int(blockIdx.x)) / int(gridDim.x) + 1 : 0;
      const int total = my_items * nq;
      if (total > 0) {
        mbar_wait(&kv_full[0], 0);
        mbar_wait(&qdo_full[0], 0);
        tc_fence_after();
        issue_s_dp(0, 0);
      }
      int n = 0, i = 0;
      for (int g = 0; g < total; ++g) {
        const int stage = g & 1;
        // S / dP of the NEXT iteration go out as soon as the compute warps have read this one's out of TMEM -- not after their
        // whole P / dS phase (r02 clocks: the compute warps then idled 1900 of 4850 cycles per iteration waiting for the scores).
        // At an item boundary the next iteration reads the other K / V buffer, loaded while this item ran.
        if (g + 1 < total) {
          const int n_next = (i + 1 == nq) ? n + 1 : n;
          mbar_wait(sdp_free, g & 1);
          if (n_next != n) mbar_wait(&kv_full[n_next & 1], (n_next >> 1) & 1);
          mbar_wait(&qdo_full[(g + 1) & 1], ((g + 1) >> 1) & 1);
          tc_fence_after();
          issue_s_dp(g + 1, n_next);
        }
        mbar_wait(pds_full, g & 1);
        if (i == 0 && n > 0) mbar_wait(dkv_free, (n - 1) & 1);   // the previous item's dV has been read out
        tc_fence_after();
        const uint32_t aO = smem_u32(sdO + stage * ATT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dV[key, d] += sum_q P[q, key] dO[q, d]
          mma_ss(tdV, make_sdesc(aP + uint32_t(k * 2048), 16384, 1024), make_sdesc(aO + uint32_t(k * 2048), 8192, 1024), idesc_t,
                 (i | k) ? 1u : 0u);
        tc_commit(&qdo_empty[stage]);
        tc_commit(mma2_done);
        if (++i == nq) {
          tc_commit(&kv_empty[n & 1]);     // (this warp's last GEMM on the item's V; S / dP of the item were issued earlier)
          i = 0;
          ++n;
        }
      }
    }
    } else if (warp == ATTB_CWARPS + 2) {
    // second issuing warp: dK and dQ (both read dS).  A single thread issues one tcgen05.mma per ~53 cycles whatever its N
    // (profiles/r02_ubench_mma_issue.txt), so the 24 N = 64 instructions of an iteration were 1270 issue cycles on one warp;
    // dV (other warp, own accumulator) and dK / dQ (this warp, own accumulators) need no order between them.
    if (lane == 0) {
      constexpr uint32_t idesc_t = make_idesc(DT, 128, 64, 1, 1);
      constexpr uint32_t idesc_q = make_idesc(DT, 128, 64, 0, 1);     // dQ: A = dS (K-major), B = K (MN-major)
      const uint32_t aDS = smem_u32(sdS);
      int g = 0, n = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
        const uint32_t aK = smem_u32(sKV + (2 * (n & 1)) * ATT_TILE_BYTES);
        mbar_wait(&kv_full[n & 1], (n >> 1) & 1);
        for (int i = 0; i < nq; ++i, ++g) {
          const int stage = g & 1;
          mbar_wait(pds_full, g & 1);
          mbar_wait(&qdo_full[stage], (g >> 1) & 1);
          if (i == 0 && n > 0) mbar_wait(dkv_free, (n - 1) & 1);   // the previous item's dK has been read out
          tc_fence_after();
          const uint32_t aQ = smem_u32(sQ + stage * ATT_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dK[key, d] += sum_q dS[q, key] Q[q, d]
            mma_ss(tdK, make_sdesc(aDS + uint32_t(k * 2048), 16384, 1024), make_sdesc(aQ + uint32_t(k * 2048), 8192, 1024), idesc_t,
                   (i | k) ? 1u : 0u);
          // Q_i is not read again: the stage goes back to the TMA warp before the dQ GEMMs (and before the wait for the dQ
          // accumulator) -- its refill is a 2000-cycle round trip that the S / dP issue two iterations on waits for
          tc_commit(&qdo_empty[stage]);
          if (g > 0) {                  // the drain warpgroup holds dQ of the previous iteration in registers
            mbar_wait(dq_free, (g - 1) & 1);
            tc_fence_after();
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dQ[q, d] = sum_key dS[q, key] K[key, d]
            mma_ss(tdQ, make_sdesc(aDS + uint32_t((k >> 2) * ATT_TILE_BYTES), 16, 1024) + uint64_t(2 * (k & 3)),
                   make_sdesc(aK + uint32_t(k * 2048), 8192, 1024), idesc_q, k ? 1u : 0u);
          tc_commit(mma2_done);
        }
        tc_commit(&kv_empty[n & 1]);
      }
    }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ATTB_REGS_COMPUTE));
    const int wg = warp >> 2;                      // which half of the keys / staging columns / (dK | dV) this warp owns
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const float sc = p.scale_log2, scale = p.scale;
    // dK_j, dV_j of a finished item -> 16-bit column blocks of dqkv; two warpgroups: one takes dK, the other dV; four:
    // (dK | dV) x (columns 0-31 | 32-63)
    auto drain_dkv = [&](int kv0, int h, int row_base) {
      const int key = kv0 + row;
      typename O16::T* dst = reinterpret_cast<typename O16::T*>(p.dqkv16) + long(row_base + key) * (3 * p.H * 64) + h * 64;
      const int which = ATTB_NWG == 2 ? wg : (wg >> 1);     // 0: dK -> column block 1, 1: dV -> column block 2
#pragma unroll 1
      for (int c = (ATTB_NWG == 2 ? 0 : (wg & 1)); c < (ATTB_NWG == 2 ? 2 : (wg & 1) + 1); ++c) {
        uint32_t v[32];
        tmem_ld32((which ? tdV : tdK) + lane_off + uint32_t(c * 32), v);
        tc_wait_ld();
        if (c == (ATTB_NWG == 2 ? 1 : (wg & 1))) {      // this warp's last read of the item's dK / dV accumulators
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dkv_free);
        }
        if (key < p.N) {
          // lane = key row: 256-bit stores write whole 32-byte sectors (16-byte ones left half-sector writes to L2)
          typename O16::T* d = dst + (which + 1) * p.H * 64 + c * 32;
#ifdef ATTB_STG128
#pragma unroll
          for (int k = 0; k < 32; k += 8)
            st_global_v4(d + k, O16::pack(__uint_as_float(v[k]), __uint_as_float(v[k + 1])),
                         O16::pack(__uint_as_float(v[k + 2]), __uint_as_float(v[k + 3])),
                         O16::pack(__uint_as_float(v[k + 4]), __uint_as_float(v[k + 5])),
                         O16::pack(__uint_as_float(v[k + 6]), __uint_as_float(v[k + 7])));
#else
#pragma unroll
          for (int k = 0; k < 32; k += 16) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = O16::pack(__uint_as_float(v[k + 2 * j]), __uint_as_float(v[k + 2 * j + 1]));
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d + k), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                         "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
          }
#endif
        }
      }
    };
#ifdef ATTB_DIAG
    unsigned dg_wait_s = 0, dg_math = 0, dg_wait_mma = 0, dg_drain = 0, dg_store = 0, dg_arrive = 0, dg_item = 0;
    const long long dg_start = clock64();
#define DG(var) do { const unsigned n__ = (unsigned)clock(); var += n__ - dg_t; dg_t = n__; } while (0)
#else
#define DG(var) do { } while (0)
#endif
    int g = 0, n = 0;
    int pv_kv0 = 0, pv_h = 0, pv_row = 0;          // the previous iteration's item and first query row (for the deferred read-outs)
    // this thread's row statistics (log-sum-exp, delta) are fetched one iteration ahead: loaded at the top of the iteration that
    // uses them, their global-load latency sat in front of the first exponential of every iteration
    float lse_n = 0.f, dlt_n = 0.f;
    auto fetch_stats = [&](int it, int i) {
      const int qrow = i * 128 + row;
      const bool ok = it < n_items && qrow < p.N;
      const long at = (long(item_b(ok ? it : 0)) * p.H + item_h(ok ? it : 0)) * p.N + qrow;
      lse_n = ok ? __ldg(p.lse + at) : 0.f;
      dlt_n = ok ? __ldg(p.delta + at) : 0.f;
    };
    fetch_stats(blockIdx.x, 0);
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
      const int kv0 = item_kv0(it), h = item_h(it), b = item_b(it);
      const int row_base = b * p.N;
      for (int i = 0; i < nq; ++i, ++g) {
#ifdef ATTB_DIAG
        unsigned dg_t = (unsigned)clock();
#endif
        const int qrow = i * 128 + row;
        const bool q_ok = qrow < p.N;
#ifdef ATTB_NO_STATS_PREFETCH
        fetch_stats(it, i);
        const float lse2 = lse_n, dlt = dlt_n;
#else
        const float lse2 = lse_n, dlt = dlt_n;
        if (i + 1 < nq) fetch_stats(it, i + 1);
        else fetch_stats(it + int(gridDim.x), 0);
#endif
        const bool full_tile = (kv0 + 128 <= p.N) && (i * 128 + 128 <= p.N);     // CTA-uniform
        mbar_wait(sdp_full, g & 1);
        tc_fence_after();
        DG(dg_wait_s);
#pragma unroll 1
        for (int c = ATTB_CPW * wg; c < ATTB_CPW * wg + ATTB_CPW; ++c) {
          uint32_t sv[32], dv[32];
          tmem_ld32(tS + lane_off + uint32_t(c * 32), sv);
          tmem_ld32(tdP + lane_off + uint32_t(c * 32), dv);
          tc_wait_ld();
          if (c == ATTB_CPW * wg + ATTB_CPW - 1) {       // this warp's last chunk of S_i / dP_i is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_free);
          }
          uint32_t pkP[16], pkD[16];
          // only the clip's last key tile and last query tile have rows / keys to mask: everywhere else the per-element
          // compare + select pairs (30 % of this kernel's instructions, r02c_prof_attention_bwd) are skipped
          if (full_tile) {
            // packed fp32x2: per PAIR of scores one FFMA2 (exponent), two MUFU, one FFMA2 (dP * scale - delta * scale), one FMUL2
            // and the two packs -- 3.5 issue slots per element instead of ~7
            const u64 sc2 = f2_packf(sc, sc), nlse2 = f2_packf(-lse2, -lse2);
            const u64 scale2 = f2_packf(scale, scale), ndlt2 = f2_packf(-dlt * scale, -dlt * scale);
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              float a0, a1, d0, d1, p0, p1;
              const u64 a = f2_fma(f2_pack(sv[k], sv[k + 1]), sc2, nlse2);
              if ((((k >> 1) * ATTB_NPOLY) & 7) < ATTB_NPOLY) {      // the polynomial pairs spread over each group of 8
                attb_ex2_poly2(a, p0, p1);
              } else {
                f2_unpack(a, a0, a1);
                p0 = ex2_approx(a0);
                p1 = ex2_approx(a1);
              }
              const u64 t = f2_fma(f2_pack(dv[k], dv[k + 1]), scale2, ndlt2);
              f2_unpack(f2_mul(f2_packf(p0, p1), t), d0, d1);
              pkP[k >> 1] = O16::pack(p0, p1);
              pkD[k >> 1] = O16::pack(d0, d1);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const int key = kv0 + c * 32 + k;
              float p0 = ex2_approx(fmaf(__uint_as_float(sv[k]), sc, -lse2));
              float p1 = ex2_approx(fmaf(__uint_as_float(sv[k + 1]), sc, -lse2));
              if (!q_ok || key >= p.N) p0 = 0.f;
              if (!q_ok || key + 1 >= p.N) p1 = 0.f;
              const float d0 = p0 * (__uint_as_float(dv[k]) - dlt) * scale;
              const float d1 = p1 * (__uint_as_float(dv[k + 1]) - dlt) * scale;
              pkP[k >> 1] = O16::pack(p0, p1);
              pkD[k >> 1] = O16::pack(d0, d1);
            }
          }
          DG(dg_math);
          // The previous iteration's GEMMs must retire before P/dS smem is overwritten (this iteration's first chunk of arithmetic
          // hides their latency; dQ leaves through the drain warpgroup).  An item boundary is no different -- the first scores of
          // the new item were issued during the old one -- except that the old item's dK / dV are read out here.
          if (c == ATTB_CPW * wg && g > 0) {
            mbar_wait(mma2_done, (g - 1) & 1);
            tc_fence_after();
            DG(dg_wait_mma);
            if (i == 0) {
              drain_dkv(pv_kv0, pv_h, pv_row - (nq - 1) * 128);
              DG(dg_item);
            }
          }
          uint8_t* bp = sP + (c >> 1) * ATT_TILE_BYTES + row * 128;
          uint8_t* bd = sdS + (c >> 1) * ATT_TILE_BYTES + row * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int chunk = (((c & 1) * 4 + q4) ^ (row & 7)) * 16;
            *reinterpret_cast<uint4*>(bp + chunk) = make_uint4(pkP[4 * q4], pkP[4 * q4 + 1], pkP[4 * q4 + 2], pkP[4 * q4 + 3]);
            *reinterpret_cast<uint4*>(bd + chunk) = make_uint4(pkD[4 * q4], pkD[4 * q4 + 1], pkD[4 * q4 + 2], pkD[4 * q4 + 3]);
          }
          DG(dg_store);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(pds_full);
        DG(dg_arrive);
        pv_kv0 = kv0; pv_h = h; pv_row = row_base + i * 128;
      }
    }
    if (g > 0) {      // the CTA's last item
      mbar_wait(mma2_done, (g - 1) & 1);
      tc_fence_after();
      drain_dkv(pv_kv0, pv_h, pv_row - (nq - 1) * 128);
    }
#ifdef ATTB_DIAG
    if (blockIdx.x == 1 && (threadIdx.x == 0 || threadIdx.x == 128) && g > 0)
      printf("ATTB_DIAG wg %d nq %d items %d per-iteration cycles: wait_s %u math %u wait_mma2 %u drain %u store %u fence+arrive %u | per item: end %u | total per iteration %lld\n",
             wg, nq, n, dg_wait_s / g, dg_math / g, dg_wait_mma / g, dg_drain / g, dg_store / g, dg_arrive / g, dg_item / n, (clock64() - dg_start) / g);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATTB_CWARPS) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mb
