// Attention forward, "three chains" kernel (round 2 default) for sm_100a: persistent, one CTA per SM, d_head = 64.
//
// Replaces models/maest.py:362-375 like attention.cuh; same inputs / outputs (packed qkv activation in, o16 and the optional
// log-sum-exp out).  What changed against attention_fwd_spec_kernel, and why (profiles/r01_attention_phase_clocks.md: that
// kernel was bound by the per-tile dependency chain S ready -> tcgen05.ld -> exp -> tcgen05.st -> PV with ONE softmax warp per
// sub-partition and CTA, 4.5 issue slots + one MUFU op per score):
//   * a work item is one 128-query tile of one (clip, head); a CTA walks items blockIdx.x, +gridDim.x, ... and its KV tiles form
//     ONE continuous stream g = 0, 1, 2, ...; tile g belongs to chain g % 3.  Each chain has its own 128-column score buffer
//     in TMEM and its own 4 softmax warps, so every sub-partition always has three softmax warps in different phases; the
//     next item's Q / K / V loads and first QK^T overlap the previous item's tail (no per-CTA prologue bubble).
//   * P is written IN PLACE over the scores it came from (16-bit pairs over the first 64 columns of the chain's buffer), the
//     tensor pipe runs PV(g) and then QK^T(g+3) into the same buffer in issue order; the O accumulator is double-buffered
//     across items (TMEM: 3 x 128 + 2 x 64 = 512 columns).
//   * the softmax never touches O: every tile of an item is exponentiated against ONE reference m_ref = the exact row max of
//     the item's first KV tile (published through shared memory by the chain that owns that tile).  No running max, no
//     rescaling, no per-tile wait on the previous PV.  P may exceed 1: fp16 P holds 2^a exactly up to a < 16 and fp32
//     accumulators do not care about the common scale.  Rows for which that is not enough (a 16-bit P overflowed to inf, the
//     row sum overflowed, or a polynomial-path exponent left [-126, 126]) are DETECTED in the epilogue (non-finite O or l)
//     and recomputed exactly by the epilogue warp on the CUDA cores (att_row_exact) -- a rare path, but it makes the kernel
//     correct for any input (tests: test_attention_sharp_scores_and_rescale_path).
//   * per PAIR of scores: one FFMA2 (scale and shift, packed fp32x2), two MUFU.EX2, one FADD2 (row sum), one F2FP; ATT_CHAIN_NPOLY
//     of every 8 pairs take the exponential on the FMA pipe instead (Cody-Waite + degree-3 minimax polynomial, relative error
//     8.0e-5, packed fp32x2 ops) because MUFU (16 / clk / SM) is the binding unit at d_head = 64.
// Warp roles (512 threads): warps 0-11 softmax (chain = warp / 4, TMEM lane quadrant = warp % 4), 12 TMA producer, 13 MMA
// issuer, 14-15 idle (they complete the fourth warpgroup so that setmaxnreg can move its registers to the softmax warps).  The epilogue of an item (O / l -> 16-bit, log-sum-exp, exact redo) is done by the chain that owns the SECOND tile
// after the item's last one, right after it has finished that tile: by then the item's last PV has retired, so the chain
// never waits for the tensor pipe, and no extra warps (and their registers) are needed.  Requires >= 2 KV tiles per item.
#pragma once
#include "attention.cuh"

namespace mb {

#ifndef ATT_CHAIN_NPOLY
#define ATT_CHAIN_NPOLY 3     // pairs of every 8 whose exponentials run on the FMA pipe
#endif

constexpr int ATC_THREADS = 512;   // 16 warps: the register file is handed out per warpgroup (setmaxnreg)
// 512 threads launch with 128 registers each (the whole file).  The producer / issuer warpgroup gives most of its share back
// (setmaxnreg.dec) and the three softmax warpgroups take it (setmaxnreg.inc): 3 x 128 x 152 + 128 x 56 = 65 536.  At 128 registers
// the softmax loop spilled its row-sum accumulators to local memory, and with ~200 KB of shared memory configured the L1 that
// backs local memory is only ~30 KB: every cold-path spill reload went to L2 (measured: a 150-instruction epilogue took 5000 cycles).
constexpr int ATC_REGS_SOFTMAX = 152;
constexpr int ATC_REGS_AUX = 56;
#ifndef ATC_R_N
#define ATC_R_N 5
#endif
constexpr int ATC_R = ATC_R_N;          // ring slots, each {K tile, V tile} = 32 KB
constexpr int ATC_NBAR = 2 + 2 + 2 * ATC_R + 3 + 3 + 2 + 2 + 2 + 2;
constexpr int ATC_OFF_KV = 2 * ATT_TILE_BYTES;
constexpr int ATC_OFF_BAR = ATC_OFF_KV + ATC_R * 2 * ATT_TILE_BYTES;
constexpr int ATC_OFF_MREF = ATC_OFF_BAR + ((ATC_NBAR * 8 + 16 + 127) / 128) * 128;
constexpr int ATC_OFF_LPART = ATC_OFF_MREF + 2 * 128 * 4;
constexpr int ATC_SMEM_BYTES = ATC_OFF_LPART + 2 * 3 * 128 * 4;

#ifdef ATC_DIAG   // timing diagnostic: per-role wait clocks, written over p.lse[blockIdx.x * 128 + ...] (results of lse are garbage)
#define ATC_T0() unsigned t__ = (unsigned)clock()
#define ATC_ACC(var) do { const unsigned n__ = (unsigned)clock(); var += n__ - t__; t__ = n__; } while (0)
#else
#define ATC_T0() do { } while (0)
#define ATC_ACC(var) do { } while (0)
#endif

typedef unsigned long long u64;
__device__ __forceinline__ u64 f2_pack(uint32_t lo, uint32_t hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ u64 f2_packf(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 f2_add(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 f2_sub(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// The values a tcgen05.ld wrote become visible at tcgen05.wait::ld; this empty asm makes every later use depend on a point
// AFTER the wait (the loads are software-pipelined, so there is real work between the ld and its wait).
__device__ __forceinline__ void reg_fence32(uint32_t (&r)[32]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
               "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
  asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
               "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}

__device__ __forceinline__ void tmem_st16_lo(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}

// S[tS] = Q K^T: four K = 16 steps in ONE asm block (descriptor advance 32 B = +2 per step), so the issuing lane moves five
// values to uniform registers per tile instead of re-deriving every operand per instruction.
__device__ __forceinline__ void mma_qk4(uint32_t tS, uint64_t qd, uint64_t kd, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p0, p1;\n\t.reg .b64 q1, q2, q3, k1, k2, k3;\n\t"
      "setp.ne.b32 p0, 0, 0;\n\tsetp.eq.b32 p1, 0, 0;\n\t"
      "add.s64 q1, %1, 2;\n\tadd.s64 q2, %1, 4;\n\tadd.s64 q3, %1, 6;\n\t"
      "add.s64 k1, %2, 2;\n\tadd.s64 k2, %2, 4;\n\tadd.s64 k3, %2, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], q1, k1, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], q2, k2, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], q3, k3, %3, p1;\n\t}\n" ::"r"(tS), "l"(qd), "l"(kd), "r"(idesc)
      : "memory");
}
// O[tO] (+)= P[tP] V: eight K = 16 steps (P advances 8 TMEM columns, V 16 rows = 2048 B = +128 per step); acc0 = 0 starts a new item.
__device__ __forceinline__ void mma_pv8(uint32_t tO, uint32_t tP, uint64_t vd, uint32_t idesc, uint32_t acc0) {
  asm volatile(
      "{\n\t.reg .pred p0, p1;\n\t.reg .b64 v1, v2, v3, v4, v5, v6, v7;\n\t.reg .b32 a1, a2, a3, a4, a5, a6, a7;\n\t"
      "setp.ne.b32 p0, %4, 0;\n\tsetp.eq.b32 p1, 0, 0;\n\t"
      "add.s64 v1, %2, 128;\n\tadd.s64 v2, %2, 256;\n\tadd.s64 v3, %2, 384;\n\tadd.s64 v4, %2, 512;\n\t"
      "add.s64 v5, %2, 640;\n\tadd.s64 v6, %2, 768;\n\tadd.s64 v7, %2, 896;\n\t"
      "add.s32 a1, %1, 8;\n\tadd.s32 a2, %1, 16;\n\tadd.s32 a3, %1, 24;\n\tadd.s32 a4, %1, 32;\n\t"
      "add.s32 a5, %1, 40;\n\tadd.s32 a6, %1, 48;\n\tadd.s32 a7, %1, 56;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], v1, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], v2, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], v3, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a4], v4, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a5], v5, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a6], v6, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [a7], v7, %3, p1;\n\t}\n" ::"r"(tO), "r"(tP), "l"(vd), "r"(idesc), "r"(acc0)
      : "memory");
}

// 32 score columns -> 16 packed P words, IN PLACE: P word i (columns 2i, 2i+1) replaces s[i], which pair i/2 has already consumed
// (saves 16 registers against a separate output array).  la / lb: packed fp32x2 partial row sums; amax: largest |a| seen on the
// polynomial path.
template <int DT, int NPOLY>
__device__ __forceinline__ void att_chain_chunk(uint32_t (&s)[32], u64& la, u64& lb, float& amax, const u64 sc2, const u64 negm2) {
  using O16 = Op16<DT>;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const u64 a = f2_fma(f2_pack(s[2 * i], s[2 * i + 1]), sc2, negm2);
    float p0, p1;
    if ((i & 7) >= 8 - NPOLY) {
      // 2^a = 2^round(a) * 2^r, r in [-0.5, 0.5]: round through the 1.5 * 2^23 magic add, minimax cubic for 2^r, the integer
      // part added into the exponent field (LEA).  Valid for |a| <= 126; amax lets the caller flag rows outside that range.
      float a0, a1;
      f2_unpack(a, a0, a1);
      amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));
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
    } else {
      float a0, a1;
      f2_unpack(a, a0, a1);
      p0 = ex2_approx(a0);
      p1 = ex2_approx(a1);
    }
    if (i & 1) lb = f2_add(lb, f2_packf(p0, p1));
    else la = f2_add(la, f2_packf(p0, p1));
    s[i] = O16::pack(p0, p1);
  }
}

// Exact attention for ONE query row on the CUDA cores (fp32 math on the 16-bit q / k / v, online softmax), executed by a whole
// warp: lane <-> key for the scores, lane <-> two output columns for P V.  Only reached for rows the fast path flagged.
template <int DT>
__device__ __noinline__ void att_row_exact(const AttnParams& p, const void* qkv_base, const int b, const int h, const int qrow,
                                           const int lane) {
  using O16 = Op16<DT>;
  using T = typename O16::T;
  const T* base = reinterpret_cast<const T*>(qkv_base) + size_t(b) * p.N * p.ld_qkv;
  const int hoff = h * ATT_D;
  uint32_t q2[32];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(base + size_t(qrow) * p.ld_qkv + hoff);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 v = __ldg(qp + i);
      q2[4 * i] = v.x; q2[4 * i + 1] = v.y; q2[4 * i + 2] = v.z; q2[4 * i + 3] = v.w;
    }
  }
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int k0 = 0; k0 < p.N; k0 += 32) {
    const int key = k0 + lane;
    float s = -INFINITY;
    if (key < p.N) {
      const uint4* kp = reinterpret_cast<const uint4*>(base + size_t(key) * p.ld_qkv + p.H * ATT_D + hoff);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 v = __ldg(kp + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 kf = O16::unpack(w[j]);
          const float2 qf = O16::unpack(q2[4 * i + j]);
          acc = fmaf(qf.x, kf.x, acc);
          acc = fmaf(qf.y, kf.y, acc);
        }
      }
      s = acc * p.scale_log2;
    }
    const float m_new = fmaxf(m, warp_max(s));
    const float f = exp2f(m - m_new);            // 0 on the first iteration (m = -inf)
    const float pr = exp2f(s - m_new);           // 0 for keys past the end
    l = l * f + warp_sum(pr);
    o0 *= f;
    o1 *= f;
    m = m_new;
    const int nk = min(32, p.N - k0);
    for (int kk = 0; kk < nk; ++kk) {
      const float pk = __shfl_sync(0xffffffffu, pr, kk);
      const uint32_t vv = __ldg(reinterpret_cast<const uint32_t*>(base + size_t(k0 + kk) * p.ld_qkv + 2 * p.H * ATT_D + hoff) + lane);
      const float2 vf = O16::unpack(vv);
      o0 = fmaf(pk, vf.x, o0);
      o1 = fmaf(pk, vf.y, o1);
    }
  }
  const float inv = 1.0f / l;
  uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<T*>(p.out) + size_t(b * p.N + qrow) * p.ld_out + hoff) + lane;
  *dst = O16::pack(o0 * inv, o1 * inv);
  if (p.lse != nullptr && lane == 0) p.lse[(size_t(b) * p.H + h) * p.N + qrow] = m + log2f(l);
}

// Coordinates (query tile, head, clip) of the items blockIdx.x, blockIdx.x + gridDim.x, ... without a division per item: the
// decomposition of gridDim.x is computed once, then each step is three adds with carries.
struct AtcItem {
  int qt, h, b;
};
struct AtcStep {
  int dqt, dh, db, nq, H;
  __device__ __forceinline__ void init(int grid, int nq_, int H_) {
    nq = nq_; H = H_;
    dqt = grid % nq;
    const int r = grid / nq;
    dh = r % H;
    db = r / H;
  }
  __device__ __forceinline__ AtcItem first(int it) const {
    AtcItem x;
    x.qt = it % nq;
    const int r = it / nq;
    x.h = r % H;
    x.b = r / H;
    return x;
  }
  __device__ __forceinline__ void next(AtcItem& x) const {
    x.qt += dqt;
    int carry = 0;
    if (x.qt >= nq) { x.qt -= nq; carry = 1; }
    x.h += dh + carry;
    carry = 0;
    if (x.h >= H) { x.h -= H; carry = 1; }
    x.b += db + carry;
  }
};

template <int DT>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_fwd_chain_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p, const void* qkv_base) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using O16 = Op16<DT>;
  uint8_t* sQ = smem;                       // [2]
  uint8_t* sKV = smem + ATC_OFF_KV;         // [ATC_R] slots of {K tile, V tile}: slot of "virtual tile" vg holds K(vg) and V(vg - 3)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATC_OFF_BAR);
  uint64_t* q_full = bars;                  // [2]  Q tile of item n landed
  uint64_t* q_empty = q_full + 2;           // [2]  last QK^T of the item retired
  uint64_t* kv_full = q_empty + 2;          // [R]  K(vg) and V(vg-3) landed
  uint64_t* kv_empty = kv_full + ATC_R;     // [R]  QK^T(vg) and PV(vg-3) retired
  uint64_t* s_full = kv_empty + ATC_R;      // [3]  scores of the chain's current tile are in TMEM
  uint64_t* p_full = s_full + 3;            // [3]  P written in place (4 warp arrivals)
  uint64_t* o_full = p_full + 3;            // [2]  last PV of the item retired
  uint64_t* o_empty = o_full + 2;           // [2]  epilogue has O, m_ref and the l partials of the item in registers (4 arrivals)
  uint64_t* mref_full = o_empty + 2;        // [2]  m_ref of the item published (4 arrivals)
  uint64_t* lpart_full = mref_full + 2;     // [2]  every chain has published its partial row sums of the item (12 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lpart_full + 2);
  float* s_mref = reinterpret_cast<float*>(smem + ATC_OFF_MREF);     // [2][128]
  float* s_lpart = reinterpret_cast<float*>(smem + ATC_OFF_LPART);   // [2][3][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nq = (p.N + ATT_BQ - 1) / ATT_BQ;
  const int nkv = (p.N + ATT_BKV - 1) / ATT_BKV;
  const int valid_last = p.N - (nkv - 1) * ATT_BKV;        // 1..128 real keys in the last KV tile
  const int nc_last = (valid_last + 31) & ~31;             // score columns computed for it
  const int n_items = p.B * p.H * nq;
  const int n_local = (n_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int n_tiles = n_local * nkv;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("attention: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
      mbar_init(&mref_full[i], 4);
      mbar_init(&lpart_full[i], 12);
    }
    for (int i = 0; i < ATC_R; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); }
    fence_mbar_init();
  }
  if (warp == 13) {
    if (lane == 0) tma_prefetch_desc(&tmap_qkv);
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: [0,384) three score / P buffers, [384,512) two O accumulators

  if (warp >= 12) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATC_REGS_AUX));
  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer (one lane)
    // Loads go out in the order the tensor pipe consumes them: K(0..2), then {V(g), K(g+3)} for g = 0, 1, ...; the pair shares a
    // ring slot and one mbarrier, so the issuer waits once per tile.
    if (lane == 0) {
      unsigned d_kw = 0, d_qw = 0;
      const long long d_start = clock64();
      (void)d_start;
      // K stream state: item of the next K tile and its coordinates (recomputed once per ITEM: no per-tile divisions)
      int kn = 0, kj = 0, k_row0 = 0, k_col = 0;
      auto k_item = [&]() {
        const int it = int(blockIdx.x) + kn * int(gridDim.x);
        const int qt = it % nq, bh = it / nq, h = bh % p.H, b = bh / p.H;
        k_row0 = b * p.N;
        k_col = h * ATT_D;
        ATC_T0();
        mbar_wait(&q_empty[kn & 1], ((kn >> 1) & 1) ^ 1);
        ATC_ACC(d_qw);
        mbar_expect_tx(&q_full[kn & 1], ATT_TILE_BYTES);
        tma_load_2d(sQ + (kn & 1) * ATT_TILE_BYTES, &tmap_qkv, &q_full[kn & 1], k_col, k_row0 + qt * ATT_BQ);
      };
      int vn = 0, vj = 0, v_row0 = 0, v_col = 0;      // V stream
      auto v_item = [&]() {
        const int it = int(blockIdx.x) + vn * int(gridDim.x);
        const int bh = it / nq, h = bh % p.H, b = bh / p.H;
        v_row0 = b * p.N;
        v_col = 2 * p.H * ATT_D + h * ATT_D;
      };
      int slot = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int vg = 0; vg < n_tiles + 3; ++vg) {
        const bool has_k = vg < n_tiles, has_v = vg >= 3;
        ATC_T0();
        mbar_wait(&kv_empty[slot], phase ^ 1);
        ATC_ACC(d_kw);
        mbar_expect_tx(&kv_full[slot], (has_k ? ATT_TILE_BYTES : 0) + (has_v ? ATT_TILE_BYTES : 0));
        uint8_t* dst = sKV + slot * (2 * ATT_TILE_BYTES);
        if (has_v) {
          if (vj == 0) v_item();
          tma_load_2d(dst + ATT_TILE_BYTES, &tmap_qkv, &kv_full[slot], v_col, v_row0 + vj * ATT_BKV);
          if (++vj == nkv) { vj = 0; ++vn; }
        }
        if (has_k) {
          if (kj == 0) k_item();
          tma_load_2d(dst, &tmap_qkv, &kv_full[slot], p.H * ATT_D + k_col, k_row0 + kj * ATT_BKV);
          if (++kj == nkv) { kj = 0; ++kn; }
        }
        if (++slot == ATC_R) { slot = 0; phase ^= 1; }
      }
#ifdef ATC_DIAG
      if (p.lse != nullptr) {
        float* d = p.lse + size_t(blockIdx.x) * 256 + 200;
        d[0] = float(d_kw); d[1] = 0.f; d[2] = float(d_qw); d[3] = float(clock64() - d_start);
      }
#endif
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer
    // The WHOLE warp runs this loop (uniform control flow: addresses, descriptors and counters stay in uniform registers, no
    // per-instruction R2UR); one elected lane issues the tcgen05 instructions.  Per tile g: wait P(g) and slot {V(g), K(g+3)},
    // issue PV(g) and QK^T(g+3), commit.
    constexpr uint32_t idesc_qk = make_idesc(DT, 128, 128, 0, 0);
    const uint32_t idesc_qk_last = make_idesc(DT, 128, nc_last, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc(DT, 128, 64, 0, 1);  // B = V, MN-major
    const uint64_t qdesc0 = make_sdesc(smem_u32(sQ), 16, 1024);
    const uint64_t kdesc0 = make_sdesc(smem_u32(sKV), 16, 1024);
    const uint64_t vdesc0 = make_sdesc(smem_u32(sKV + ATT_TILE_BYTES), 8192, 1024);
    unsigned d_q = 0, d_o = 0, d_p = 0, d_kv = 0;
    const long long d_start = clock64();
    (void)d_start;
    int qn = 0, qj = 0, qc = 0;        // next QK^T: item, tile in item, chain
    // QK^T of the next tile from the K half of `slot` into chain qc's buffer; caller has waited for the slot
    auto issue_qk = [&](const int slot) {
      const uint64_t qd = qdesc0 + uint64_t((qn & 1) * (ATT_TILE_BYTES >> 4));
      const uint64_t kd = kdesc0 + uint64_t(slot * (2 * ATT_TILE_BYTES >> 4));
      const uint32_t tS = tmem_base + uint32_t(qc) * 128u;
      const bool last = (qj == nkv - 1);
      const uint32_t idesc = last ? idesc_qk_last : idesc_qk;
      if (elect_one()) {
        mma_qk4(tS, qd, kd, idesc);
        tc_commit(&s_full[qc]);
        if (last) tc_commit(&q_empty[qn & 1]);
      }
      __syncwarp();
      if (++qc == 3) qc = 0;
      if (++qj == nkv) { qj = 0; ++qn; }
    };
    int slot = 0;
    uint32_t phase = 0;
    auto next_slot = [&]() { if (++slot == ATC_R) { slot = 0; phase ^= 1; } };
    // prologue: QK^T(0..2) (their slots hold a K tile only)
    for (int vg = 0; vg < 3 && vg < n_tiles; ++vg) {
      if (qj == 0) mbar_wait(&q_full[qn & 1], (qn >> 1) & 1);
      mbar_wait(&kv_full[slot], phase);
      tc_fence_after();
      issue_qk(slot);
      if (elect_one()) tc_commit(&kv_empty[slot]);
      __syncwarp();
      next_slot();
    }
    if (n_tiles < 3) { for (int vg = n_tiles; vg < 3; ++vg) next_slot(); }   // (unreachable: n_tiles >= 2 * n_local and nkv >= 2 ... kept for safety)
    int n = 0, j = 0, c = 0;
    uint32_t pbits = 0;                // per-chain phase parity of p_full
#pragma unroll 1
    for (int g = 0; g < n_tiles; ++g) {
      ATC_T0();
      if (j == 0) mbar_wait(&o_empty[n & 1], ((n >> 1) & 1) ^ 1);     // epilogue of item n-2 has drained this accumulator
      ATC_ACC(d_o);
      const bool more_qk = g + 3 < n_tiles;
      if (more_qk && qj == 0) mbar_wait(&q_full[qn & 1], (qn >> 1) & 1);
      ATC_ACC(d_q);
      mbar_wait(&p_full[c], (pbits >> c) & 1u);
      pbits ^= 1u << c;
      ATC_ACC(d_p);
      mbar_wait(&kv_full[slot], phase);
      ATC_ACC(d_kv);
      tc_fence_after();
      const uint32_t tP = tmem_base + uint32_t(c) * 128u;
      const uint32_t tO = tmem_base + 384u + uint32_t(n & 1) * 64u;
      const uint64_t vd = vdesc0 + uint64_t(slot * (2 * ATT_TILE_BYTES >> 4));
      const bool last = (j == nkv - 1);
      if (elect_one()) {
        if (!last) {
          mma_pv8(tO, tP, vd, idesc_pv, uint32_t(j));
        } else {
          const int ksteps = nc_last >> 4;
#pragma unroll 1
          for (int k = 0; k < ksteps; ++k) mma_ts(tO, tP + uint32_t(8 * k), vd + uint64_t(k * 128), idesc_pv, (j | k) ? 1u : 0u);
          tc_commit(&o_full[n & 1]);
        }
      }
      __syncwarp();
      if (more_qk) issue_qk(slot);      // QK^T(g+3) overwrites this chain's buffer; the tensor pipe executes in issue order
      if (elect_one()) tc_commit(&kv_empty[slot]);
      __syncwarp();
      next_slot();
      if (++c == 3) c = 0;
      if (++j == nkv) { j = 0; ++n; }
    }
#ifdef ATC_DIAG
    if (p.lse != nullptr && lane == 0) {
      float* d = p.lse + size_t(blockIdx.x) * 256 + 208;
      d[0] = float(d_q); d[1] = float(d_kv); d[2] = float(d_o); d[3] = float(d_p); d[4] = 0.f; d[5] = float(clock64() - d_start);
    }
#endif
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ATC_REGS_SOFTMAX));
    // ------------------------------------------------------------------ softmax chains: thread <-> query row (TMEM lane)
    const int c = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tS = tmem_base + uint32_t(c) * 128u + (uint32_t((warp & 3) * 32) << 16);
    const float sc = p.scale_log2;
    const u64 sc2 = f2_packf(sc, sc);
    int cur_n = -1, flushed = 0;
    unsigned d_s = 0, d_mref = 0, d_flush = 0, d_epi = 0, d_exp = 0, d_lead = 0, d_e1 = 0, d_e2 = 0, d_f1 = 0, d_bad = 0, d_e3 = 0, d_e4 = 0, d_e0 = 0;
    const long long d_start = clock64();
    u64 la = 0ull, lb = 0ull;
    float amax = 0.f;
    float m_ref = 0.f;
    auto flush_to = [&](int upto) {       // publish this chain's partial row sums of items [flushed, upto)
      while (flushed < upto) {
        const int m = flushed;
        float lt = 0.f;
        if (m == cur_n) {
          float x0, x1, y0, y1;
          f2_unpack(la, x0, x1);
          f2_unpack(lb, y0, y1);
          lt = (x0 + x1) + (y0 + y1);
          if (amax > 126.0f) lt = INFINITY;          // a polynomial-path exponent left its valid range: force the exact redo
        }
        ATC_T0();
        mbar_wait(&o_empty[m & 1], ((m >> 1) & 1) ^ 1);                // the epilogue of item m-2 has read its slots
        ATC_ACC(d_f1);
        s_lpart[(m & 1) * 384 + c * 128 + row] = lt;
        __syncwarp();
        if (lane == 0) mbar_arrive(&lpart_full[m & 1]);
        ++flushed;
      }
    };
    AtcStep step;
    step.init(int(gridDim.x), nq, p.H);
    AtcItem cur_it = step.first(int(blockIdx.x)), prev_it = cur_it;   // coordinates of items cur_n (0 before the first tile) / cur_n - 1
    int n_bad = 0;
    // item n (coordinates `x`): O / l -> 16-bit, log-sum-exp.  Rows the fast path could not represent get a NaN sentinel in
    // their first output word and are recomputed exactly after the main loop (keeps the function call and its register
    // traffic out of the pipelined part of the kernel).
    auto epilogue = [&](const int n, const AtcItem x) {
      ATC_T0();
      const int par = n & 1;
      const uint32_t ph = (n >> 1) & 1;
      ATC_ACC(d_e0);
      mbar_wait(&lpart_full[par], ph);
      ATC_ACC(d_e1);
      const float l = s_lpart[par * 384 + row] + s_lpart[par * 384 + 128 + row] + s_lpart[par * 384 + 256 + row];
      const float mr = s_mref[par * 128 + row];
      mbar_wait(&o_full[par], ph);
      ATC_ACC(d_e2);
      tc_fence_after();
      const uint32_t tO = tmem_base + 384u + uint32_t(par) * 64u + (uint32_t((warp & 3) * 32) << 16);
      const int qrow = x.qt * ATT_BQ + row;
      const float inv_l = 1.0f / l;
      bool good = (l > 0.f) && (l < INFINITY);
      typename O16::T* dst = reinterpret_cast<typename O16::T*>(p.out) + size_t(x.b * p.N + qrow) * p.ld_out + x.h * ATT_D;
      uint32_t v[32], w[32];
      tmem_ld32(tO, v);
      tmem_ld32(tO + 32u, w);
      tc_wait_ld();
      reg_fence32(v);
      reg_fence32(w);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[par]);       // O, m_ref and the l partials of this slot are in registers
      ATC_ACC(d_e3);
      // one column suffices for the finiteness test: an inf / NaN in a row of P reaches all 64 columns of that row of O
      good = good && (fabsf(__uint_as_float(v[0]) * inv_l) < INFINITY);
      if (qrow < p.N) {
        if (!good) {
          ++n_bad;                // the first output word of the row becomes the NaN sentinel 0x7fff7fff (both 16-bit halves)
        }
        if (p.lse != nullptr) p.lse[(size_t(x.b) * p.H + x.h) * p.N + qrow] = mr + log2f(l);
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          st_global_v4(dst + i, (i == 0 && !good) ? 0x7fff7fffu : O16::pack(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l),
                       O16::pack(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l),
                       O16::pack(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l),
                       O16::pack(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l));
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          st_global_v4(dst + 32 + i, O16::pack(__uint_as_float(w[i]) * inv_l, __uint_as_float(w[i + 1]) * inv_l),
                       O16::pack(__uint_as_float(w[i + 2]) * inv_l, __uint_as_float(w[i + 3]) * inv_l),
                       O16::pack(__uint_as_float(w[i + 4]) * inv_l, __uint_as_float(w[i + 5]) * inv_l),
                       O16::pack(__uint_as_float(w[i + 6]) * inv_l, __uint_as_float(w[i + 7]) * inv_l));
      }
      ATC_ACC(d_e4);
    };
    // Masked keys (past the end of the clip, last KV tile only) get the score that maps to a = -120: p = 2^-120 is zero for
    // every purpose (0 in fp16, 7.5e-37 in bf16 against row sums >= 1) and stays inside the polynomial path's valid range.
    const float inv_sc = 1.0f / sc;
    int n = 0, j = c;
    while (j >= nkv) { j -= nkv; ++n; }
    // three "virtual" tiles past the end give every chain one more pass through the item-change / epilogue logic below, so the
    // flush and the epilogue have exactly one call site each (code size: the whole kernel has to live in the instruction cache)
#pragma unroll 1
    for (int g = c; g < n_tiles + 3; g += 3) {
      bool new_item = false;
      ATC_T0();
      if (n != cur_n) {
        flush_to(n < n_local ? n : n_local);
        int k = cur_n < 0 ? 0 : cur_n;
        while (k < n) { prev_it = cur_it; step.next(cur_it); ++k; }    // cur_it = item n, prev_it = item n - 1
        cur_n = n;
        la = 0ull; lb = 0ull; amax = 0.f;
        new_item = true;
      }
      ATC_ACC(d_flush);
      if (g < n_tiles) {
        const bool last = (j == nkv - 1);
        const int nch = (last ? nc_last : ATT_BKV) >> 5;
        const int valid = last ? valid_last : ATT_BKV;
        mbar_wait(&s_full[c], ((g - c) / 3) & 1);
        ATC_ACC(d_s);
        tc_fence_after();
        uint32_t sa[32], sb[32];
        if (j == 0) {
          // this chain owns the item's first tile (never the ragged last one: nkv >= 2): exact row max -> the item's reference
          float mx = -INFINITY;
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            tmem_ld32(tS + uint32_t(ch * 32), sa);
            tc_wait_ld();
            reg_fence32(sa);
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              m0 = fmaxf(m0, __uint_as_float(sa[i]));
              m1 = fmaxf(m1, __uint_as_float(sa[i + 1]));
            }
            mx = fmaxf(mx, fmaxf(m0, m1));
          }
          m_ref = mx * sc;
          mbar_wait(&o_empty[n & 1], ((n >> 1) & 1) ^ 1);                // readers of the slot's previous item are done
          s_mref[(n & 1) * 128 + row] = m_ref;
          __syncwarp();
          if (lane == 0) mbar_arrive(&mref_full[n & 1]);
          ATC_ACC(d_lead);
        } else if (new_item) {
          // first tile of this chain in the item (tiles 1 or 2): pick the reference up
          mbar_wait(&mref_full[n & 1], (n >> 1) & 1);
          m_ref = s_mref[(n & 1) * 128 + row];
          ATC_ACC(d_mref);
        }
        const u64 negm2 = f2_packf(-m_ref, -m_ref);
        if (last && valid < (nch << 5)) {
          // ragged last tile: overwrite the score columns of keys past the end of the clip (in TMEM, one column per store)
          const uint32_t s_mask = __float_as_uint((m_ref - 120.0f) * inv_sc);
#pragma unroll 1
          for (int col = valid; col < (nch << 5); ++col) tmem_st1(tS + uint32_t(col), s_mask);
          tc_wait_st();
        }
        // chunks of 32 columns, two per iteration (registers ping-pong: the next chunk's tcgen05.ld is in flight while this one
        // is exponentiated); P is stored over score columns that have already been consumed
        tmem_ld32(tS, sa);
        tc_wait_ld();
        reg_fence32(sa);
#pragma unroll 1
        for (int ch = 0; ch < nch; ch += 2) {
          const bool has_b = ch + 1 < nch, has_a2 = ch + 2 < nch;
          if (has_b) tmem_ld32(tS + uint32_t((ch + 1) * 32), sb);
          att_chain_chunk<DT, ATT_CHAIN_NPOLY>(sa, la, lb, amax, sc2, negm2);
          tmem_st16_lo(tS + uint32_t(ch * 16), sa);
          if (has_b) {
            tc_wait_ld();
            reg_fence32(sb);
            if (has_a2) tmem_ld32(tS + uint32_t((ch + 2) * 32), sa);
            att_chain_chunk<DT, ATT_CHAIN_NPOLY>(sb, la, lb, amax, sc2, negm2);
            tmem_st16_lo(tS + uint32_t((ch + 1) * 16), sb);
            if (has_a2) {
              tc_wait_ld();
              reg_fence32(sa);
            }
          }
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[c]);
        ATC_ACC(d_exp);
      }
      // deferred epilogue of the previous item: its last tile was g - 2, i.e. this (real or virtual) tile is tile 1 of item n
      if (j == 1 && n >= 1) epilogue(n - 1, prev_it);
      ATC_ACC(d_epi);
      j += 3;
      while (j >= nkv) { j -= nkv; ++n; }
    }
    // exact redo of the rows flagged above (none in the common case: one ballot).  Every warp re-walks the items whose epilogue
    // it ran (tile 1 of item n + 1 belongs to chain (n * nkv + nkv + 1) % 3) and looks for the sentinel in its own rows.
    if (__any_sync(0xffffffffu, n_bad != 0)) {
      AtcItem x = step.first(int(blockIdx.x));
      for (int m = 0; m < n_local; ++m) {
        if (((m + 1) * nkv + 1) % 3 == c) {
          const int qrow = x.qt * ATT_BQ + row;
          bool flagged = false;
          if (qrow < p.N) {
            const uint32_t w0 = *reinterpret_cast<const volatile uint32_t*>(reinterpret_cast<typename O16::T*>(p.out) + size_t(x.b * p.N + qrow) * p.ld_out + x.h * ATT_D);
            flagged = (w0 == 0x7fff7fffu);
          }
          uint32_t bad = __ballot_sync(0xffffffffu, flagged);
          while (bad) {
            const int r = __ffs(bad) - 1;
            bad &= bad - 1;
            att_row_exact<DT>(p, qkv_base, x.b, x.h, x.qt * ATT_BQ + (warp & 3) * 32 + r, lane);
          }
        }
        step.next(x);
      }
    }
#ifdef ATC_DIAG
    if (p.lse != nullptr && lane == 0) {
      float* d = p.lse + size_t(blockIdx.x) * 256 + warp * 16;
      d[0] = float(d_s); d[1] = float(d_mref); d[2] = float(d_flush); d[3] = float(d_epi); d[4] = float(d_exp); d[5] = float(d_lead);
      d[6] = float(clock64() - d_start); d[7] = float(d_e1); d[8] = float(d_e2); d[9] = float(d_f1); d[10] = float(n_bad); d[11] = float(d_e0); d[12] = float(d_e3); d[13] = float(d_e4);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mb
