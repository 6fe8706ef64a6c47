// Attention forward, "chains" kernel (round-2 default) for sm_100a: persistent, one CTA per SM, d_head = 64.
//
// Replaces models/maest.py:362-375 like attention.cuh; same inputs / outputs (packed qkv activation in, o16 and the optional
// log-sum-exp out).  What changed against attention_fwd_spec_kernel, and why (profiles/r01_attention_phase_clocks.md: that
// kernel was bound by the per-tile dependency chain S ready -> tcgen05.ld -> exp -> tcgen05.st -> PV with ONE softmax warp per
// sub-partition and CTA, 4.5 issue slots + one MUFU op per score):
//   * a work item is one 128-query tile of one (clip, head); a CTA walks items blockIdx.x, +gridDim.x, ... and its KV tiles form
//     ONE continuous stream g = 0, 1, 2, ...; tile g belongs to chain g % NCH.  Each chain has its own score buffer in TMEM and
//     its own 4 softmax warps, so every sub-partition always has NCH softmax warps in different phases; the next item's
//     Q / K / V loads and first QK^T overlap the previous item's tail (no per-CTA prologue bubble).
//   * P is written IN PLACE over the scores it came from (16-bit pairs over the first half of the chain's buffer); PV(g) reads
//     it there and QK^T(g + NCH) refills the buffer; the O accumulator is double-buffered across items
//     (TMEM: NCH x BKV + 2 x 64 = 512 columns for 3 x 128 and for 4 x 96).
//   * the softmax never touches O: every tile of an item is exponentiated against ONE reference m_ref = the exact row max of
//     the item's first KV tile (published through shared memory by the chain that owns that tile).  No running max, no
//     rescaling, no per-tile wait on the previous PV.  P may exceed 1: fp16 P holds 2^a exactly up to a < 16 and fp32
//     accumulators do not care about the common scale.  Rows for which that is not enough (a 16-bit P overflowed to inf, the
//     row sum overflowed, or a polynomial-path exponent left [-126, 126]) are DETECTED in the epilogue (non-finite O or l),
//     flagged with a NaN sentinel and recomputed exactly on the CUDA cores after the main loop (att_row_exact) -- a rare path,
//     but it makes the kernel correct for any input (tests: test_attention_sharp_scores_and_rescale_path).
//   * per PAIR of scores: one FFMA2 (scale and shift, packed fp32x2), two MUFU.EX2, one FADD2 (row sum), one F2FP; ATT_CHAIN_NPOLY
//     of every 8 pairs take the exponential on the FMA pipe instead (Cody-Waite + degree-3 minimax polynomial, relative error
//     8.0e-5, packed fp32x2 ops) because MUFU (16 / clk / SM) is the binding unit at d_head = 64.
// Warp roles: 4 x NCH softmax warps (chain = warp / 4, TMEM lane quadrant = warp % 4), then one warpgroup of helpers: TMA
// producer, PV issuer, QK^T issuer, one idle warp (completes the warpgroup so that setmaxnreg can move its registers to the
// softmax warps).  The epilogue of an item (O / l -> 16-bit, log-sum-exp) is done by the chain that owns the SECOND tile after
// the item's last one, right after it has finished that tile: by then the item's last PV has retired, so the chain never waits
// for the tensor pipe.  Requires >= 2 KV tiles per item (api.cu falls back to attention_fwd_spec_kernel below that).
#pragma once
#include "attention.cuh"

namespace mb {

#ifndef ATT_CHAIN_NPOLY
#define ATT_CHAIN_NPOLY 2     // pairs of every 8 whose exponentials run on the FMA pipe (re-tuned with the epilogue warpgroup:
                              // 1 / 2 / 3 / 4 -> 0.772 / 0.745-0.764 / 0.756-0.777 / 0.789 ms at config 3, box-dependent)
#endif

// NCH chains x BKV keys per tile; CW = score columns per tcgen05.ld chunk; R = ring slots of {K tile, V tile}.
// Registers: the CTA launches with the whole file split evenly, the helper warpgroup gives most of its share back
// (setmaxnreg.dec) and the softmax warpgroups take it (setmaxnreg.inc).  With spills the kernel was 30 % slower: the L1 that
// backs local memory is only ~30 KB when ~200 KB of shared memory are configured, so spill reloads went to L2.
// NSPLIT = 2: TWO softmax warps per (chain, TMEM lane quadrant), each owning half of the tile's score columns (its packed P stays
// inside its own half: PV reads k-steps 0..3 from the first half, 4..7 from the second).  Six instead of three softmax warps per
// sub-partition: the per-tile dependency chain S -> exp -> P of a chain is half as long and the sub-partition always has a warp
// to issue from (r02 phase clocks: with one warp per chain the softmax warps kept their sub-partition 42 % busy).
// NALT = 2: every softmax warp set owns TWO score buffers and alternates between them (tile g: set g % NCH, buffer g % (2 NCH)).
// While a set exponentiates the tile in one buffer, PV of its previous tile and QK^T of its next one run on the tensor pipe and
// refill the other buffer, so the set goes from tile to tile without the P -> PV -> QK^T -> S round trip (~1300 cycles per tile
// with one buffer per set, r02 phase clocks) -- at the price of half-size tiles (twice the per-tile bookkeeping).
// REGS_EPI > 0: a dedicated EPILOGUE warpgroup (4 more warps) turns every finished item's O / l into the 16-bit output.  With the
// epilogue inside the softmax chains, one chain disappeared for ~5000 cycles per item (r02 clocks: 370 cycles per tile and
// chain) and, because PV / QK^T are issued in tile order, the other two chains ran into its missing P a tile or two later.
// PHALF: the softmax warps also arrive on p_half once the packed P of the first half of the tile's keys is in TMEM (three
// quarters through the tile: the tcgen05.wait::st sits behind the third chunk's arithmetic, where it is free), and the PV issuer
// starts the first BKV/32 k-steps then.  PV is ISSUE-bound (one tcgen05.mma per ~53 cycles from one thread, r02_ubench_mma_issue),
// so half of its 426 issue cycles leave the P -> PV -> QK^T -> S round trip.
template <int NCH_, int BKV_, int CW_, int R_, int REGS_SM_, int REGS_AUX_, int NKV_MAX_, int NSPLIT_ = 1, int NALT_ = 1, int REGS_EPI_ = 0,
          bool PHALF_ = false>
struct AtcCfgT {
  static constexpr bool PHALF = PHALF_;
  static constexpr bool EPIWG = REGS_EPI_ > 0;
  static constexpr int REGS_EPI = REGS_EPI_;
  static constexpr int NCH = NCH_, BKV = BKV_, CW = CW_, R = R_, REGS_SM = REGS_SM_, REGS_AUX = REGS_AUX_, NSPLIT = NSPLIT_;
  static constexpr int NBUF = NCH * NALT_;                 // score buffers in TMEM
  // LEAN protocol (R == NBUF: ring slot == score buffer): the TMA producer reloads slot b when s_full[b] of the slot's previous
  // tile has completed -- QK^T(vg - NBUF) retired, and it was only issued after PV(vg - 2 NBUF) retired, so both halves of the slot
  // are free -- instead of waiting on kv_empty.  The issuers lose two commits per tile and walk the buffers in an unrolled loop
  // (buffer, slot, barrier and descriptor offsets are compile-time constants; lane 0 issues, no elect).
  static constexpr bool LEAN = R_ == NCH_ * NALT_;
  static constexpr int HW = BKV / NSPLIT;                  // score columns owned by one softmax warp
  static constexpr bool PINGPONG = NSPLIT == 1;            // two register chunks in flight (one warp per chain and quadrant needs the
                                                           // overlap; with split columns the other warps cover the tcgen05.ld latency)
  static constexpr int THREADS = (4 * NCH * NSPLIT + 4 + (EPIWG ? 4 : 0)) * 32;
  static constexpr int KV_BYTES = BKV * ATT_D * 2;
  static constexpr int SLOT_BYTES = 2 * KV_BYTES;
  static constexpr int NBAR = 2 + 2 + 2 * R + 4 * NBUF + 2 + 2 + 2 + 2;
  static constexpr int OFF_KV = 2 * ATT_TILE_BYTES;
  static constexpr int OFF_BAR = OFF_KV + R * SLOT_BYTES;
  static constexpr int OFF_MREF = OFF_BAR + ((NBAR * 8 + 16 + 127) / 128) * 128;
  static constexpr int OFF_LPART = OFF_MREF + 2 * NSPLIT * 128 * 4;
  static constexpr int NKV_MAX = NKV_MAX_;      // KV tiles per item the per-tile row-sum slots are sized for
  static constexpr int SMEM_BYTES = OFF_LPART + 2 * NKV_MAX * NSPLIT * 128 * 4;
  static_assert(NBUF * BKV + 128 <= 512, "TMEM: score buffers + two O accumulators");
  static_assert(BKV % 32 == 0 && (CW == 32 || CW == 16), "tile / chunk shape");
  static_assert(NSPLIT == 1 || (NSPLIT == 2 && HW % 32 == 0), "column split");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
  static_assert(!PHALF || (NSPLIT == 1 && CW == 32 && BKV == 128 && R_ != NCH_ * NALT_), "PHALF: generic ring, 4 chunks of 32 per tile");
};
using AtcCfg3 = AtcCfgT<3, 128, 32, 5, 152, 56, 24>;     // 512 threads: 3 x 128 x 152 + 128 x 56 = 65 536 registers
using AtcCfg4 = AtcCfgT<4, 96, 16, 6, 104, 56, 32>;      // 640 threads: 4 x 128 x 104 + 128 x 56 = 60 416 registers
#ifndef ATC_X2_REGS_SM
#define ATC_X2_REGS_SM 80
#define ATC_X2_REGS_AUX 24
#endif
#ifndef ATC_X2_CW
#define ATC_X2_CW 16
#endif
// (3 sets x 2 alternating buffers x 64 keys, AtcCfgT<3, 64, 32, 6, 152, 56, 32, 1, 2>, was measured at 0.90 - 1.01 ms -- issue-bound, see
// profiles/r02_attention_notes.md -- and is not instantiated: its fast build also showed an unresolved run-to-run race.)
using AtcCfg3E = AtcCfgT<3, 128, 32, 5, 128, 40, 24, 1, 1, 56>;   // 3 x 128 + epilogue warpgroup; 640 threads: 384 x 128 + 128 x 40 + 128 x 56 = 61 440 registers
using AtcCfg3EH = AtcCfgT<3, 128, 32, 5, 128, 40, 24, 1, 1, 56, true>;   // + PV starts on the first half of P
using AtcCfg3L = AtcCfgT<3, 128, 32, 3, 152, 56, 24>;      // 3 x 128, lean protocol
using AtcCfg3x2 = AtcCfgT<3, 128, ATC_X2_CW, 5, ATC_X2_REGS_SM, ATC_X2_REGS_AUX, 16, 2>;   // 896 threads: 6 x 128 x 80 + 128 x 24 = 64 512 registers

// ATC_LOADDIV = 2 (timing diagnostic only, results are garbage): every K / V load fetches only the first half of the tile's rows
// (the rest of the zero-initialised ring slot stays zero) -- tells whether the kernel is bound by L2 -> SM bandwidth.
#ifndef ATC_LOADDIV
#define ATC_LOADDIV 1
#endif
#ifdef ATC_DIAG   // timing diagnostic: per-role wait clocks, written over p.lse[blockIdx.x * 512 + ...] (the lse values are garbage)
#define ATC_T0() unsigned t__ = (unsigned)clock()
#define ATC_ACC(var) do { const unsigned n__ = (unsigned)clock(); var += n__ - t__; t__ = n__; } while (0)
#else
#define ATC_T0() do { } while (0)
#define ATC_ACC(var) do { } while (0)
#endif

// A chunk of W score columns of one TMEM lane in registers.  ld() is asynchronous: the values are defined after
// tcgen05.wait::ld + fence() (the empty asm makes every later use depend on a point AFTER the wait; the loads are
// software-pipelined, so there is real work between a ld and its wait).  st_lo() stores the first W/2 words (the packed P).
template <int W> struct AtcChunk;
template <> struct AtcChunk<32> {
  uint32_t r[32];
  __device__ __forceinline__ void ld(uint32_t taddr) { tmem_ld32(taddr, r); }
  __device__ __forceinline__ void fence() {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
    asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
  }
  __device__ __forceinline__ void st_lo(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
  }
};
template <> struct AtcChunk<16> {
  uint32_t r[16];
  __device__ __forceinline__ void ld(uint32_t taddr) { tmem_ld16(taddr, r); }
  __device__ __forceinline__ void fence() {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
  }
  __device__ __forceinline__ void st_lo(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  }
};
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}

// k-steps [K0, K1) of P V (PHALF: the two halves of a tile are issued at different times)
template <int K0, int K1, int KH>
__device__ __forceinline__ void mma_pv_range(uint32_t tO, uint32_t tP, uint64_t vd, uint32_t idesc, uint32_t acc0) {
#pragma unroll
  for (int k = K0; k < K1; ++k)
    mma_ts(tO, tP + uint32_t((k / KH) * (16 * KH) + 8 * (k % KH)), vd + uint64_t(128 * k), idesc, k == K0 ? acc0 : 1u);
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
// O[tO] (+)= P[tP] V: KS steps of K = 16 (P advances 8 TMEM columns, V 16 rows = 2048 B = +128 per step); acc0 = 0 starts an item.
// KH = k-steps per column half (KS when the tile is not split): the packed P of half h starts at column h * 16 * KH.
template <int KS, int KH>
__device__ __forceinline__ void mma_pv(uint32_t tO, uint32_t tP, uint64_t vd, uint32_t idesc, uint32_t acc0) {
  mma_ts(tO, tP, vd, idesc, acc0);
#pragma unroll
  for (int k = 1; k < KS; ++k) mma_ts(tO, tP + uint32_t((k / KH) * (16 * KH) + 8 * (k % KH)), vd + uint64_t(128 * k), idesc, 1u);
}

// W score columns -> W/2 packed P words, IN PLACE: P word i (columns 2i, 2i+1) replaces s[i], which pair i/2 has already consumed
// (no separate output array).  la / lb: packed fp32x2 partial row sums; amax: largest |a| seen on the polynomial path.
template <int DT, int NPOLY, int W>
__device__ __forceinline__ void att_chain_chunk(uint32_t (&s)[W], u64& la, u64& lb, float& amax, const u64 sc2, const u64 negm2) {
  using O16 = Op16<DT>;
#pragma unroll
  for (int i = 0; i < W / 2; ++i) {
    const u64 a = f2_fma(f2_pack(s[2 * i], s[2 * i + 1]), sc2, negm2);
    float p0, p1;
#ifdef ATC_NOEXP      // timing diagnostic: no exponentials at all (results are garbage)
    if (true) {
      float a0, a1;
      f2_unpack(a, a0, a1);
      p0 = fabsf(a0) * 1e-3f + 1e-3f;      // positive and finite: the row sums stay valid, no exact redo
      p1 = fabsf(a1) * 1e-3f + 1e-3f;
    } else
#endif
#ifdef ATT_CHAIN_POLY_SPREAD
    if (((i * NPOLY) & 7) < NPOLY) {      // the polynomial pairs spread over each group of 8
#else
    if ((i & 7) >= 8 - NPOLY) {
#endif
      // 2^a = 2^round(a) * 2^r, r in [-0.5, 0.5]: round through the 1.5 * 2^23 magic add, minimax cubic for 2^r, the integer
      // part added into the exponent field.  Valid for |a| <= 126; amax lets the caller flag rows outside that range.
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

template <int DT, typename Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
attention_fwd_chain_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, const AttnParams p,
                           const void* qkv_base) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using O16 = Op16<DT>;
  // NCH: score buffers (what the TMA / MMA-issuing warps step through); NSET: softmax warp sets (tile g -> set g % NSET)
  constexpr int NCH = Cfg::NBUF, NSET = Cfg::NCH, BKV = Cfg::BKV, CW = Cfg::CW, R = Cfg::R;
  constexpr int KV_BYTES = Cfg::KV_BYTES, SLOT_BYTES = Cfg::SLOT_BYTES;
  constexpr int NSPLIT = Cfg::NSPLIT, HW = Cfg::HW, KH = HW / 16;
  constexpr int SMW = 4 * NSET * NSPLIT;    // softmax warps
  uint8_t* sQ = smem;                       // [2]
  uint8_t* sKV = smem + Cfg::OFF_KV;        // [R] slots of {K tile, V tile}: the slot of "virtual tile" vg holds K(vg) and V(vg - NCH)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars;                  // [2]    Q tile of item n landed
  uint64_t* q_empty = q_full + 2;           // [2]    last QK^T of the item retired
  uint64_t* kv_full = q_empty + 2;          // [R]    K(vg) and V(vg - NCH) landed
  uint64_t* kv_empty = kv_full + R;         // [R]    QK^T(vg) and PV(vg - NCH) retired (two arrivals: one per issuing warp)
  uint64_t* s_full = kv_empty + R;          // [NCH]  scores of the chain's current tile are in TMEM
  uint64_t* p_full = s_full + NCH;          // [NCH]  P written in place (4 warp arrivals)
  uint64_t* pv_done = p_full + NCH;         // [NCH]  PV of the chain's tile retired: its buffer may take the next scores
  uint64_t* p_half = pv_done + NCH;         // [NCH]  (PHALF) P of the first half of the tile's keys written
  uint64_t* o_full = p_half + NCH;          // [2]    last PV of the item retired
  uint64_t* o_empty = o_full + 2;           // [2]    epilogue has O, m_ref and the l partials of the item in registers (4 arrivals)
  uint64_t* mref_full = o_empty + 2;        // [2]    m_ref of the item published (4 arrivals)
  uint64_t* lpart_full = mref_full + 2;     // [2]    the row sums of all tiles of the item are in shared memory (one arrival: PV issuer)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lpart_full + 2);
  float* s_mref = reinterpret_cast<float*>(smem + Cfg::OFF_MREF);     // [2][NSPLIT][128]
  float* s_lpart = reinterpret_cast<float*>(smem + Cfg::OFF_LPART);   // [2][NKV_MAX][NSPLIT][128]: row sum of (tile j, column half) of the item

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nq = (p.N + ATT_BQ - 1) / ATT_BQ;
  const int nkv = (p.N + BKV - 1) / BKV;
  const int valid_last = p.N - (nkv - 1) * BKV;            // 1..BKV real keys in the last KV tile
  const int nc_last = (valid_last + 31) & ~31;             // score columns computed for it
  const int n_items = p.B * p.H * nq;
  const int n_local = (n_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int n_tiles = n_local * nkv;

  if (ATC_LOADDIV > 1) {
    for (int i = threadIdx.x; i < R * SLOT_BYTES / 4; i += Cfg::THREADS) reinterpret_cast<uint32_t*>(sKV)[i] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("attention: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4 * NSPLIT);
      mbar_init(&mref_full[i], 4 * NSPLIT);
      mbar_init(&lpart_full[i], 1);
    }
    for (int i = 0; i < R; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2); }
    for (int i = 0; i < NCH; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4 * NSPLIT); mbar_init(&pv_done[i], 1); mbar_init(&p_half[i], 4 * NSPLIT); }
    fence_mbar_init();
  }
  if (warp == SMW + 1) {
    if (lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_kv); }
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: [0, NCH * BKV) score / P buffers, [384, 512) two O accumulators

  const int half = (warp >> 2) % NSPLIT;                 // softmax / epilogue warps: which HW-column part of a tile the warp owns
  const int row = (warp & 3) * 32 + lane;                // ... and their TMEM lane = query row of the tile
  int n_bad = 0;
  // coordinates of this CTA's item n, recomputed where they are needed (once per epilogue) instead of carried in registers
  auto item_of = [&](const int n) {
    const int it = int(blockIdx.x) + n * int(gridDim.x);
    AtcItem x;
    x.qt = it % nq;
    const int r = it / nq;
    x.h = r % p.H;
    x.b = r / p.H;
    return x;
  };
  // item n (coordinates `x`): O / l -> 16-bit, log-sum-exp.  Rows the fast path could not represent get a NaN sentinel in
  // their first output word and are recomputed exactly after the main loop (keeps the function call and its register
  // traffic out of the pipelined part of the kernel).
  auto epilogue = [&](const int n) {
    const AtcItem x = item_of(n);
    const int par = n & 1;
    const uint32_t ph = (n >> 1) & 1;
    mbar_wait(&lpart_full[par], ph);
    // the tiles' row sums are added in tile order, whichever chain produced them: the result does not depend on how this
    // CTA's items happened to line up with the chains (a clip computed alone or inside a batch gives identical bits)
    float l = 0.f;
    for (int k = 0; k < nkv * NSPLIT; ++k) l += s_lpart[(par * Cfg::NKV_MAX * NSPLIT + k) * 128 + row];
    float mr = s_mref[par * NSPLIT * 128 + row];
    if (NSPLIT > 1) mr = fmaxf(mr, s_mref[(par * NSPLIT + 1) * 128 + row]);
    mbar_wait(&o_full[par], ph);
    tc_fence_after();
    const uint32_t tO = tmem_base + 384u + uint32_t(par) * 64u + (uint32_t((warp & 3) * 32) << 16);
    const int qrow = x.qt * ATT_BQ + row;
    const float inv_l = 1.0f / l;
    bool good = (l > 0.f) && (l < INFINITY);
    typename O16::T* dst = reinterpret_cast<typename O16::T*>(p.out) + size_t(x.b * p.N + qrow) * p.ld_out + x.h * ATT_D;
    if (qrow < p.N && p.lse != nullptr) p.lse[(size_t(x.b) * p.H + x.h) * p.N + qrow] = mr + log2f(l);
    // the 64 columns of O in parts of 32: both when the warp owns whole rows, part `half` when the columns are split
#pragma unroll
    for (int q = 0; q < 2 / NSPLIT; ++q) {
      const int oh = half * (2 / NSPLIT) + q;
      uint32_t v[32];
      tmem_ld32(tO + uint32_t(oh * 32), v);
      tc_wait_ld();
      if (q == 2 / NSPLIT - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[par]);       // O, m_ref and the l partials of this slot are in registers
      }
      // one column suffices for the finiteness test: an inf / NaN in a row of P reaches all 64 columns of that row of O
      // (the warp that owns column 0 decides, sets the sentinel and later redoes the WHOLE row)
      if (oh == 0) {
        good = good && (fabsf(__uint_as_float(v[0]) * inv_l) < INFINITY);
#ifndef ATC_NOSOFTMAX
        if (!good && qrow < p.N) ++n_bad;   // the first output word of the row becomes the NaN sentinel 0x7fff7fff
#endif
      }
      if (qrow < p.N) {
#ifdef ATC_STG128
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          st_global_v4(dst + oh * 32 + i,
                       (oh == 0 && i == 0 && !good) ? 0x7fff7fffu : O16::pack(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l),
                       O16::pack(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l),
                       O16::pack(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l),
                       O16::pack(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l));
#else
        // lane = query row: 256-bit stores write whole 32-byte sectors (the backward's dK / dV read-out gained 4.7 % from this)
#pragma unroll
        for (int i = 0; i < 32; i += 16) {
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = O16::pack(__uint_as_float(v[i + 2 * j]) * inv_l, __uint_as_float(v[i + 2 * j + 1]) * inv_l);
          if (oh == 0 && i == 0 && !good) w[0] = 0x7fff7fffu;
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + oh * 32 + i), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                       "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
        }
#endif
      }
    }
  };
  // exact redo of the rows flagged by the epilogues this warp ran (none in the common case: one ballot).  Epilogue warpgroup: every
  // item; chains: tile epi_j of item m + 1 belongs to chain ((m + 1) * nkv + 1) % NSET, and with split columns the warp that
  // owns column 0 (the other half's warps may still be storing their part of a flagged row -- wait for all softmax warps).
  auto redo_flagged_rows = [&](const int c) {
    if (NSPLIT > 1) asm volatile("bar.sync 1, %0;" ::"n"(SMW * 32) : "memory");
    if (__any_sync(0xffffffffu, n_bad != 0)) {
      for (int m = 0; m < n_local; ++m) {
        if (Cfg::EPIWG || (((m + 1) * nkv + 1) % NSET == c && half == 0)) {
          const AtcItem x = item_of(m);
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
      }
    }
  };

  if (Cfg::EPIWG && warp >= SMW + 4) {
    // ------------------------------------------------------------------ epilogue warpgroup: item after item, off the chains' path
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::EPIWG ? Cfg::REGS_EPI : 24));
#pragma unroll 1
    for (int n = 0; n < n_local; ++n) epilogue(n);
    redo_flagged_rows(0);
  } else if (warp >= SMW) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_AUX));
  if (warp == SMW) {
    // ------------------------------------------------------------------ TMA producer
    // Loads go out in the order the tensor pipe consumes them: K(0..NCH-1), then {V(g), K(g+NCH)} for g = 0, 1, ...; the pair
    // shares a ring slot and one mbarrier.  The whole warp runs the loop (uniform control flow), one elected lane issues.
    unsigned d_kw = 0, d_qw = 0;
    const long long d_start = clock64();
    (void)d_start;
    AtcStep step;
    step.init(int(gridDim.x), nq, p.H);
    AtcItem kit = step.first(int(blockIdx.x)), vit = kit;      // items of the next K tile / next V tile
    int kn = 0, kj = 0, vj = 0;
    int slot = 0;
    uint32_t phase = 0;
    if (Cfg::LEAN) {
      uint32_t round = 0;
#pragma unroll 1
      for (int base = 0; base < n_tiles + NCH; base += NCH, ++round) {
#pragma unroll
        for (int b = 0; b < NCH; ++b) {
          const int vg = base + b;
          if (vg >= n_tiles + NCH) break;
          const bool has_k = vg < n_tiles, has_v = vg >= NCH;
          ATC_T0();
          if (has_k && kj == 0) mbar_wait(&q_empty[kn & 1], ((kn >> 1) & 1) ^ 1);
          ATC_ACC(d_qw);
          if (round > 0) mbar_wait(&s_full[b], (round - 1) & 1);         // QK^T(vg - NCH) retired => PV(vg - 2 NCH) retired
          ATC_ACC(d_kw);
          if (lane == 0) {
            uint8_t* dst = sKV + b * SLOT_BYTES;
            mbar_expect_tx(&kv_full[b], ((has_k ? KV_BYTES : 0) + (has_v ? KV_BYTES : 0)) / ATC_LOADDIV);
            if (has_k) {
              if (kj == 0) {
                mbar_expect_tx(&q_full[kn & 1], ATT_TILE_BYTES);
                tma_load_2d(sQ + (kn & 1) * ATT_TILE_BYTES, &tmap_q, &q_full[kn & 1], kit.h * ATT_D, kit.b * p.N + kit.qt * ATT_BQ);
              }
              tma_load_2d(dst, &tmap_kv, &kv_full[b], (p.H + kit.h) * ATT_D, kit.b * p.N + kj * BKV);
            }
            if (has_v) tma_load_2d(dst + KV_BYTES, &tmap_kv, &kv_full[b], (2 * p.H + vit.h) * ATT_D, vit.b * p.N + vj * BKV);
          }
          __syncwarp();
          if (has_v && ++vj == nkv) { vj = 0; step.next(vit); }
          if (has_k && ++kj == nkv) { kj = 0; ++kn; step.next(kit); }
        }
      }
    }
#pragma unroll 1
    for (int vg = 0; !Cfg::LEAN && vg < n_tiles + NCH; ++vg) {
      const bool has_k = vg < n_tiles, has_v = vg >= NCH;
      ATC_T0();
      if (has_k && kj == 0) mbar_wait(&q_empty[kn & 1], ((kn >> 1) & 1) ^ 1);
      ATC_ACC(d_qw);
      mbar_wait(&kv_empty[slot], phase ^ 1);
      ATC_ACC(d_kw);
      if (elect_one()) {
        uint8_t* dst = sKV + slot * SLOT_BYTES;
        mbar_expect_tx(&kv_full[slot], ((has_k ? KV_BYTES : 0) + (has_v ? KV_BYTES : 0)) / ATC_LOADDIV);
        if (has_v) tma_load_2d(dst + KV_BYTES, &tmap_kv, &kv_full[slot], (2 * p.H + vit.h) * ATT_D, vit.b * p.N + vj * BKV);
        if (has_k) {
          if (kj == 0) {
            mbar_expect_tx(&q_full[kn & 1], ATT_TILE_BYTES);
            tma_load_2d(sQ + (kn & 1) * ATT_TILE_BYTES, &tmap_q, &q_full[kn & 1], kit.h * ATT_D, kit.b * p.N + kit.qt * ATT_BQ);
          }
          tma_load_2d(dst, &tmap_kv, &kv_full[slot], (p.H + kit.h) * ATT_D, kit.b * p.N + kj * BKV);
        }
      }
      __syncwarp();
      if (has_v && ++vj == nkv) { vj = 0; step.next(vit); }
      if (has_k && ++kj == nkv) { kj = 0; ++kn; step.next(kit); }
      if (++slot == R) { slot = 0; phase ^= 1; }
    }
#ifdef ATC_DIAG
    if (p.lse != nullptr && lane == 0) {
      float* d = p.lse + size_t(blockIdx.x) * 512 + 400;
      d[0] = float(d_kw); d[1] = 0.f; d[2] = float(d_qw); d[3] = float(clock64() - d_start);
    }
#endif
  } else if (warp == SMW + 1) {
    // ------------------------------------------------------------------ PV issuer: O(item) += P(g) V(g)
    // Two issuing warps (this one and the QK^T issuer below) share the tensor pipe's queue: tcgen05.mma issue blocks while the
    // queue is full, and with one issuer the per-tile bookkeeping (barrier tests, descriptors, commits: ~600 cycles of a single
    // latency-bound thread) ran with the pipe idle.  Cross-warp order is by COMPLETION: QK^T(g+NCH) waits for pv_done of PV(g).
    constexpr uint32_t idesc_pv = make_idesc(DT, 128, 64, 0, 1);  // B = V, MN-major
    const uint64_t vdesc0 = make_sdesc(smem_u32(sKV + KV_BYTES), 8192, 1024);
    unsigned d_o = 0, d_p = 0, d_kv = 0, d_pv = 0, d_cm = 0, d_kv2 = 0;
    (void)d_kv2;
    const long long d_start = clock64();
    (void)d_start;
    int slot = NCH % R;
    uint32_t phase = (NCH / R) & 1;
    int n = 0, j = 0, c = 0;
    uint32_t pbits = 0;                // per-chain phase parity of p_full
    if (Cfg::LEAN) {
      uint32_t round = 0;
#pragma unroll 1
      for (int base = 0; base < n_tiles; base += NCH, ++round) {
#pragma unroll
        for (int b = 0; b < NCH; ++b) {
          if (base + b >= n_tiles) break;
          ATC_T0();
          if (j == 0) mbar_wait(&o_empty[n & 1], ((n >> 1) & 1) ^ 1);     // epilogue of item n-2 has drained this accumulator
          ATC_ACC(d_o);
          // P(g): completion `round` of p_full[b]; V(g) came with virtual tile g + NCH: completion round + 1 of kv_full[b]
          const bool ok_p = mbar_try_wait(&p_full[b], round & 1), ok_kv = mbar_try_wait(&kv_full[b], (round + 1) & 1);
          if (!ok_p) mbar_wait(&p_full[b], round & 1);
          ATC_ACC(d_p);
          if (!ok_kv) mbar_wait(&kv_full[b], (round + 1) & 1);
          ATC_ACC(d_kv);
          const bool last = (j == nkv - 1);
          if (last && lane == 0) mbar_arrive(&lpart_full[n & 1]);         // (see the generic path below)
          tc_fence_after();
          // (opaque copies: without them ptxas hoists all NCH x k-step operand values out of the tile loop and spills them)
          uint32_t tb = tmem_base;
          uint64_t vd0 = vdesc0;
          asm volatile("" : "+r"(tb), "+l"(vd0));
          const uint32_t tP = tb + uint32_t(b * BKV);
          const uint32_t tO = tb + 384u + uint32_t(n & 1) * 64u;
          const uint64_t vd = vd0 + uint64_t(b * (SLOT_BYTES >> 4));
          if (lane == 0) {
#ifdef ATC_DIAG
            const unsigned m0 = (unsigned)clock();
#endif
            if (!last) {
              mma_pv<BKV / 16, KH>(tO, tP, vd, idesc_pv, uint32_t(j));
            } else {
              const int ksteps = nc_last >> 4;
#pragma unroll 1
              for (int k = 0; k < ksteps; ++k)
                mma_ts(tO, tP + uint32_t((k / KH) * (16 * KH) + 8 * (k % KH)), vd + uint64_t(k * 128), idesc_pv, (j | k) ? 1u : 0u);
              tc_commit(&o_full[n & 1]);
            }
#ifdef ATC_DIAG
            const unsigned m1 = (unsigned)clock();
#endif
            tc_commit(&pv_done[b]);
#ifdef ATC_DIAG
            d_cm += m1 - m0;                        // "next" column: MMA issue only
            d_kv2 += (unsigned)clock() - m1;        // "-" column 7: the commit
#endif
          }
          __syncwarp();
          ATC_ACC(d_pv);
          if (++j == nkv) { j = 0; ++n; }
        }
      }
    }
#pragma unroll 1
    for (int g = 0; !Cfg::LEAN && g < n_tiles; ++g) {
      ATC_T0();
      if (j == 0) mbar_wait(&o_empty[n & 1], ((n >> 1) & 1) ^ 1);     // epilogue of item n-2 has drained this accumulator
      ATC_ACC(d_o);
      // the two per-tile barriers are tested together (a try_wait costs ~90 cycles even when the phase is long complete)
      const uint32_t ppar = (pbits >> c) & 1u;
      const bool early = Cfg::PHALF && j != nkv - 1;                    // full tile: start on the first half of P
      uint64_t* pbar = early ? &p_half[c] : &p_full[c];
      const bool ok_p = mbar_try_wait(pbar, ppar), ok_kv = mbar_try_wait(&kv_full[slot], phase);
      if (!ok_p) mbar_wait(pbar, ppar);
      ATC_ACC(d_p);
      if (!ok_kv) mbar_wait(&kv_full[slot], phase);
      ATC_ACC(d_kv);
      pbits ^= 1u << c;
      // Every softmax warp stores its tile's row sums BEFORE it arrives on p_full, and this warp has now acquired the p_full of
      // every tile of the item: one release-arrive here hands all of them to the epilogue (no per-chain flush protocol).
      if (j == nkv - 1 && lane == 0) mbar_arrive(&lpart_full[n & 1]);
      tc_fence_after();
      const uint32_t tP = tmem_base + uint32_t(c * BKV);
      const uint32_t tO = tmem_base + 384u + uint32_t(n & 1) * 64u;
      const uint64_t vd = vdesc0 + uint64_t(slot * (SLOT_BYTES >> 4));
      const bool last = (j == nkv - 1);
      if (early) {
        if (elect_one()) mma_pv_range<0, BKV / 32, KH>(tO, tP, vd, idesc_pv, uint32_t(j));
        __syncwarp();
        mbar_wait(&p_full[c], ppar);
        tc_fence_after();
      }
      if (elect_one()) {
        if (early) {
          mma_pv_range<BKV / 32, BKV / 16, KH>(tO, tP, vd, idesc_pv, 1u);
        } else if (!last) {
          mma_pv<BKV / 16, KH>(tO, tP, vd, idesc_pv, uint32_t(j));
        } else {
          const int ksteps = nc_last >> 4;
#pragma unroll 1
          for (int k = 0; k < ksteps; ++k)
            mma_ts(tO, tP + uint32_t((k / KH) * (16 * KH) + 8 * (k % KH)), vd + uint64_t(k * 128), idesc_pv, (j | k) ? 1u : 0u);
          tc_commit(&o_full[n & 1]);
        }
        tc_commit(&pv_done[c]);
        tc_commit(&kv_empty[slot]);
        if (g + NCH >= n_tiles) tc_commit(&kv_empty[slot]);     // no QK^T(g+NCH): this warp supplies the slot's second arrival too
      }
      __syncwarp();
      ATC_ACC(d_pv);
      if (++slot == R) { slot = 0; phase ^= 1; }
      if (++c == NCH) c = 0;
      if (++j == nkv) { j = 0; ++n; }
      ATC_ACC(d_cm);
    }
#ifdef ATC_DIAG
    if (p.lse != nullptr && lane == 0) {
      float* d = p.lse + size_t(blockIdx.x) * 512 + 408;
      d[0] = 0.f; d[1] = float(d_kv); d[2] = float(d_o); d[3] = float(d_p); d[4] = 0.f; d[5] = float(clock64() - d_start); d[6] = float(d_pv); d[7] = float(d_kv2); d[8] = float(d_cm);
    }
#endif
  } else if (warp == SMW + 2) {
    // ------------------------------------------------------------------ QK^T issuer: S(chain) = Q(item) K(vg)^T
    constexpr uint32_t idesc_qk = make_idesc(DT, 128, BKV, 0, 0);
    const uint32_t idesc_qk_last = make_idesc(DT, 128, nc_last, 0, 0);
    const uint64_t qdesc0 = make_sdesc(smem_u32(sQ), 16, 1024);
    const uint64_t kdesc0 = make_sdesc(smem_u32(sKV), 16, 1024);
    unsigned d_q = 0, d_kv = 0, d_pvd = 0, d_qk = 0;
    const long long d_start = clock64();
    (void)d_start;
    int slot = 0;
    uint32_t phase = 0;
    int qn = 0, qj = 0, qc = 0;        // item, tile in item, chain of the tile
    uint32_t dbits = 0;                // per-chain phase parity of pv_done
    if (Cfg::LEAN) {
      uint32_t round = 0;
#pragma unroll 1
      for (int base = 0; base < n_tiles; base += NCH, ++round) {
#pragma unroll
        for (int b = 0; b < NCH; ++b) {
          if (base + b >= n_tiles) break;
          ATC_T0();
          if (qj == 0) mbar_wait(&q_full[qn & 1], (qn >> 1) & 1);
          ATC_ACC(d_q);
          // buffer b still holds P(vg - NCH) until PV(vg - NCH) retires: completion round - 1 of pv_done[b]
          const bool ok_d = round > 0 ? mbar_try_wait(&pv_done[b], (round - 1) & 1) : true, ok_kv = mbar_try_wait(&kv_full[b], round & 1);
          if (!ok_d) mbar_wait(&pv_done[b], (round - 1) & 1);
          ATC_ACC(d_pvd);
          if (!ok_kv) mbar_wait(&kv_full[b], round & 1);
          ATC_ACC(d_kv);
          tc_fence_after();
          uint32_t tb = tmem_base;
          uint64_t kd0 = kdesc0;
          asm volatile("" : "+r"(tb), "+l"(kd0));
          const uint64_t qd = qdesc0 + uint64_t((qn & 1) * (ATT_TILE_BYTES >> 4));
          const uint64_t kd = kd0 + uint64_t(b * (SLOT_BYTES >> 4));
          const bool last = (qj == nkv - 1);
          if (lane == 0) {
            mma_qk4(tb + uint32_t(b * BKV), qd, kd, last ? idesc_qk_last : idesc_qk);
            tc_commit(&s_full[b]);
            if (last) tc_commit(&q_empty[qn & 1]);
          }
          __syncwarp();
          ATC_ACC(d_qk);
          if (++qj == nkv) { qj = 0; ++qn; }
        }
      }
    }
#pragma unroll 1
    for (int vg = 0; !Cfg::LEAN && vg < n_tiles; ++vg) {
      ATC_T0();
      if (qj == 0) mbar_wait(&q_full[qn & 1], (qn >> 1) & 1);
      ATC_ACC(d_q);
      const uint32_t dpar = (dbits >> qc) & 1u;
      const bool need_pv = vg >= NCH;                                 // the chain's buffer still holds P(vg - NCH) until PV(vg - NCH) retires
      const bool ok_d = need_pv ? mbar_try_wait(&pv_done[qc], dpar) : true, ok_kv = mbar_try_wait(&kv_full[slot], phase);
      if (!ok_d) mbar_wait(&pv_done[qc], dpar);
      ATC_ACC(d_pvd);
      if (!ok_kv) mbar_wait(&kv_full[slot], phase);
      ATC_ACC(d_kv);
      if (need_pv) dbits ^= 1u << qc;
      tc_fence_after();
      const uint64_t qd = qdesc0 + uint64_t((qn & 1) * (ATT_TILE_BYTES >> 4));
      const uint64_t kd = kdesc0 + uint64_t(slot * (SLOT_BYTES >> 4));
      const uint32_t tS = tmem_base + uint32_t(qc * BKV);
      const bool last = (qj == nkv - 1);
      if (elect_one()) {
        mma_qk4(tS, qd, kd, last ? idesc_qk_last : idesc_qk);
        tc_commit(&s_full[qc]);
        tc_commit(&kv_empty[slot]);
        if (vg < NCH) tc_commit(&kv_empty[slot]);                     // K-only slot of the prologue: no PV uses it
        if (last) tc_commit(&q_empty[qn & 1]);
      }
      __syncwarp();
      ATC_ACC(d_qk);
      if (++slot == R) { slot = 0; phase ^= 1; }
      if (++qc == NCH) qc = 0;
      if (++qj == nkv) { qj = 0; ++qn; }
    }
#ifdef ATC_DIAG
    if (p.lse != nullptr && lane == 0) {
      float* d = p.lse + size_t(blockIdx.x) * 512 + 420;
      d[0] = float(d_q); d[1] = float(d_kv); d[2] = float(d_pvd); d[3] = float(d_qk); d[4] = float(clock64() - d_start);
    }
#endif
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_SM));
    // ------------------------------------------------------------------ softmax chains: thread <-> query row (TMEM lane)
    const int c = warp / (4 * NSPLIT);
    const uint32_t tS0 = tmem_base + uint32_t(half * HW) + (uint32_t((warp & 3) * 32) << 16);    // + buffer * BKV
    const float sc = p.scale_log2;
    const u64 sc2 = f2_packf(sc, sc);
    int cur_n = -1;
    unsigned d_s = 0, d_mref = 0, d_flush = 0, d_epi = 0, d_exp = 0, d_lead = 0;
    const long long d_start = clock64();
    (void)d_start;
    float m_ref = 0.f;
    // Masked keys (past the end of the clip, last KV tile only) get the score that maps to a = -120: p = 2^-120 is zero for
    // every purpose (0 in fp16, 7.5e-37 in bf16 against row sums >= 1) and stays inside the polynomial path's valid range.
    const float inv_sc = 1.0f / sc;
    // The epilogue of item n - 1 runs after tile epi_j of item n: the same chain as tile 1 (whose predecessor's PV is the
    // item's last), one chain round later when the item is long enough -- by then that PV has long retired (r02 clocks: at tile 1
    // the epilogue still waited ~3000 cycles for o_full).
    const int epi_j = nkv > 1 + NSET ? 1 + NSET : 1;
    int n = 0, j = c;
    while (j >= nkv) { j -= nkv; ++n; }
    // NCH "virtual" tiles past the end give every chain one more pass through the item-change / epilogue logic below, so the
    // flush and the epilogue have exactly one call site each.
#pragma unroll 1
    int buf = c;                       // score buffer of tile g: g % NCH
    for (int g = c; g < n_tiles + 2 * NSET; g += NSET) {
      const uint32_t tS = tS0 + uint32_t(buf * BKV);
      bool new_item = false;
      ATC_T0();
      if (n != cur_n) {
        cur_n = n;
        new_item = true;
        // the row-sum and m_ref slots of this parity were last read by the epilogue of item n - 2
        if (n < n_local) mbar_wait(&o_empty[n & 1], ((n >> 1) & 1) ^ 1);
      }
      ATC_ACC(d_flush);
      if (g < n_tiles) {
        const bool last = (j == nkv - 1);
        // this warp's columns [half * HW, half * HW + ncols) of the tile, the first `valid` of them real keys
        const int ncols = max(0, min(HW, (last ? nc_last : BKV) - half * HW));
        const int valid = max(0, min(HW, (last ? valid_last : BKV) - half * HW));
#ifdef ATC_SLEEP_NS
        // the scores arrive a whole P -> PV -> QK^T round trip after this chain's last arrive: sleep through most of it instead
        // of polling (a failed try_wait + re-arm costs ~5 issue slots of the sub-partition the other chains are computing on)
        if (!mbar_try_wait(&s_full[buf], (g / NCH) & 1)) {
          __nanosleep(ATC_SLEEP_NS);
          mbar_wait(&s_full[buf], (g / NCH) & 1);
        }
#else
        mbar_wait(&s_full[buf], (g / NCH) & 1);
#endif
        ATC_ACC(d_s);
        tc_fence_after();
        AtcChunk<CW> ca, cb;
        u64 la = 0ull, lb = 0ull;      // packed fp32x2 partial row sums of this tile
        float amax = 0.f;
        if (j == 0) {
          // this chain owns the item's first tile (never the ragged last one: nkv >= 2): exact row max -> the item's reference
          float mx = -INFINITY;
#pragma unroll 1
          for (int col = 0; col < HW; col += CW) {
            ca.ld(tS + uint32_t(col));
            tc_wait_ld();
            ca.fence();
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < CW; i += 2) {
              m0 = fmaxf(m0, __uint_as_float(ca.r[i]));
              m1 = fmaxf(m1, __uint_as_float(ca.r[i + 1]));
            }
            mx = fmaxf(mx, fmaxf(m0, m1));
          }
          m_ref = mx * sc;
          s_mref[((n & 1) * NSPLIT + half) * 128 + row] = m_ref;
          __syncwarp();
          if (lane == 0) mbar_arrive(&mref_full[n & 1]);
          ATC_ACC(d_lead);
        }
        if (j == 0 ? NSPLIT > 1 : new_item) {
          // first tile of this chain in the item: pick the reference up (split columns: the owners of the first tile too -- the
          // reference is the max over both halves)
          mbar_wait(&mref_full[n & 1], (n >> 1) & 1);
          m_ref = s_mref[(n & 1) * NSPLIT * 128 + row];
          if (NSPLIT > 1) m_ref = fmaxf(m_ref, s_mref[((n & 1) * NSPLIT + 1) * 128 + row]);
          ATC_ACC(d_mref);
        }
        const u64 negm2 = f2_packf(-m_ref, -m_ref);
        if (last && valid < ncols) {
          // ragged last tile: overwrite the score columns of keys past the end of the clip (in TMEM, one column per store)
          const uint32_t s_mask = __float_as_uint((m_ref - 120.0f) * inv_sc);
#pragma unroll 1
          for (int col = valid; col < ncols; ++col) tmem_st1(tS + uint32_t(col), s_mask);
          tc_wait_st();
        }
        // chunks of CW columns, two per iteration (registers ping-pong: the next chunk's tcgen05.ld is in flight while this one
        // is exponentiated); P is stored over score columns that have already been consumed
#ifdef ATC_NOSOFTMAX     // timing diagnostic: no tcgen05.ld / exp / tcgen05.st at all -- the floor set by the MMA / barrier pipeline
        if (true) { la = f2_packf(1.f, 1.f); } else
#endif
        if (!Cfg::PINGPONG) {
#pragma unroll 1
          for (int col = 0; col < ncols; col += CW) {
            ca.ld(tS + uint32_t(col));
            tc_wait_ld();
            ca.fence();
            att_chain_chunk<DT, ATT_CHAIN_NPOLY, CW>(ca.r, la, lb, amax, sc2, negm2);
            ca.st_lo(tS + uint32_t(col >> 1));
          }
        } else if (ncols > 0) {
          ca.ld(tS);
          tc_wait_ld();
          ca.fence();
        }
#ifdef ATC_NOSOFTMAX
        if (false)
#endif
        bool half_sent = false;
#pragma unroll 1
        for (int col = 0; Cfg::PINGPONG && col < ncols; col += 2 * CW) {
          const bool has_b = col + CW < ncols, has_a2 = col + 2 * CW < ncols;
          if (has_b) cb.ld(tS + uint32_t(col + CW));
          att_chain_chunk<DT, ATT_CHAIN_NPOLY, CW>(ca.r, la, lb, amax, sc2, negm2);
          if (Cfg::PHALF && col == 2 * CW) {     // P of keys 0 .. 2 CW - 1 was stored a whole chunk ago: its wait::st is free here
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_half[buf]);
            half_sent = true;
          }
          ca.st_lo(tS + uint32_t(col >> 1));
          if (has_b) {
            tc_wait_ld();
            cb.fence();
            if (has_a2) ca.ld(tS + uint32_t(col + 2 * CW));
            att_chain_chunk<DT, ATT_CHAIN_NPOLY, CW>(cb.r, la, lb, amax, sc2, negm2);
            cb.st_lo(tS + uint32_t((col + CW) >> 1));
            if (has_a2) {
              tc_wait_ld();
              ca.fence();
            }
          }
        }
        {
          float x0, x1, y0, y1;
          f2_unpack(la, x0, x1);
          f2_unpack(lb, y0, y1);
          float lt = (x0 + x1) + (y0 + y1);
          if (amax > 126.0f) lt = INFINITY;          // a polynomial-path exponent left its valid range: force the exact redo
          s_lpart[(((n & 1) * Cfg::NKV_MAX + j) * NSPLIT + half) * 128 + row] = lt;
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (Cfg::PHALF && !half_sent && lane == 0) mbar_arrive(&p_half[buf]);   // short (last) tile: both at once
        if (lane == 0) mbar_arrive(&p_full[buf]);        // release: the row sums above and P in TMEM
        ATC_ACC(d_exp);
      }
      // deferred epilogue of the previous item: its last tile was g - 2, i.e. this (real or virtual) tile is tile 1 of item n
      if (!Cfg::EPIWG && j == epi_j && n >= 1 && n <= n_local) epilogue(n - 1);  // (n > n_local: a virtual tile past the virtual item)
      ATC_ACC(d_epi);
      j += NSET;
      buf += NSET;
      if (buf >= NCH) buf -= NCH;
      while (j >= nkv) { j -= nkv; ++n; }
    }
    if (!Cfg::EPIWG) redo_flagged_rows(c);
#ifdef ATC_DIAG
    if (p.lse != nullptr && lane == 0) {
      float* d = p.lse + size_t(blockIdx.x) * 512 + warp * 16;
      d[0] = float(d_s); d[1] = float(d_mref); d[2] = float(d_flush); d[3] = float(d_epi); d[4] = float(d_exp); d[5] = float(d_lead);
      d[6] = float(clock64() - d_start); d[10] = float(n_bad);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SMW + 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mb
