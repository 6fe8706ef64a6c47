// K2 helpers: patch gather (im2col for the 16x16 / stride-10 conv), position table, CLS/DIST rows.
//
// The patch-embed conv (models/maest.py:238-250) is run as the GEMM  [B*P, 256] x [768, 256]^T  on the
// tcgen05 kernel (gemm.cuh) whose epilogue adds  conv bias + freq_new_pos_embed[f] + time_new_pos_embed[off+t]
// (:645-675) from a pre-summed [P,768] table and scatters rows straight into the packed token buffer
// [B, 2+P, 768] (rows 0/1 = cls/dist + new_pos_embed, :785-796).  Structured / unstructured patchout
// (:678-780) is a list of kept (f,t) grid cells: only kept patches are gathered and multiplied.
#pragma once
#include "common.cuh"

namespace mb {

struct PatchGatherParams {
  const void* mel;       // [B, 96, T] fp32 or fp16
  int mel_is_half;
  int B, T, P;           // P kept patches per clip
  int Tp;                // full patch-grid width (T-16)/10+1 (for the identity mapping when keep_ft == null)
  const int32_t* keep_ft;  // [P] (f << 16 | t) or null -> p = f*Tp + t
  void* a16;             // [B*P, 256] 16-bit, K index = kh*16 + kw
  // cls/dist rows
  const float* cls_token; const float* dist_token; const float* new_pos_embed;  // [768], [768], [2,768]
  float* tokens;         // [B, 2+P, 768]
};

// one warp per patch: lane -> (kh = lane / 2, 8 consecutive kw); 16-byte stores, 32-byte-sector loads
template <int DT>
__global__ void __launch_bounds__(256) patch_gather_kernel(const PatchGatherParams p) {
  using O = Op16<DT>;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.B * p.P;
  if (warp_global < total) {
    const int b = warp_global / p.P, pi = warp_global - b * p.P;
    int f, t;
    if (p.keep_ft) { const int ft = p.keep_ft[pi]; f = ft >> 16; t = ft & 0xffff; }
    else { f = pi / p.Tp; t = pi - f * p.Tp; }
    const int kh = lane >> 1, kw0 = (lane & 1) * 8;
    const long src = (long(b) * 96 + 10 * f + kh) * p.T + 10 * t + kw0;
    float v[8];
    if (p.mel_is_half) {
      const __half* m = reinterpret_cast<const __half*>(p.mel) + src;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __half2float(m[i]);
    } else {
      const float* m = reinterpret_cast<const float*>(p.mel) + src;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(m + i);
    }
    typename O::T* dst = reinterpret_cast<typename O::T*>(p.a16) + long(warp_global) * 256 + kh * 16 + kw0;
    st_global_v4(dst, O::pack(v[0], v[1]), O::pack(v[2], v[3]), O::pack(v[4], v[5]), O::pack(v[6], v[7]));
  }
  // CLS / DIST rows: the first B*2 warps of the grid also write one token row each
  if (warp_global < 2 * p.B) {
    const int b = warp_global >> 1, which = warp_global & 1;
    const float* tok = which ? p.dist_token : p.cls_token;
    float* dst = p.tokens + (long(b) * (2 + p.P) + which) * 768;
    for (int c = lane; c < 768; c += 32) dst[c] = tok[c] + p.new_pos_embed[which * 768 + c];
  }
}

// pos[p][c] = conv_bias[c] + freq_pe[c][f_p] + time_pe[c][t_off + t_p]
struct PosTableParams {
  const float* conv_bias; const float* freq_pe; const float* time_pe;  // [768], [768,Fp], [768,Wt]
  int Fp, Wt, Tp, P, t_off;
  const int32_t* keep_ft;
  float* pos;  // [P,768]
};
__global__ void __launch_bounds__(256) pos_table_kernel(const PosTableParams p) {
  const int pi = blockIdx.x;
  int f, t;
  if (p.keep_ft) { const int ft = p.keep_ft[pi]; f = ft >> 16; t = ft & 0xffff; }
  else { f = pi / p.Tp; t = pi - f * p.Tp; }
  for (int c = threadIdx.x; c < 768; c += blockDim.x)
    p.pos[long(pi) * 768 + c] = p.conv_bias[c] + p.freq_pe[c * p.Fp + f] + p.time_pe[c * p.Wt + p.t_off + t];
}

}  // namespace mb
