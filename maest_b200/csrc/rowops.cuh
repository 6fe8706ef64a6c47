// Row-wise kernels around the GEMMs: LayerNorm (fp32 residual stream -> 16-bit GEMM operand),
// token mean, pooling + classification head.  All HBM-bound; one warp per 768-wide row, 16-byte accesses.
#pragma once
#include "common.cuh"

namespace mb {

constexpr int D_MODEL = 768;

// y16[r] = LN(x[r]; w, b, eps)     models/maest.py:395,405 (eps 1e-6, :499)
template <int DT>
__global__ void __launch_bounds__(256) layernorm_to16_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bta, void* __restrict__ y,
                                                             int rows, float eps, float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out) {
  using O = Op16<DT>;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + long(row) * D_MODEL);
  float4 v[6];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = warp_sum(s) * (1.0f / D_MODEL);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D_MODEL) + eps);
  if (mean_out && lane == 0) { mean_out[row] = mu; rstd_out[row] = rstd; }
  uint2* yr = reinterpret_cast<uint2*>(reinterpret_cast<typename O::T*>(y) + long(row) * D_MODEL);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
    const float4 be = __ldg(reinterpret_cast<const float4*>(bta) + lane + 32 * i);
    uint2 o;
    o.x = O::pack((v[i].x - mu) * rstd * g.x + be.x, (v[i].y - mu) * rstd * g.y + be.y);
    o.y = O::pack((v[i].z - mu) * rstd * g.z + be.z, (v[i].w - mu) * rstd * g.w + be.w);
    yr[lane + 32 * i] = o;
  }
}

// LayerNorm folding, weight side (done once per weight version): for the Linear that follows a LayerNorm(gamma, beta),
//   wg[n] = sum_k gamma[k] W16[n,k]      bf[n] = bias[n] + sum_k beta[k] W16[n,k]
// computed from the 16-bit operand copy the GEMM really multiplies.  One warp per output feature n.
template <int DT>
__global__ void __launch_bounds__(256) ln_fold_kernel(const void* __restrict__ w16, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ bias, int N, int K,
                                                      float* __restrict__ wg, float* __restrict__ bf) {
  using O = Op16<DT>;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const typename O::T* w = reinterpret_cast<const typename O::T*>(w16) + long(n) * K;
  float a = 0.f, b = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = O::to_f(w[k]);
    a = fmaf(gamma[k], wv, a);
    b = fmaf(beta[k], wv, b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { wg[n] = a; bf[n] = (bias ? bias[n] : 0.f) + b; }
}

// LayerNorm folding, statistics side: merge the partials (pivot, sum d, sum d^2 with d = x - pivot, number of features; 128
// features per partial) written by the producer GEMM epilogue into (rstd, -mean * rstd) per row (pairwise update of Chan et al., fixed order).
// Partials are [chunk][rows]: consecutive threads read consecutive 16-byte records.
__global__ void __launch_bounds__(256) ln_finalize_kernel(const float4* __restrict__ part, int rows, long row_pitch, int nparts, float eps,
                                                          float2* __restrict__ stats) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float mean = 0.f, m2 = 0.f, cnt = 0.f;
#pragma unroll 8
  for (int c = 0; c < nparts; ++c) {
    const float4 q = __ldg(part + long(c) * row_pitch + row);
    const float nc = q.w, inv = 1.0f / q.w;
    const float mc = q.x + q.y * inv;                          // partial mean
    const float m2c = fmaxf(q.z - q.y * q.y * inv, 0.f);       // partial sum of squared deviations from its mean
    const float tot = cnt + nc;
    const float d = mc - mean;
    mean += d * (nc / tot);
    m2 += m2c + d * d * (cnt * nc / tot);
    cnt = tot;
  }
  const float rstd = rsqrtf(m2 / cnt + eps);
  stats[row] = make_float2(rstd, -mean * rstd);
}

// emb[b, 0:768] = x[b,0], emb[b,768:1536] = x[b,1], emb[b,1536:2304] = mean(x[b,2:])   models/maest.py:825-829
// grid (B, 768/64), 256 threads = 4 row-groups x 64 columns
__global__ void __launch_bounds__(256) block_embedding_kernel(const float* __restrict__ x, int N, float* __restrict__ emb) {
  __shared__ float red[4][64];
  const int b = blockIdx.x, c = blockIdx.y * 64 + (threadIdx.x & 63), g = threadIdx.x >> 6;
  const float* xb = x + long(b) * N * D_MODEL;
  float s = 0.f;
  for (int r = 2 + g; r < N; r += 4) s += xb[long(r) * D_MODEL + c];
  red[g][threadIdx.x & 63] = s;
  __syncthreads();
  if (g == 0) {
    const float tot = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    float* e = emb + long(b) * 3 * D_MODEL;
    e[c] = xb[c];
    e[D_MODEL + c] = xb[D_MODEL + c];
    e[2 * D_MODEL + c] = tot / float(N - 2);
  }
}

struct HeadParams {
  const float* x; int N;      // [B, N, 768] residual stream after the last block
  const float* norm_w; const float* norm_b;        // final LN, eps 1e-6 (models/maest.py:553,806)
  const float* hln_w; const float* hln_b;          // head.0 LayerNorm, eps 1e-5 (:570-575)
  const float* head_w; const float* head_b;        // head.1 [C,768]
  const float* hdist_w; const float* hdist_b;      // head_dist [C,768] ("separated" only)
  int C; int separated;
  float* logits; float* logits_dist; float* feats; // [B,C], [B,C] (separated), [B,768]
  float* ln_cls; float* ln_dist;                   // optional [B,768] saves for backward
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

// one CTA (256 threads) per clip: final LN of rows 0/1 only, (cls+dist)/2, head LN + Linear  (models/maest.py:804-810, 905-925)
__global__ void __launch_bounds__(256) pool_head_kernel(const HeadParams p) {
  __shared__ float red[8];
  __shared__ float z[2][D_MODEL];   // LN'd cls / dist
  __shared__ float hz[D_MODEL];     // head-LN output (mean mode: of feats; separated: of cls)
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* xb = p.x + long(b) * p.N * D_MODEL;
  for (int which = 0; which < 2; ++which) {
    float v[3], s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) { v[i] = xb[which * D_MODEL + tid + 256 * i]; s += v[i]; }
    const float mu = block_sum_256(s, red) * (1.0f / D_MODEL);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) q += (v[i] - mu) * (v[i] - mu);
    const float rstd = rsqrtf(block_sum_256(q, red) * (1.0f / D_MODEL) + 1e-6f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      z[which][c] = (v[i] - mu) * rstd * p.norm_w[c] + p.norm_b[c];
    }
  }
  __syncthreads();
  {
    float v[3], s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      const float f = (z[0][c] + z[1][c]) * 0.5f;
      p.feats[long(b) * D_MODEL + c] = f;
      if (p.ln_cls) { p.ln_cls[long(b) * D_MODEL + c] = z[0][c]; p.ln_dist[long(b) * D_MODEL + c] = z[1][c]; }
      v[i] = p.separated ? z[0][c] : f;
      s += v[i];
    }
    const float mu = block_sum_256(s, red) * (1.0f / D_MODEL);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) q += (v[i] - mu) * (v[i] - mu);
    const float rstd = rsqrtf(block_sum_256(q, red) * (1.0f / D_MODEL) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      hz[c] = (v[i] - mu) * rstd * p.hln_w[c] + p.hln_b[c];
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int j = warp; j < p.C; j += 8) {
    const float4* wr = reinterpret_cast<const float4*>(p.head_w + long(j) * D_MODEL);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 w4 = __ldg(wr + lane + 32 * i);
      const float* h4 = hz + 4 * (lane + 32 * i);
      acc += w4.x * h4[0] + w4.y * h4[1] + w4.z * h4[2] + w4.w * h4[3];
    }
    acc = warp_sum(acc);
    if (lane == 0) p.logits[long(b) * p.C + j] = acc + p.head_b[j];
    if (p.separated) {
      const float4* wd = reinterpret_cast<const float4*>(p.hdist_w + long(j) * D_MODEL);
      float a2 = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float4 w4 = __ldg(wd + lane + 32 * i);
        const float* h4 = z[1] + 4 * (lane + 32 * i);
        a2 += w4.x * h4[0] + w4.y * h4[1] + w4.z * h4[2] + w4.w * h4[3];
      }
      a2 = warp_sum(a2);
      if (lane == 0) p.logits_dist[long(b) * p.C + j] = a2 + p.hdist_b[j];
    }
  }
}

// dst16 = cast(src32): fp32 master weights -> 16-bit GEMM operands
template <int DT>
__global__ void __launch_bounds__(256) cast_to16_kernel(const float* __restrict__ src, void* __restrict__ dst, long n) {
  using O = Op16<DT>;
  const long i = (long(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    uint2 o;
    o.x = O::pack(v.x, v.y);
    o.y = O::pack(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<typename O::T*>(dst) + i) = o;
  } else {
    for (long j = i; j < n; ++j) reinterpret_cast<typename O::T*>(dst)[j] = O::from_f(src[j]);
  }
}

}  // namespace mb
