// Row-wise / reduction kernels of the training step (Module.training_step, models/module.py:73-102 and the autograd
// mirrors of the forward path).  All HBM-bound.  Gradients are ACCUMULATED into caller-provided fp32 buffers.
#pragma once
#include "common.cuh"
#include "rowops.cuh"

namespace mb {

// ---------------------------------------------------------------------------------------------------------------
// mixup (models/module.py:77-86): out[b,i] = x[b,i]*lam[b] + x[perm[b],i]*(1-lam[b]); fp16 or fp32 in, fp32 out
// ---------------------------------------------------------------------------------------------------------------
template <typename TIN>
__global__ void __launch_bounds__(256) mixup_kernel(const TIN* __restrict__ x, const int32_t* __restrict__ perm,
                                                    const float* __restrict__ lam, float* __restrict__ out, int B, long L) {
  const int b = blockIdx.y;
  const float l = lam[b];
  const TIN* xa = x + long(b) * L;
  const TIN* xb = x + long(perm[b]) * L;
  float* o = out + long(b) * L;
  for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < L; i += long(gridDim.x) * blockDim.x)
    o[i] = float(xa[i]) * l + float(xb[i]) * (1.0f - l);
}

// ---------------------------------------------------------------------------------------------------------------
// BCE-with-logits, mean reduction (models/module.py:90).  One CTA.  Also emits dlogits = (sigmoid(z) - y) / n.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bce_logits_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                          float* __restrict__ loss, float* __restrict__ dz, int n) {
  __shared__ float red[32];
  float acc = 0.f;
  const float inv_n = 1.0f / float(n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float zi = z[i], yi = y[i];
    const float e = expf(-fabsf(zi));
    acc += fmaxf(zi, 0.f) - zi * yi + log1pf(e);
    const float sig = zi >= 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);
    dz[i] = (sig - yi) * inv_n;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *loss = v * inv_n;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// head backward, part 1: one CTA (256 threads) per clip.  Autograd of models/maest.py:806-810 (final LN, rows 0/1),
// :906 ((cls+dist)/2) and :909 (head = LayerNorm(eps 1e-5) + Linear), "mean" distillation mode; with `separated` set,
// :914-925: logits = head(cls), logits_dist = head_dist(dist) (head_dist is a bare Linear), two independent gradients.
// ---------------------------------------------------------------------------------------------------------------
struct HeadBwdParams {
  const float* x; int N;                 // [B, N, 768] residual stream after the last block (saved)
  const float* dlogits; const float* gscale;   // [B, C] from bce_logits_kernel; *gscale = upstream d(loss)
  const float* norm_w; const float* norm_b; const float* hln_w; const float* hln_b; const float* head_w;  // params
  int C;
  float* dx;                             // [B, N, 768] gradient stream: rows 0/1 of each clip are WRITTEN here
  float* hz;                             // [B, 768] scratch: head-LN output, consumed by head_wgrad_kernel
  float* d_norm_w; float* d_norm_b; float* d_hln_w; float* d_hln_b;   // accumulated (atomics)
  int separated;                         // 0: "mean" mode; 1: "separated" (the three fields below are used)
  const float* dlogits_dist;             // [B, C] gradient of the head_dist logits
  const float* hdist_w;                  // [C, 768]
  float* z1out;                          // [B, 768] scratch: final-LN'd dist row, consumed by head_wgrad_kernel for head_dist
};

__global__ void __launch_bounds__(256) head_bwd_clip_kernel(const HeadBwdParams p) {
  __shared__ float red[8];
  __shared__ float dl[1024];
  __shared__ float dl2[1024];        // separated mode: d(logits_dist)
  __shared__ float xh[2][D_MODEL];   // normalised (pre-affine) rows 0/1
  __shared__ float rs[2];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float gs = *p.gscale;
  for (int j = tid; j < p.C; j += 256) {
    dl[j] = p.dlogits[long(b) * p.C + j] * gs;
    if (p.separated) dl2[j] = p.dlogits_dist[long(b) * p.C + j] * gs;
  }
  const float* xb = p.x + long(b) * p.N * D_MODEL;
  // recompute the final LN of rows 0/1
  float z[2][3];
  for (int r = 0; r < 2; ++r) {
    float v[3], s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) { v[i] = xb[r * D_MODEL + tid + 256 * i]; s += v[i]; }
    const float mu = block_sum_256(s, red) * (1.0f / D_MODEL);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) q += (v[i] - mu) * (v[i] - mu);
    const float rstd = rsqrtf(block_sum_256(q, red) * (1.0f / D_MODEL) + 1e-6f);
    if (tid == 0) rs[r] = rstd;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      const float h = (v[i] - mu) * rstd;
      xh[r][c] = h;
      z[r][i] = h * p.norm_w[c] + p.norm_b[c];
    }
  }
  // feats, head LN forward (recomputed), d(hz) = dlogits . W
  float f[3], fh[3], s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) { f[i] = p.separated ? z[0][i] : 0.5f * (z[0][i] + z[1][i]); s += f[i]; }
  const float fmu = block_sum_256(s, red) * (1.0f / D_MODEL);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) q += (f[i] - fmu) * (f[i] - fmu);
  const float frstd = rsqrtf(block_sum_256(q, red) * (1.0f / D_MODEL) + 1e-5f);
  float dhz[3] = {0.f, 0.f, 0.f};
  float dz1[3] = {0.f, 0.f, 0.f};    // separated mode: d loss / d (final-LN'd dist row) = dlogits_dist . W_dist
  __syncthreads();   // dl[] visible
  for (int j = 0; j < p.C; ++j) {
    const float d = dl[j];
    const float* w = p.head_w + long(j) * D_MODEL;
#pragma unroll
    for (int i = 0; i < 3; ++i) dhz[i] = fmaf(d, __ldg(w + tid + 256 * i), dhz[i]);
  }
  if (p.separated) {
    for (int j = 0; j < p.C; ++j) {
      const float d = dl2[j];
      const float* w = p.hdist_w + long(j) * D_MODEL;
#pragma unroll
      for (int i = 0; i < 3; ++i) dz1[i] = fmaf(d, __ldg(w + tid + 256 * i), dz1[i]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) p.z1out[long(b) * D_MODEL + tid + 256 * i] = z[1][i];
  }
  float g[3], sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = tid + 256 * i;
    fh[i] = (f[i] - fmu) * frstd;
    p.hz[long(b) * D_MODEL + c] = fh[i] * p.hln_w[c] + p.hln_b[c];
    atomicAdd(p.d_hln_w + c, dhz[i] * fh[i]);
    atomicAdd(p.d_hln_b + c, dhz[i]);
    g[i] = dhz[i] * p.hln_w[c];
    sg += g[i];
    sgx += g[i] * fh[i];
  }
  const float mg = block_sum_256(sg, red) * (1.0f / D_MODEL);
  const float mgx = block_sum_256(sgx, red) * (1.0f / D_MODEL);
  float dfe[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) dfe[i] = frstd * (g[i] - mg - fh[i] * mgx);   // d loss / d feats
  // rows 0/1 through the final LN: dy = dfeats / 2 ("mean") or (d cls, d dist) ("separated")
  for (int r = 0; r < 2; ++r) {
    float gg[3], a = 0.f, ax = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      const float dy = p.separated ? (r == 0 ? dfe[i] : dz1[i]) : 0.5f * dfe[i];
      atomicAdd(p.d_norm_w + c, dy * xh[r][c]);
      atomicAdd(p.d_norm_b + c, dy);
      gg[i] = dy * p.norm_w[c];
      a += gg[i];
      ax += gg[i] * xh[r][c];
    }
    const float m1 = block_sum_256(a, red) * (1.0f / D_MODEL);
    const float m2 = block_sum_256(ax, red) * (1.0f / D_MODEL);
    const float rstd = rs[r];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      p.dx[(long(b) * p.N + r) * D_MODEL + c] = rstd * (gg[i] - m1 - xh[r][c] * m2);
    }
  }
}

// head backward, part 2: d head.1.weight[j,:] += sum_b dlogits[b,j] hz[b,:],  d head.1.bias[j] += sum_b dlogits[b,j]
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ gscale,
                                                         const float* __restrict__ hz, int B, int C, float* __restrict__ dW,
                                                         float* __restrict__ db) {
  const int j = blockIdx.x, tid = threadIdx.x;
  const float gs = *gscale;
  float acc[3] = {0.f, 0.f, 0.f}, sb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = dlogits[long(b) * C + j] * gs;
    sb += d;
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] = fmaf(d, hz[long(b) * D_MODEL + tid + 256 * i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) dW[long(j) * D_MODEL + tid + 256 * i] += acc[i];
  if (tid == 0) db[j] += sb;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward (autograd of norm1/norm2, models/maest.py:418-419): warp per row, strided over rows.
//   dx[r] += rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;   dgamma += sum dy*xhat, dbeta += sum dy
// Optionally also writes the updated dx as a 16-bit GEMM operand (dx16) for the next backward GEMMs, and accumulates the column
// sums of the UPDATED dx into dx_colsum: that is the bias gradient of the linear layer whose output gradient this dx is (fc2 of
// the previous block after norm1, proj of this block after norm2) -- it replaces a separate colsum pass over dx (331 MB read).
// All three operands of a row (x, dy, dx) are requested before the first reduction, so a warp has 18 float4 loads in flight.
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, float* __restrict__ dx,
                                                            void* __restrict__ dx16, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, float* __restrict__ dx_colsum, int rows) {
  using O = Op16<DT>;
  __shared__ float sred[8][D_MODEL];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * 8;
  float4 gam[6], ag[6], ab[6], ao[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    gam[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ao[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += warps_total) {
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + long(row) * D_MODEL);
    const float4* dyr = reinterpret_cast<const float4*>(dy + long(row) * D_MODEL);
    float4* dxr = reinterpret_cast<float4*>(dx + long(row) * D_MODEL);
    float4 xh[6], g[6], dxv[6];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) dxv[i] = dxr[lane + 32 * i];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 xv = xr[lane + 32 * i], d = dyr[lane + 32 * i];
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[i] = make_float4(d.x * gam[i].x, d.y * gam[i].y, d.z * gam[i].z, d.w * gam[i].w);
      ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    const float m1 = warp_sum(s1) * (1.0f / D_MODEL), m2 = warp_sum(s2) * (1.0f / D_MODEL);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float4 o = dxv[i];
      o.x += rs * (g[i].x - m1 - xh[i].x * m2);
      o.y += rs * (g[i].y - m1 - xh[i].y * m2);
      o.z += rs * (g[i].z - m1 - xh[i].z * m2);
      o.w += rs * (g[i].w - m1 - xh[i].w * m2);
      dxr[lane + 32 * i] = o;
      ao[i].x += o.x; ao[i].y += o.y; ao[i].z += o.z; ao[i].w += o.w;
      if (dx16 != nullptr) {
        uint2 pk;
        pk.x = O::pack(o.x, o.y);
        pk.y = O::pack(o.z, o.w);
        reinterpret_cast<uint2*>(reinterpret_cast<typename O::T*>(dx16) + long(row) * D_MODEL)[lane + 32 * i] = pk;
      }
    }
  }
  // CTA-level reduction of dgamma / dbeta partials, then one atomic per column per CTA
  for (int pass = 0; pass < (dx_colsum != nullptr ? 3 : 2); ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 v = pass == 0 ? ag[i] : pass == 1 ? ab[i] : ao[i];
      *reinterpret_cast<float4*>(&sred[warp][4 * (lane + 32 * i)]) = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D_MODEL; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sred[w][c];
      atomicAdd((pass == 0 ? dgamma : pass == 1 ? dbeta : dx_colsum) + c, t);
    }
  }
}

// out[n] += sum_m in[m, n]   (bias gradients).  CTA = 256 threads = 8 row-lanes x 32 column-groups of 8 columns (16-byte /
// 32-byte vector loads, 4 rows in flight per thread); grid (ceil(N/256), row_splits); smem reduce, one atomic per column.
template <typename TIN>
__global__ void __launch_bounds__(256) colsum_kernel(const TIN* __restrict__ in, long ld, int M, int N, float* __restrict__ out) {
  __shared__ float red[8][256 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 256 + cg * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (n0 < N) {
    const int stride = gridDim.y * 8;
    for (int m = blockIdx.y * 8 + rl; m < M; m += stride * 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int mm = m + u * stride;
        if (mm < M) {
          const TIN* src = in + long(mm) * ld + n0;
          if constexpr (sizeof(TIN) == 2) {
            const uint4 v = *reinterpret_cast<const uint4*>(src);
            const TIN* e = reinterpret_cast<const TIN*>(&v);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += float(e[k]);
          } else {
            const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][cg * 8 + k] = acc[k];
  __syncthreads();
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n < N) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
    atomicAdd(out + n, t);
  }
}

// dst16[g*P + p, :] = src32[(g*group_stride + row_offset + p), :]   (compact the patch-token rows of the gradient stream)
template <int DT>
__global__ void __launch_bounds__(256) cast_rows16_kernel(const float* __restrict__ src, void* __restrict__ dst, long dst_ld, int rows,
                                                          int rows_per_group, int group_stride, int row_offset) {
  using O = Op16<DT>;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long srow = long(row / rows_per_group) * group_stride + row_offset + row % rows_per_group;
  const float4* s = reinterpret_cast<const float4*>(src + srow * D_MODEL);
  uint2* d = reinterpret_cast<uint2*>(reinterpret_cast<typename O::T*>(dst) + long(row) * dst_ld);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 v = s[lane + 32 * i];
    uint2 o;
    o.x = O::pack(v.x, v.y);
    o.y = O::pack(v.z, v.w);
    d[lane + 32 * i] = o;
  }
}

// Gradients of the token-assembly stage (autograd of models/maest.py:670-675, :785-796): CTA per sequence position.
//   position 0/1: d cls_token / d dist_token (+ d new_pos_embed);  position 2+p: conv bias, freq_pe[:, f_p], time_pe[:, off+t_p]
struct TokenGradParams {
  const float* dx; int B, N, P, Tp, Fp, Wt, t_off;
  const int32_t* keep_ft;
  float* d_cls; float* d_dist; float* d_new_pos; float* d_conv_bias; float* d_freq; float* d_time;
};
__global__ void __launch_bounds__(256) token_grad_kernel(const TokenGradParams p) {
  const int pos = blockIdx.x;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int b = 0; b < p.B; ++b) {
    const float* r = p.dx + (long(b) * p.N + pos) * D_MODEL;
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] += r[threadIdx.x + 256 * i];
  }
  if (pos < 2) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = threadIdx.x + 256 * i;
      (pos == 0 ? p.d_cls : p.d_dist)[c] += acc[i];
      p.d_new_pos[pos * D_MODEL + c] += acc[i];
    }
    return;
  }
  const int pi = pos - 2;
  int f, t;
  if (p.keep_ft) { const int ft = p.keep_ft[pi]; f = ft >> 16; t = ft & 0xffff; }
  else { f = pi / p.Tp; t = pi - f * p.Tp; }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = threadIdx.x + 256 * i;
    atomicAdd(p.d_conv_bias + c, acc[i]);
    atomicAdd(p.d_freq + c * p.Fp + f, acc[i]);
    atomicAdd(p.d_time + c * p.Wt + p.t_off + t, acc[i]);
  }
}

// delta[(b*H + h)*N + n] = sum_d dO[row, h*64+d] * O[row, h*64+d]     (softmax backward row term)
template <int DT>
__global__ void __launch_bounds__(256) attn_delta_kernel(const void* __restrict__ o16, const void* __restrict__ do16,
                                                         float* __restrict__ delta, int B, int N, int H) {
  using O = Op16<DT>;
  const long idx = long(blockIdx.x) * blockDim.x + threadIdx.x;   // (row, head)
  const long total = long(B) * N * H;
  if (idx >= total) return;
  const long row = idx / H;
  const int h = int(idx - row * H);
  const uint4* a = reinterpret_cast<const uint4*>(reinterpret_cast<const typename O::T*>(o16) + row * (H * 64) + h * 64);
  const uint4* d = reinterpret_cast<const uint4*>(reinterpret_cast<const typename O::T*>(do16) + row * (H * 64) + h * 64);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 x = a[i], y = d[i];
    const float2 x0 = O::unpack(x.x), x1 = O::unpack(x.y), x2 = O::unpack(x.z), x3 = O::unpack(x.w);
    const float2 y0 = O::unpack(y.x), y1 = O::unpack(y.y), y2 = O::unpack(y.z), y3 = O::unpack(y.w);
    acc += x0.x * y0.x + x0.y * y0.y + x1.x * y1.x + x1.y * y1.y + x2.x * y2.x + x2.y * y2.y + x3.x * y3.x + x3.y * y3.y;
  }
  const int b = int(row / N), n = int(row - long(b) * N);
  delta[(long(b) * H + h) * N + n] = acc;
}

}  // namespace mb
