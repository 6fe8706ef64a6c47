// Fused AdamW (+ optional SWA running average) over every parameter of the model in ONE launch.
// Replaces torch.optim.AdamW as configured by Module.get_optimizer (models/module.py:237-243) -- ~10 element-wise kernels per
// parameter tensor, 152 tensors -- and the per-tensor running average of Lightning's StochasticWeightAveraging that
// helpers/swa_callback.py:11-15 copies into net_swa.
//
// Multi-tensor layout: a device table of per-tensor pointers and a chunk list (tensor index, first element); one CTA per
// chunk of OPT_CHUNK elements.  The arithmetic follows torch's single-tensor AdamW step operation by operation in fp32:
//   p *= 1 - lr*wd;  m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;  denom = sqrt(v)/sqrt(1-b2^t) + eps;  p -= (lr/(1-b1^t)) * m/denom
//   swa += (p - swa) / (n_averaged + 1)                                   (torch.optim.swa_utils default avg_fn)
#pragma once
#include "common.cuh"

namespace mb {

constexpr int OPT_CHUNK = 8192;

struct OptTensor {
  float* p; const float* g; float* m; float* v; float* swa;   // swa may be null
  long n;
};
struct OptChunk { int tensor; int pad; long start; };

struct AdamWParams {
  const OptTensor* tensors;
  const OptChunk* chunks;
  float lr, beta1, beta2, eps, weight_decay;
  float bias_c1, bias_c2_sqrt;     // 1 - beta1^t,  sqrt(1 - beta2^t)
  float grad_scale;                // gradients are multiplied by this first (1 = none; 1/loss_scale for fp16 training)
  float swa_inv;                   // 1 / (n_averaged + 1), or 0 = no SWA update this step
};

__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamWParams a) {
  const OptChunk ch = a.chunks[blockIdx.x];
  const OptTensor t = a.tensors[ch.tensor];
  const long end = ch.start + OPT_CHUNK < t.n ? ch.start + OPT_CHUNK : t.n;
  const float step_size = a.lr / a.bias_c1;
  const float decay = 1.0f - a.lr * a.weight_decay;
  for (long i = ch.start + threadIdx.x; i < end; i += 256) {
    const float g = t.g[i] * a.grad_scale;
    float p = t.p[i] * decay;
    float m = t.m[i];
    m = m + (1.0f - a.beta1) * (g - m);                    // lerp
    const float v = a.beta2 * t.v[i] + (1.0f - a.beta2) * g * g;
    const float denom = sqrtf(v) / a.bias_c2_sqrt + a.eps;
    p -= step_size * (m / denom);
    t.p[i] = p;
    t.m[i] = m;
    t.v[i] = v;
    if (a.swa_inv > 0.f && t.swa != nullptr) {
      const float s = t.swa[i];
      t.swa[i] = s + (p - s) * a.swa_inv;
    }
  }
}

// net_swa <- running average of net (torch.optim.swa_utils semantics: avg += (p - avg) * swa_inv), one launch for all tensors;
// the epoch-end update of helpers/swa_callback.py / Lightning's StochasticWeightAveraging without a second model copy.
__global__ void __launch_bounds__(256) swa_fold_multi_kernel(const OptTensor* tensors, const OptChunk* chunks, const float swa_inv) {
  const OptChunk ch = chunks[blockIdx.x];
  const OptTensor t = tensors[ch.tensor];
  const long end = ch.start + OPT_CHUNK < t.n ? ch.start + OPT_CHUNK : t.n;
  for (long i = ch.start + threadIdx.x; i < end; i += 256) {
    const float s = t.swa[i];
    t.swa[i] = s + (t.p[i] - s) * swa_inv;
  }
}

}  // namespace mb
