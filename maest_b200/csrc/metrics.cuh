// Validation metrics on the device: per-class average precision and ROC AUC with
// scikit-learn's semantics (models/module.py:189-190 calls metrics.average_precision_score / roc_auc_score on the
// gathered [n_samples, n_classes] arrays on the host).
//
// Input: per class, the labels re-ordered by DESCENDING score and the sorted scores themselves, both [n, C] with the class
// index contiguous (one thread per class reads coalesced rows).  A threshold exists at the last element of every run of
// equal scores (sklearn's `distinct_value_indices`); with tp / fp the cumulative counts at threshold k:
//   AP  = sum_k (tp_k - tp_{k-1}) / P * tp_k / (tp_k + fp_k)            (precision_recall_curve + -sum(diff(recall) * precision))
//   AUC = sum_k (fp_k - fp_{k-1}) * (tp_k + tp_{k-1}) / 2 / (P * N)     (trapezoid over roc_curve; collinear points dropped by
//                                                                        sklearn do not change the area)
// Counts are integers, the two sums run in double precision in a fixed order: results match sklearn to ~1e-15.
#pragma once
#include "common.cuh"

namespace mb {

__global__ void __launch_bounds__(128) ap_roc_kernel(const float* __restrict__ score_sorted, const float* __restrict__ label_sorted,
                                                     int n, int C, double* __restrict__ ap, double* __restrict__ auc,
                                                     int32_t* __restrict__ n_pos) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  long P = 0;
  for (int i = 0; i < n; ++i) P += label_sorted[long(i) * C + c] > 0.5f ? 1 : 0;
  const long Nn = long(n) - P;
  n_pos[c] = int32_t(P);
  if (P == 0 || Nn == 0) {          // scikit-learn >= 1.6: roc_auc_score warns and returns nan; AP is 0 (no positives) or 1
    ap[c] = P == 0 ? 0.0 : 1.0;
    auc[c] = nan("");
    return;
  }
  long tp = 0, fp = 0, tp_prev = 0, fp_prev = 0;
  double ap_sum = 0.0, auc_sum = 0.0;
  float s = score_sorted[c];
  for (int i = 0; i < n; ++i) {
    if (label_sorted[long(i) * C + c] > 0.5f) ++tp; else ++fp;
    const float s_next = i + 1 < n ? score_sorted[long(i + 1) * C + c] : 0.f;
    if (i + 1 == n || s_next != s) {
      ap_sum += double(tp - tp_prev) * (double(tp) / double(tp + fp));
      auc_sum += double(fp - fp_prev) * double(tp + tp_prev);
      tp_prev = tp;
      fp_prev = fp;
    }
    s = s_next;
  }
  ap[c] = ap_sum / double(P);
  auc[c] = 0.5 * auc_sum / (double(P) * double(Nn));
}

}  // namespace mb
