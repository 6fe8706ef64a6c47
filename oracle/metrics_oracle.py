"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the two scikit-learn metrics the reference's validation epoch uses
(models/module.py:189-190: `metrics.average_precision_score(y, y_hat, average="macro")`, `metrics.roc_auc_score(...)`).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

The arithmetic lives in a third-party dependency that is not under /root/reference: scikit-learn (unpinned in the
reference's pyproject.toml; 1.9.0 in this image), `sklearn/metrics/_ranking.py`: `_binary_clf_curve` (stable sort by
descending score, one threshold per distinct score, cumulative tp / fp), `precision_recall_curve` +
`average_precision_score` (AP = -sum(diff(recall) * precision[:-1])), `roc_curve` + `auc` (trapezoid).  Pinned against
the installed scikit-learn itself in tests/test_metrics_cpu.py."""
from __future__ import annotations

import numpy as np


def binary_clf_curve(y_true: np.ndarray, y_score: np.ndarray):
    order = np.argsort(-y_score.astype(np.float64), kind="mergesort")
    y_score = y_score[order]
    y_true = (y_true[order] > 0.5).astype(np.float64)
    distinct = np.where(np.diff(y_score))[0]
    idx = np.r_[distinct, y_true.size - 1]
    tps = np.cumsum(y_true)[idx]
    fps = 1 + idx - tps
    return fps, tps


def average_precision(y_true: np.ndarray, y_score: np.ndarray) -> float:
    fps, tps = binary_clf_curve(y_true, y_score)
    precision = tps / (tps + fps)
    if tps[-1] == 0:                      # no positives: sklearn sets recall to one for all thresholds -> AP = precision[0] = 0
        return 0.0
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def roc_auc(y_true: np.ndarray, y_score: np.ndarray) -> float:
    fps, tps = binary_clf_curve(y_true, y_score)
    if tps[-1] == 0 or fps[-1] == 0:      # sklearn >= 1.6 warns (UndefinedMetricWarning) and returns nan; older versions raise
        return float("nan")
    fpr = np.r_[0.0, fps / fps[-1]]
    tpr = np.r_[0.0, tps / tps[-1]]
    return float(np.trapezoid(tpr, fpr) if hasattr(np, "trapezoid") else np.trapz(tpr, fpr))


def macro_ap_roc(y: np.ndarray, y_hat: np.ndarray):
    """[n, C] arrays -> (macro AP, macro ROC AUC), as models/module.py:189-190."""
    C = y.shape[1]
    return (float(np.mean([average_precision(y[:, c], y_hat[:, c]) for c in range(C)])),
            float(np.mean([roc_auc(y[:, c], y_hat[:, c]) for c in range(C)])))
