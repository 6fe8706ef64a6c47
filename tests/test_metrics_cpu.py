"""Validation metrics without a GPU: the numpy restatement in oracle/metrics_oracle.py against the
installed scikit-learn (the third-party dependency the reference calls, models/module.py:189-190)."""
import numpy as np
import pytest

from oracle import metrics_oracle as MO

sk = pytest.importorskip("sklearn.metrics")


def _data(n, C, seed, ties):
    rng = np.random.RandomState(seed)
    y = (rng.rand(n, C) < 0.15).astype(np.float32)
    y[0, :] = 1
    y[1, :] = 0
    s = rng.rand(n, C).astype(np.float32)
    if ties:
        s = np.round(s * 8) / 8          # many equal scores: thresholds are shared
    return y, s


@pytest.mark.parametrize("ties", [False, True])
def test_oracle_matches_sklearn(ties):
    y, s = _data(500, 7, 3, ties)
    for c in range(y.shape[1]):
        assert abs(MO.average_precision(y[:, c], s[:, c]) - sk.average_precision_score(y[:, c], s[:, c])) < 1e-12
        assert abs(MO.roc_auc(y[:, c], s[:, c]) - sk.roc_auc_score(y[:, c], s[:, c])) < 1e-12
    ap, roc = MO.macro_ap_roc(y, s)
    assert abs(ap - sk.average_precision_score(y, s, average="macro")) < 1e-12
    assert abs(roc - sk.roc_auc_score(y, s, average="macro")) < 1e-12


def test_single_class_behaves_like_sklearn():
    import warnings
    s = np.linspace(0, 1, 10).astype(np.float32)
    for y in (np.zeros(10, dtype=np.float32), np.ones(10, dtype=np.float32)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_roc, ref_ap = sk.roc_auc_score(y, s), sk.average_precision_score(y, s)
        assert np.isnan(MO.roc_auc(y, s)) and np.isnan(ref_roc)          # scikit-learn 1.9: nan + UndefinedMetricWarning
        assert MO.average_precision(y, s) == ref_ap
