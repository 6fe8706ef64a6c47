"""Validation path on the GPU: the AP / ROC AUC kernel against the numpy oracle (itself pinned to
scikit-learn) and the Module's twin-net validation epoch against the reference's formulae."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,C,ties", [(500, 7, False), (500, 7, True), (3000, 400, True), (2, 3, False)])
def test_ap_roc_kernel_vs_oracle(n, C, ties):
    from maest_b200 import ops
    rng = np.random.RandomState(n + C)
    y = (rng.rand(n, C) < 0.2).astype(np.float32)
    y[0, :], y[1, :] = 1, 0
    s = rng.rand(n, C).astype(np.float32)
    if ties:
        s = np.round(s * 16) / 16
    ap, auc, npos = ops.ap_roc(torch.from_numpy(y).cuda(), torch.from_numpy(s).cuda())
    assert np.array_equal(npos.cpu().numpy(), y.sum(0).astype(np.int32))
    for c in range(0, C, max(1, C // 25)):
        assert abs(float(ap[c]) - MO.average_precision(y[:, c], s[:, c])) < 1e-12, c
        assert abs(float(auc[c]) - MO.roc_auc(y[:, c], s[:, c])) < 1e-12, c
    # a class with a single label value: NaN AUC and AP 0, as scikit-learn >= 1.6 returns
    y[:, 0] = 0
    _, auc, _ = ops.ap_roc(torch.from_numpy(y).cuda(), torch.from_numpy(s).cuda())
    assert bool(torch.isnan(auc[0])) and not bool(torch.isnan(auc[1:]).any())
    assert float(ops.ap_roc(torch.from_numpy(y).cuda(), torch.from_numpy(s).cuda())[0][0]) == 0.0


def test_module_validation_epoch_with_swa_twin():
    import copy
    from maest_b200 import get_maest, synth
    from maest_b200.module import Module
    net = get_maest(arch="discogs-maest-5s-pw-129e", pretrained=False)
    net.load_state_dict(synth.synth_state_dict(31, 400, seed=0), strict=False)
    mod = Module(net=net.cuda().eval(), do_swa=True)
    mod.net_swa = copy.deepcopy(mod.net)                      # what helpers/swa_callback.py:43-44 does on fit start
    with torch.no_grad():
        mod.net_swa.head[1].bias.add_(0.05)                   # make the twin differ
    g = torch.Generator().manual_seed(0)
    ys, yh, yh_swa, losses = [], [], [], []
    for i in range(3):
        x = (torch.rand(4, 1, 96, 312, generator=g) - 0.3).cuda()
        y = (torch.rand(4, 400, generator=g) < 0.3).half().cuda()
        y[0], y[1] = 1, 0
        out = mod.validation_step((x, ["f"] * 4, y), i)
        ys.append(y.float().cpu().numpy())
        yh.append(out["y_hat"].cpu().numpy())
        yh_swa.append(out["swa_y_hat"].cpu().numpy())
        with torch.no_grad():
            lo, _ = mod.net(x)
        losses.append(float(torch.nn.functional.binary_cross_entropy_with_logits(lo, y.float())))
        assert abs(float(out["loss"]) - losses[-1]) < 1e-5
        assert np.allclose(yh[-1], torch.sigmoid(lo).cpu().numpy(), atol=1e-6)
    res = mod.on_validation_epoch_end()
    y, a, b = np.concatenate(ys), np.concatenate(yh), np.concatenate(yh_swa)
    ap, roc = MO.macro_ap_roc(y, a)
    ap_s, roc_s = MO.macro_ap_roc(y, b)
    assert abs(res["val_ap"] - ap) < 1e-9 and abs(res["val_roc"] - roc) < 1e-9
    assert abs(res["val_ap_swa"] - ap_s) < 1e-9 and abs(res["val_roc_swa"] - roc_s) < 1e-9
    assert abs(res["val_loss"] - np.mean(losses)) < 1e-5 and "val_loss_swa" in res
    assert mod.validation_outputs == []
