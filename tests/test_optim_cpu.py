"""Optimiser side without a GPU: the learning-rate lambdas against values produced by the
reference's helpers/ramp.py (tests/golden/c7_sched.npz, tests/golden/make_golden_sched.py), Module.configure_optimizers' shape."""
import numpy as np
import pytest
import torch


def test_lr_lambdas_match_reference(golden):
    from maest_b200 import optim
    g = golden["c7_sched"]
    ep = g["epochs"]
    f = optim.get_scheduler_lambda(5, 50, 50, 0.01, "exp_lin")
    assert np.array_equal(np.array([f(int(e)) for e in ep]), g["exp_lin"])
    f = optim.exp_warmup_linear_down(20, 100, 50, 0.001)
    assert np.array_equal(np.array([f(int(e)) for e in ep]), g["exp_lin_b"])
    f = optim.get_scheduler_lambda(5, 50, 50, 0.01, "cos_cyc")
    assert np.array_equal(np.array([f(int(e)) for e in ep]), g["cos_cyc"])
    with pytest.raises(RuntimeError):
        optim.get_scheduler_lambda(5, 50, 50, 0.01, "nope")


def test_configure_optimizers_mirrors_reference():
    from maest_b200 import get_maest
    from maest_b200.module import Module
    mod = Module(net=get_maest(arch="discogs-maest-5s-pw-129e", pretrained=False))
    cfg = mod.configure_optimizers()                       # models/module.py:245-254: optimizer + LambdaLR
    opt, sch = cfg["optimizer"], cfg["lr_scheduler"]
    assert isinstance(opt, torch.optim.AdamW) and not type(opt).__name__.startswith("Fused")      # CPU parameters: torch's AdamW
    assert opt.defaults["lr"] == 2e-5 and opt.defaults["weight_decay"] == 1e-4
    assert isinstance(sch, torch.optim.lr_scheduler.LambdaLR)
    lam = mod.get_scheduler_lambda()
    assert abs(opt.param_groups[0]["lr"] - 2e-5 * lam(0)) < 1e-12


def test_fused_adamw_has_no_cpu_fallback():
    from maest_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.ones(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdamW([p]).step()
