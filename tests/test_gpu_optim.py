"""Fused AdamW + SWA average on the GPU against torch.optim.AdamW / swa_utils.AveragedModel's
update rule on identical gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch_adamw_and_swa():
    from maest_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(0)
    shapes = [(768, 768), (3072,), (1, 1, 768), (9000,), (5,)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    frozen = torch.nn.Parameter(torch.ones(3).cuda())           # never receives a gradient (like head_dist.* in "mean" mode)
    swa = [p.detach().clone() for p in our_p] + [frozen.detach().clone()]
    ref = torch.optim.AdamW(ref_p, lr=3e-3, weight_decay=0.05)
    ours = FusedAdamW(our_p + [frozen], lr=3e-3, weight_decay=0.05, swa_params=swa)
    swa_ref = [p.detach().clone() for p in ref_p]
    n_avg = 0
    for step in range(6):
        for a, b in zip(ref_p, our_p):
            gr = torch.randn(a.shape, generator=g).cuda() * (0.1 + step)
            a.grad, b.grad = gr.clone(), gr.clone()
        ref.step()
        do_swa = step >= 2
        ours.step(update_swa=do_swa)
        if do_swa:
            for s, p in zip(swa_ref, ref_p):
                s += (p.detach() - s) / (n_avg + 1)           # torch.optim.swa_utils default avg_fn
            n_avg += 1
        for a, b in zip(ref_p, our_p):
            assert float((a - b).detach().abs().max()) <= 2e-6 * float(a.detach().abs().max()), step
    for s, r in zip(swa, swa_ref):
        assert float((s - r).abs().max()) <= 2e-6 * float(r.abs().max())
    assert torch.equal(frozen.detach(), torch.ones(3).cuda()) and ours.n_averaged == 4
    st = ours.state[our_p[0]]
    rm = ref.state[ref_p[0]]["exp_avg"]
    assert st["step"] == 6 and float((st["exp_avg"] - rm).abs().max()) <= 2e-6 * float(rm.abs().max())


def test_module_uses_the_fused_optimizer_on_cuda():
    from maest_b200 import get_maest
    from maest_b200.module import Module
    from maest_b200.optim import FusedAdamW
    mod = Module(net=get_maest(arch="discogs-maest-5s-pw-129e", pretrained=False)).cuda()
    cfg = mod.configure_optimizers()
    assert isinstance(cfg["optimizer"], FusedAdamW)
