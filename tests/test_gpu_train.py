"""GPU parity of the training step (BASELINE.json configs[3]: maest_30s_from_passt_pretrain, mel [B,1,96,1875] fp16,
s_patchout_t=90, mixup 0.3, BCE) against the reference's own `Module.training_step` + backward (tests/golden/c4.npz,
fp32 CPU run of the unmodified reference with identical seeds / host RNG draws).

Tolerances: bf16 operands (the training dtype) loss 1e-4 abs, gradients 1e-2 rel-L2; fp16 operands with static loss
scaling 2e-3.  The reference's own 16-mixed run is not bit-comparable to its fp32 run either."""
import numpy as np
import pytest
import torch

from maest_b200 import _lib, get_maest, ops, synth
from maest_b200.module import Module, my_mixup
from maest_b200.train import training_forward

pytestmark = pytest.mark.gpu

VEC = ["cls_token", "time_new_pos_embed", "freq_new_pos_embed", "patch_embed.proj.bias", "blocks.0.norm1.weight",
       "blocks.0.attn.qkv.bias", "blocks.5.mlp.fc1.bias", "blocks.11.attn.proj.bias", "norm.weight", "head.0.bias", "head.1.bias"]
MAT = ["patch_embed.proj.weight", "blocks.0.attn.qkv.weight", "blocks.5.mlp.fc1.weight", "blocks.11.mlp.fc2.weight", "head.1.weight"]


def rel(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().flatten().cpu()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def make_train_model(dt):
    m = get_maest(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=400, input_f=96, input_t=1875,
                  s_patchout_t=90, op_dtype=dt)
    m.load_state_dict(synth.synth_state_dict(187, 400, seed=0), strict=False)
    return m.cuda().train()


@pytest.mark.parametrize("dt,tol", [("bf16", 1e-2), ("fp16", 2e-3)])
def test_training_step_vs_reference(golden, dt, tol):
    g = golden["c4"]
    m = make_train_model(dt)
    x, y = synth.train_batch(2)
    torch.manual_seed(1)
    np.random.seed(1)
    mix = my_mixup(2, 0.3)                      # same host RNG order as the reference
    loss, logits = training_forward(m, x.cuda(), y.cuda(), mix)
    assert logits.shape == (2, 400) and not logits.requires_grad
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4
    loss.backward()
    grads = {n: p.grad for n, p in m.named_parameters()}
    assert grads["head_dist.weight"] is None and grads["head_dist.bias"] is None     # unused in "mean" mode, as in the reference
    for k in VEC:
        assert rel(grads[k], g["grad." + k]) < tol, k
    for k in MAT:
        gk = grads[k]
        assert rel(gk.reshape(gk.shape[0], -1)[::37, ::29], g["grad." + k + ".sub"]) < tol, k
        assert abs(float(gk.double().norm()) / float(g["gnorm." + k]) - 1) < tol, k


def test_module_training_step_and_optimizer():
    """Lightning-style surface: Module.training_step -> loss.backward() -> AdamW.step(); loss goes down on a fixed batch."""
    net = make_train_model("bf16")
    mod = Module(net=net, mixup_alpha=0.3, do_swa=False)
    opt = torch.optim.AdamW(mod.parameters(), lr=1e-4, weight_decay=1e-4)
    x, y = synth.train_batch(4)
    batch = (x.cuda(), ["f"] * 4, y.cuda())
    torch.manual_seed(0)
    np.random.seed(0)
    losses = []
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        loss = mod.training_step(batch, 0)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    mod.eval()
    with torch.no_grad():
        out = mod.predict_step(batch, 0)
    assert out["logits"].shape == (4, 400) and out["embeddings"].shape == (4, 768)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_backward_gemm_variants(dt):
    g = torch.Generator().manual_seed(5)
    Mt, Nout, Kin = 1732, 768, 3072
    dY = (torch.randn(Mt, Nout, generator=g) * 0.1).to(dt).cuda()
    X = (torch.randn(Mt, Kin, generator=g) * 0.5).to(dt).cuda()
    W = (torch.randn(Nout, Kin, generator=g) * 0.05).to(dt).cuda()
    tol16 = 1e-3 if dt == torch.float16 else 5e-3
    assert rel(ops.gemm(dY, W, _lib.EPI_STORE32, Mt, Kin, Nout, b_mn=True), dY.double() @ W.double()) < 1e-5
    assert rel(ops.gemm(dY, W, _lib.EPI_STORE16, Mt, Kin, Nout, b_mn=True), dY.double() @ W.double()) < tol16
    upre = (torch.randn(Mt, Kin, generator=g) * 1.5).to(dt).cuda()
    up = upre.double().requires_grad_(True)
    torch.nn.functional.gelu(up).backward(dY.double() @ W.double())
    assert rel(ops.gemm(dY, W, _lib.EPI_GELUBWD16, Mt, Kin, Nout, b_mn=True, aux16=upre), up.grad) < tol16
    for splits in (1, 3, 7):
        dW = torch.zeros(Nout, Kin, device="cuda")
        ops.gemm(dY, X, _lib.EPI_ATOMIC32, Nout, Kin, Mt, a_mn=True, b_mn=True, out=dW, k_splits=splits)
        assert rel(dW, dY.double().t() @ X.double()) < 1e-5


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("BN", [(1, 128), (2, 100), (2, 866), (1, 300)])
def test_attention_backward_vs_autograd(dt, BN):
    B, N = BN
    g = torch.Generator().manual_seed(B * 100 + N)
    qkv = torch.randn(B * N, 2304, generator=g).to(dt).cuda()
    d_o = (torch.randn(B * N, 768, generator=g) * 0.1).to(dt).cuda()
    o, lse = ops.attention(qkv, B, N, 12, 0, save_lse=True)
    qd = qkv.double().requires_grad_(True)
    q, k, v = qd.view(B, N, 3, 12, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, 768).backward(d_o.double())
    assert float((lse.double() - torch.logsumexp(s, -1) * 1.4426950408889634).abs().max()) < 1e-4
    dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, N, 12)
    tol = 1e-3 if dt == torch.float16 else 6e-3
    for lo, hi in ((0, 768), (768, 1536), (1536, 2304)):
        assert rel(dqkv[:, lo:hi], qd.grad[:, lo:hi]) < tol


def test_layernorm_bwd_mixup_bce_units():
    g = torch.Generator().manual_seed(7)
    rows = 1000
    x = torch.randn(rows, 768, generator=g) * 2 + 0.3
    gam, dy, dx0 = torch.randn(768, generator=g), torch.randn(rows, 768, generator=g), torch.randn(rows, 768, generator=g)
    xd, gd = x.double().requires_grad_(True), gam.double().requires_grad_(True)
    bd = torch.zeros(768, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.layer_norm(xd, (768,), gd, bd, 1e-6).backward(dy.double())
    _, mean, rstd = ops.layernorm16(x.cuda(), gam.cuda(), torch.zeros(768).cuda(), 1e-6, "fp16", save_stats=True)
    dx, dg, db = dx0.clone().cuda(), torch.zeros(768).cuda(), torch.zeros(768).cuda()
    ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, gam.cuda(), dx, dg, db, "fp16")
    assert rel(dx, dx0.double() + xd.grad) < 1e-5 and rel(dg, gd.grad) < 1e-5 and rel(db, bd.grad) < 1e-5
    xm = torch.randn(4, 1, 96, 100, generator=g).half()
    perm, lam = torch.tensor([2, 0, 3, 1]), torch.tensor([0.9, 0.6, 0.75, 0.51])
    ref = xm.float() * lam.view(4, 1, 1, 1) + xm.float()[perm] * (1 - lam.view(4, 1, 1, 1))
    assert float((ops.mixup(xm.cuda(), perm.cuda(), lam.cuda()).cpu() - ref).abs().max()) < 1e-6
    lz, ly = torch.randn(8, 400, generator=g) * 3, (torch.rand(8, 400, generator=g) > 0.9).float()
    loss, dz = ops.bce_logits(lz.cuda(), ly.cuda())
    lzd = lz.double().requires_grad_(True)
    lref = torch.nn.functional.binary_cross_entropy_with_logits(lzd, ly.double())
    lref.backward()
    assert abs(float(loss) - float(lref)) < 1e-6 and rel(dz, lzd.grad) < 1e-5
