"""Round-2 debugging helpers (GPU box): batch invariance of the attention variants, run-to-run determinism of the training backward."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maest_b200 import get_maest, ops, synth
from maest_b200.module import my_mixup
from maest_b200.train import training_forward


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def attn_batch_invariance():
    g = torch.Generator().manual_seed(0)
    B, N = 64, 1685
    qkv = torch.randn(B * N, 2304, generator=g).half().cuda()
    for v in (0, 3, 4):
        ob, lb = ops.attention(qkv, B, N, 12, v, save_lse=True)
        for clip in (0, 5, 63):
            os_, ls = ops.attention(qkv[clip * N:(clip + 1) * N].contiguous(), 1, N, 12, v, save_lse=True)
            d = (ob[clip * N:(clip + 1) * N].float() - os_.float())
            dl = (lb[clip] - ls[0]).abs().max()
            print(f"variant {v} clip {clip}: rel {rel(ob[clip * N:(clip + 1) * N], os_):.3e} differing {int((d != 0).sum())} of {d.numel()} maxabs {float(d.abs().max()):.3e} lse maxabs {float(dl):.3e}")


def train_determinism():
    x, y = synth.train_batch(2)
    for v in (0,):
        flats = []
        for rep in range(3):
            m = get_maest(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=400, input_f=96, input_t=1875, s_patchout_t=90, op_dtype="bf16")
            m.load_state_dict(synth.synth_state_dict(187, 400, seed=0), strict=False)
            m = m.cuda().train()
            m.attn_variant = v
            torch.manual_seed(1); np.random.seed(1)
            loss, _ = training_forward(m, x.cuda(), y.cuda(), my_mixup(2, 0.3))
            loss.backward()
            torch.cuda.synchronize()
            flats.append({n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
        for rep in (1, 2):
            bad = [(n, rel(flats[rep][n], flats[0][n])) for n in flats[0] if rel(flats[rep][n], flats[0][n]) > 1e-6]
            print(f"variant {v} run {rep} vs 0: {len(bad)} tensors differ > 1e-6:", [(n, f"{r:.2e}") for n, r in bad[:4]])
            rl = {n: rel(flats[rep][n], flats[0][n]) for n in flats[0]}
            print("   last blocks:", [(n, f"{r:.1e}") for n, r in rl.items() if n.startswith("blocks.11") or n.startswith("blocks.10.mlp") or n.startswith("norm") or n.startswith("head")])


if __name__ == "__main__":
    {"attn": attn_batch_invariance, "train": train_determinism}[sys.argv[1]]()
