"""Deterministic synthetic weights and inputs (there is no network: no checkpoints, no datasets).

Everything here is generated with CPU `torch.Generator`s, so the same call reproduces bit-identical
tensors in the dev container and on the GPU box (same torch build).  Used by the golden-fixture
generator, the parity tests, `smoke()` and `bench.py`.

Weights follow the reference's state-dict layout (probed from
`models/maest.py:516-586`).  Unlike the reference's init (`models/maest.py:942-976`: zero biases,
unit LayerNorm) every bias / LayerNorm parameter is randomised so that parity tests exercise them.
GEMM weights are rounded so they are exactly representable in BOTH fp16 and bf16: the CUDA path
and the fp32 oracle then start from identical numbers (SURVEY.md §7 "Tolerance").
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

EMBED = 768
DEPTH = 12
HEADS = 12
MLP = 3072
N_MELS = 96
PATCH = 16


def round_operand(w: torch.Tensor) -> torch.Tensor:
    """Round to a value representable in both bf16 and fp16 (fp32 container)."""
    return w.float().bfloat16().float().half().float()


def _tn(gen, shape, std=0.02):
    # truncated normal at +-2 std, like trunc_normal_(std=.02) (models/helpers/vit_helpers.py:110-166)
    x = torch.randn(shape, generator=gen, dtype=torch.float32)
    x = x.clamp_(-2.0, 2.0) * std
    return x


def synth_state_dict(grid_t: int, n_classes: int = 400, seed: int = 0, grid_f: int = 9,
                     depth: int = DEPTH, round_gemm_weights: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """State dict with the reference's keys/shapes. `grid_t` = width of time_new_pos_embed (31/62/125/187)."""
    g = torch.Generator().manual_seed(seed)
    rw = round_operand if round_gemm_weights else (lambda t: t)
    sd = OrderedDict()
    sd["cls_token"] = _tn(g, (1, 1, EMBED))
    sd["dist_token"] = _tn(g, (1, 1, EMBED))
    sd["new_pos_embed"] = _tn(g, (1, 2, EMBED))
    sd["freq_new_pos_embed"] = _tn(g, (1, EMBED, grid_f, 1))
    sd["time_new_pos_embed"] = _tn(g, (1, EMBED, 1, grid_t))
    bound = 1.0 / math.sqrt(PATCH * PATCH)
    sd["patch_embed.proj.weight"] = rw((torch.rand((EMBED, 1, PATCH, PATCH), generator=g) * 2 - 1) * bound)
    sd["patch_embed.proj.bias"] = (torch.rand((EMBED,), generator=g) * 2 - 1) * bound
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + 0.1 * torch.randn((EMBED,), generator=g)
        sd[p + "norm1.bias"] = 0.05 * torch.randn((EMBED,), generator=g)
        sd[p + "attn.qkv.weight"] = rw(_tn(g, (3 * EMBED, EMBED)) * 2.0)
        sd[p + "attn.qkv.bias"] = 0.02 * torch.randn((3 * EMBED,), generator=g)
        sd[p + "attn.proj.weight"] = rw(_tn(g, (EMBED, EMBED)))
        sd[p + "attn.proj.bias"] = 0.02 * torch.randn((EMBED,), generator=g)
        sd[p + "norm2.weight"] = 1.0 + 0.1 * torch.randn((EMBED,), generator=g)
        sd[p + "norm2.bias"] = 0.05 * torch.randn((EMBED,), generator=g)
        sd[p + "mlp.fc1.weight"] = rw(_tn(g, (MLP, EMBED)))
        sd[p + "mlp.fc1.bias"] = 0.02 * torch.randn((MLP,), generator=g)
        sd[p + "mlp.fc2.weight"] = rw(_tn(g, (EMBED, MLP)))
        sd[p + "mlp.fc2.bias"] = 0.02 * torch.randn((EMBED,), generator=g)
    sd["norm.weight"] = 1.0 + 0.1 * torch.randn((EMBED,), generator=g)
    sd["norm.bias"] = 0.05 * torch.randn((EMBED,), generator=g)
    sd["head.0.weight"] = 1.0 + 0.1 * torch.randn((EMBED,), generator=g)
    sd["head.0.bias"] = 0.05 * torch.randn((EMBED,), generator=g)
    sd["head.1.weight"] = _tn(g, (n_classes, EMBED)) * 2.0
    sd["head.1.bias"] = 0.02 * torch.randn((n_classes,), generator=g)
    sd["head_dist.weight"] = _tn(g, (n_classes, EMBED)) * 2.0
    sd["head_dist.bias"] = 0.02 * torch.randn((n_classes,), generator=g)
    return sd


def wave_a(batch: int, samples: int, seed: int = 1234) -> torch.Tensor:
    """wave-A of SURVEY.md §8(d): full-scale uniform noise in [-1, 1), fp32 [B, S]."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, samples, generator=g) * 2 - 1


def wave_b(samples: int, seed: int = 4321) -> torch.Tensor:
    """wave-B of SURVEY.md §8(d): quiet tonal mixture; exercises log10(1+1e4*x) near zero. fp32 [S]."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(samples, dtype=torch.float64) / 16000.0
    f = torch.logspace(math.log10(55.0), math.log10(7000.0), 8, dtype=torch.float64)
    phi = torch.rand(8, generator=g, dtype=torch.float64) * 2 * math.pi
    x = 0.05 * torch.sin(2 * math.pi * f[:, None] * t[None, :] + phi[:, None]).sum(0)
    x = x + 1e-3 * torch.randn(samples, generator=g, dtype=torch.float64)
    x[: samples // 10] = 0.0
    return x.float()


def train_batch(batch: int, frames: int = 1875, n_classes: int = 400, seed_x: int = 7, seed_y: int = 8):
    """Config-4 batch of SURVEY.md §8(d): fp16 mel [B,1,96,frames], fp16 multi-hot targets [B,C]."""
    gx = torch.Generator().manual_seed(seed_x)
    gy = torch.Generator().manual_seed(seed_y)
    x = (0.5 * torch.randn(batch, 1, N_MELS, frames, generator=gx)).half()
    y = torch.zeros(batch, n_classes)
    for b in range(batch):
        k = int(torch.randint(1, 6, (1,), generator=gy))
        idx = torch.randperm(n_classes, generator=gy)[:k]
        y[b, idx] = 1.0
    return x, y.half()
