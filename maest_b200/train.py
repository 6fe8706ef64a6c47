"""Training step of the MAEST hot path on the B200 kernels: forward with saved activations + hand-written backward,
wrapped in ONE `torch.autograd.Function` so that Lightning / DDP / AdamW see ordinary parameters and `.grad`s.

Forward (train mode): [mixup] -> patch tokens (+pos-embed, +patchout) -> 12 x (LN, qkv, attention, proj+residual, LN,
fc1+GELU, fc2+residual) -> final LN of rows 0/1, (cls+dist)/2, head -> BCE-with-logits   (models/module.py:73-102).
Backward mirrors it with the same tcgen05 GEMM kernel reading every operand in its natural layout
(input gradients: B MN-major; weight gradients: both operands MN-major, split-K with fp32 atomics).
torch is used for memory, autograd bookkeeping and (in DDP) the NCCL all-reduce only.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops

E = 768


def _flat_params(model):
    return [(n, p) for n, p in model.named_parameters()]


class MaestTrainStep(torch.autograd.Function):
    """loss, logits = MaestTrainStep.apply(model, mel, targets, keep_ft, t_offset, *parameters)

    `targets` is one tensor ("mean" mode: BCE(head((cls + dist) / 2), y), models/module.py:88-90) or a pair (y, y_teacher) for
    distilled_type = "separated" (teacher-student step, models/module.py:279-313):
    loss = (BCE(head(cls), y) + BCE(head_dist(dist), y_teacher)) / 2; then `logits` is the pair (logits, logits_dist)."""

    @staticmethod
    def forward(ctx, model, mel, targets, keep_ft, t_offset, *params):
        with ops.launch_scope(mel.device):
            return MaestTrainStep._forward(ctx, model, mel, targets, keep_ft, t_offset, *params)

    @staticmethod
    def _forward(ctx, model, mel, targets, keep_ft, t_offset, *params):
        dt = model.op_dtype
        dev = mel.device
        names = [n for n, _ in _flat_params(model)]
        P_ = dict(zip(names, params))
        f32 = lambda n: P_[n].detach().float().contiguous()            # noqa: E731
        w16 = lambda n: model._weight16(n, P_[n + ".weight"])          # noqa: E731
        if mel.dim() == 4:
            mel = mel[:, 0]
        if mel.dtype not in (torch.float32, torch.float16):
            mel = mel.float()
        tok, a16 = ops.patch_tokens(
            mel.contiguous(), w16("patch_embed.proj"), f32("patch_embed.proj.bias"),
            f32("freq_new_pos_embed").reshape(E, -1), f32("time_new_pos_embed").reshape(E, -1), f32("cls_token").reshape(-1),
            f32("dist_token").reshape(-1), f32("new_pos_embed").reshape(2, E), keep_ft=keep_ft, t_offset=t_offset,
            return_patches=True)
        B, N, _ = tok.shape
        M = B * N
        x = tok.view(M, E)
        saved = []
        for i in range(len(model.blocks)):
            pre = f"blocks.{i}."
            h1, mean1, rstd1 = ops.layernorm16(x, f32(pre + "norm1.weight"), f32(pre + "norm1.bias"), 1e-6, dt, save_stats=True)
            qkv = ops.linear(h1, w16(pre + "attn.qkv"), f32(pre + "attn.qkv.bias"), _lib.EPI_STORE16)
            o, lse = ops.attention(qkv, B, N, 12, model.attn_variant, save_lse=True)
            x_mid = torch.empty_like(x)
            ops.linear(o, w16(pre + "attn.proj"), f32(pre + "attn.proj.bias"), _lib.EPI_RESID32, resid=x, out=x_mid)
            h2, mean2, rstd2 = ops.layernorm16(x_mid, f32(pre + "norm2.weight"), f32(pre + "norm2.bias"), 1e-6, dt, save_stats=True)
            upre = torch.empty((M, 4 * E), device=dev, dtype=h2.dtype)
            u = ops.gemm(h2, w16(pre + "mlp.fc1"), _lib.EPI_GELU16, M, 4 * E, E, bias=f32(pre + "mlp.fc1.bias"), aux16=upre)
            x_out = torch.empty_like(x)
            ops.linear(u, w16(pre + "mlp.fc2"), f32(pre + "mlp.fc2.bias"), _lib.EPI_RESID32, resid=x_mid, out=x_out)
            saved.append((x, mean1, rstd1, h1, qkv, lse, o, x_mid, mean2, rstd2, h2, upre, u))
            x = x_out
        sep = isinstance(targets, (tuple, list))
        if sep:
            logits, logits_dist, feats = ops.pool_head(x, B, N, f32("norm.weight"), f32("norm.bias"), f32("head.0.weight"),
                                                       f32("head.0.bias"), f32("head.1.weight"), f32("head.1.bias"),
                                                       f32("head_dist.weight"), f32("head_dist.bias"), separated=True)
            loss_s, dlogits = ops.bce_logits(logits, targets[0])
            loss_t, dlogits_dist = ops.bce_logits(logits_dist, targets[1])
            loss = (loss_s + loss_t) * 0.5
            ctx.dlogits_dist = dlogits_dist
            ctx.loss_parts = (loss_s, loss_t)
        else:
            logits, _, feats = ops.pool_head(x, B, N, f32("norm.weight"), f32("norm.bias"), f32("head.0.weight"), f32("head.0.bias"),
                                             f32("head.1.weight"), f32("head.1.bias"))
            loss, dlogits = ops.bce_logits(logits, targets)
            ctx.dlogits_dist = None
        ctx.model, ctx.names, ctx.saved, ctx.x_final = model, names, saved, x
        ctx.dims = (B, N, int(a16.shape[0] // B), mel.shape[-1], int(t_offset))
        ctx.keep_ft, ctx.a16, ctx.dlogits = keep_ft, a16, dlogits
        ctx.params = P_
        if sep:
            ls_d, lt_d = loss_s.detach().clone(), loss_t.detach().clone()
            ctx.mark_non_differentiable(logits, logits_dist, ls_d, lt_d)
            return loss, logits, logits_dist, ls_d, lt_d
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, dloss, *_unused):
        with ops.launch_scope(dloss.device):
            return MaestTrainStep._backward(ctx, dloss, *_unused)

    @staticmethod
    def _backward(ctx, dloss, *_unused):
        model, names, P_ = ctx.model, ctx.names, ctx.params
        dt = model.op_dtype
        B, N, P, T, t_off = ctx.dims
        M = B * N
        dev = dloss.device
        f32 = lambda n: P_[n].detach().float().contiguous()            # noqa: E731
        w16 = lambda n: model._weight16(n, P_[n + ".weight"])          # noqa: E731
        # One flat fp32 gradient buffer in parameter order (embedding params | blocks 0..11 | norm + head); every
        # parameter gradient is a view into it.  With `model.grad_allreduce` set (data-parallel run without the DDP
        # wrapper) each block's slice is all-reduced over NCCL as soon as its backward has been enqueued, so the
        # collective overlaps the remaining backward kernels; everything else about DDP stays outside.
        sep = ctx.dlogits_dist is not None
        gnames = [n for n in names if sep or not n.startswith("head_dist")]
        offs, total = {}, 0
        for n in gnames:
            offs[n] = total
            total += P_[n].numel()
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        G = {n: flat[offs[n]: offs[n] + P_[n].numel()].view(P_[n].shape) for n in gnames}
        g2 = lambda n, r, c: G[n].view(r, c)                           # noqa: E731
        sync = getattr(model, "grad_allreduce", None)
        works = []

        # "overlap": issue each slice's all-reduce as soon as it is complete.  Measured on 2xB200 this is SLOWER than one
        # all-reduce at the end (76.9 vs ~63 ms/step): the NCCL kernels take SMs away from the persistent, statically
        # scheduled GEMM CTAs (1 per SM), whose late CTAs then become a long tail.  Default: a single flat all-reduce
        # after the last backward kernel; per-slice overlap stays available as model.grad_allreduce = "overlap".
        overlap = sync == "overlap"
        # "bf16": the flat buffer is all-reduced as bfloat16 (172 instead of 343 MB on the wire, like DDP's bf16_compress_hook:
        # every rank's gradient is rounded to 8 mantissa bits before the sum) -- opt-in, the default keeps the reference's fp32.
        compress = sync == "bf16"

        def reduce_range(first, last_exclusive):
            if sync and overlap:
                import torch.distributed as dist
                grp = None if sync in (True, "overlap") else sync
                works.append(dist.all_reduce(flat[first:last_exclusive], op=dist.ReduceOp.SUM, group=grp, async_op=True))

        blk_first = [offs[f"blocks.{i}.norm1.weight"] for i in range(len(model.blocks))]
        blk_end = blk_first[1:] + [offs["norm.weight"]]
        # fp16 operands need loss scaling (the reference's "16-mixed" uses GradScaler); bf16 does not.  The scale is applied
        # to d(loss) on the way in and divided out of the fp32 parameter gradients on the way out.
        ls = float(getattr(model, "train_loss_scale", None) or (4096.0 if ops.op_dtype_code(dt) == ops.F16 else 1.0))
        gscale = (dloss.detach().float().reshape(1) * (ls * (0.5 if sep else 1.0))).contiguous()   # separated: loss = (l_s + l_t) / 2
        lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
        st = torch.cuda.current_stream().cuda_stream
        C_ = P_["head.1.weight"].shape[0]

        dx = torch.zeros((M, E), device=dev, dtype=torch.float32)
        hz = torch.empty((B, E), device=dev, dtype=torch.float32)
        if sep:
            z1 = torch.empty((B, E), device=dev, dtype=torch.float32)
            _lib.check(lib.maest_head_bwd_separated(
                ctx.x_final.data_ptr(), B, N, ctx.dlogits.data_ptr(), ctx.dlogits_dist.data_ptr(), gscale.data_ptr(),
                f32("norm.weight").data_ptr(), f32("norm.bias").data_ptr(), f32("head.0.weight").data_ptr(),
                f32("head.0.bias").data_ptr(), f32("head.1.weight").data_ptr(), f32("head_dist.weight").data_ptr(), C_,
                dx.data_ptr(), hz.data_ptr(), z1.data_ptr(), G["norm.weight"].data_ptr(), G["norm.bias"].data_ptr(),
                G["head.0.weight"].data_ptr(), G["head.0.bias"].data_ptr(), G["head.1.weight"].data_ptr(),
                G["head.1.bias"].data_ptr(), G["head_dist.weight"].data_ptr(), G["head_dist.bias"].data_ptr(), st), "head_bwd_separated")
        else:
            _lib.check(lib.maest_head_bwd(ctx.x_final.data_ptr(), B, N, ctx.dlogits.data_ptr(), gscale.data_ptr(),
                                          f32("norm.weight").data_ptr(), f32("norm.bias").data_ptr(), f32("head.0.weight").data_ptr(),
                                          f32("head.0.bias").data_ptr(), f32("head.1.weight").data_ptr(), C_, dx.data_ptr(), hz.data_ptr(),
                                          G["norm.weight"].data_ptr(), G["norm.bias"].data_ptr(), G["head.0.weight"].data_ptr(),
                                          G["head.0.bias"].data_ptr(), G["head.1.weight"].data_ptr(), G["head.1.bias"].data_ptr(), st),
                       "head_bwd")
        reduce_range(offs["norm.weight"], total)                      # final norm + head gradients are complete
        dx16 = ops.cast_rows16(dx, M, dt)
        dh = torch.empty((M, E), device=dev, dtype=torch.float32)
        dupre = torch.empty((M, 4 * E), device=dev, dtype=dx16.dtype)
        d_o = torch.empty((M, E), device=dev, dtype=dx16.dtype)
        dqkv = torch.empty((M, 3 * E), device=dev, dtype=dx16.dtype)

        def wgrad(dy16, x16, n_out, n_in, gname):
            ops.gemm(dy16, x16, _lib.EPI_ATOMIC32, n_out, n_in, M, a_mn=True, b_mn=True, out=g2(gname, n_out, n_in),
                     k_splits=ops.wgrad_splits(n_out, n_in, M))

        for i in reversed(range(len(model.blocks))):
            pre = f"blocks.{i}."
            x_in, mean1, rstd1, h1, qkv, lse, o, x_mid, mean2, rstd2, h2, upre, u = ctx.saved[i]
            # ---- MLP branch: x_out = x_mid + fc2(gelu(fc1(LN2(x_mid))))
            wgrad(dx16, u, E, 4 * E, pre + "mlp.fc2.weight")
            if i == len(model.blocks) - 1:     # every other fc2 / proj bias gradient comes out of the LayerNorm-backward kernel
                ops.colsum(dx, G[pre + "mlp.fc2.bias"])   # that produced this dx (dx_colsum below)
            ops.gemm(dx16, w16(pre + "mlp.fc2"), _lib.EPI_GELUBWD16, M, 4 * E, E, b_mn=True, out=dupre, aux16=upre,
                     colsum_out=G[pre + "mlp.fc1.bias"])       # fc1 bias gradient from the same epilogue (no colsum pass over dupre)
            wgrad(dupre, h2, 4 * E, E, pre + "mlp.fc1.weight")
            ops.gemm(dupre, w16(pre + "mlp.fc1"), _lib.EPI_STORE32, M, E, 4 * E, b_mn=True, out=dh)
            ops.layernorm_bwd(dh, x_mid, mean2, rstd2, f32(pre + "norm2.weight"), dx, G[pre + "norm2.weight"], G[pre + "norm2.bias"],
                              dt, dx16=dx16, dx_colsum=G[pre + "attn.proj.bias"])
            # ---- attention branch: x_mid = x_in + proj(attn(qkv(LN1(x_in))))
            wgrad(dx16, o, E, E, pre + "attn.proj.weight")
            ops.gemm(dx16, w16(pre + "attn.proj"), _lib.EPI_STORE16, M, E, E, b_mn=True, out=d_o)
            ops.attention_bwd(qkv, o, d_o, lse, B, N, 12, out=dqkv)
            wgrad(dqkv, h1, 3 * E, E, pre + "attn.qkv.weight")
            ops.colsum(dqkv, G[pre + "attn.qkv.bias"])
            ops.gemm(dqkv, w16(pre + "attn.qkv"), _lib.EPI_STORE32, M, E, 3 * E, b_mn=True, out=dh)
            ops.layernorm_bwd(dh, x_in, mean1, rstd1, f32(pre + "norm1.weight"), dx, G[pre + "norm1.weight"], G[pre + "norm1.bias"],
                              dt, dx16=dx16, dx_colsum=G[f"blocks.{i - 1}.mlp.fc2.bias"] if i > 0 else None)
            ctx.saved[i] = None
            reduce_range(blk_first[i], blk_end[i])
        # ---- token assembly + patch embedding
        Fp, Tp = 9, (T - 16) // 10 + 1
        Wt = P_["time_new_pos_embed"].shape[-1]
        dtok16 = ops.cast_rows16(dx, B * P, dt, rows_per_group=P, group_stride=2 + P, row_offset=2)
        ops.gemm(dtok16, ctx.a16, _lib.EPI_ATOMIC32, E, 256, B * P, a_mn=True, b_mn=True, out=g2("patch_embed.proj.weight", E, 256),
                 k_splits=ops.wgrad_splits(E, 256, B * P))
        _lib.check(lib.maest_token_grad(dx.data_ptr(), B, N, P, Tp, Fp, Wt, t_off, ops._p(ctx.keep_ft), G["cls_token"].data_ptr(),
                                        G["dist_token"].data_ptr(), G["new_pos_embed"].data_ptr(), G["patch_embed.proj.bias"].data_ptr(),
                                        G["freq_new_pos_embed"].data_ptr(), G["time_new_pos_embed"].data_ptr(), st), "token_grad")
        reduce_range(0, blk_first[0])
        scale = 1.0 / ls
        if sync:
            import torch.distributed as dist
            grp = None if sync in (True, "overlap", "bf16") else sync
            if not overlap:
                ev = getattr(model, "allreduce_events", None)      # bench.py: CUDA events around the exposed all-reduce
                if ev is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                if compress:
                    half = flat.to(torch.bfloat16)
                    dist.all_reduce(half, op=dist.ReduceOp.SUM, group=grp)
                    flat.copy_(half)
                else:
                    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=grp)
                if ev is not None:
                    e1.record()
                    ev.append((e0, e1))
            for w in works:
                w.wait()
            scale /= dist.get_world_size(grp)
        if scale != 1.0:
            flat.mul_(scale)
        if ls != 1.0 and getattr(model, "train_check_overflow", True):
            # fp16 operands (opt-in; training defaults to bf16): the reference's "16-mixed" uses a dynamic GradScaler.  Here the
            # scale is static, so an overflow must not reach the optimiser: zero this step's gradients (AdamW then only
            # decays), halve the scale for the next step and say so.  One host sync per step, fp16 training only.
            if not bool(torch.isfinite(flat).all()):
                import warnings
                model.train_loss_scale = ls * 0.5
                flat.zero_()
                warnings.warn(f"maest_b200: fp16 gradient overflow, step skipped; loss scale {ls:g} -> {ls * 0.5:g}")
        grads = tuple(G[n].to(P_[n].dtype) if n in G else None for n in names)   # head_dist.* get no gradient ("mean" mode)
        return (None, None, None, None, None) + grads


def training_forward(model, mel: torch.Tensor, targets, mix: Optional[tuple] = None):
    """Train-mode forward + BCE loss with autograd support.  Host RNG draws follow the reference's order: the mixup draws
    (if any) are made by the caller BEFORE this function (helpers/mixup.py:6-7), then time offset / patchout here.

    distilled_type "mean": `targets` = y, returns (loss, logits).  distilled_type "separated" (teacher-student step,
    models/module.py:279-313): `targets` = (y, y_teacher), returns (loss, logits, logits_dist, loss_standard, loss_teacher)."""
    sep = model.distilled_type == "separated"
    if model.distilled_type not in ("mean", "separated"):
        raise ValueError(f"unknown distilled_type {model.distilled_type!r}")
    if sep != isinstance(targets, (tuple, list)):
        raise ValueError("distilled_type='separated' takes targets=(y, y_teacher); 'mean' takes a single target tensor")
    dev = model.cls_token.device
    if dev.type != "cuda":
        raise RuntimeError("maest_b200: training runs on CUDA (B200) only")
    mel = ops.h2d(mel, dev)
    targets = tuple(ops.h2d(t, dev) for t in targets) if sep else ops.h2d(targets, dev)
    if mel.dim() == 4:
        mel = mel[:, 0]
    if mix is not None:
        perm, lam = mix
        perm_d, lam_d = ops.h2d(perm.to(torch.int32), dev), ops.h2d(lam.float(), dev)
        mel = ops.mixup(mel, perm_d, lam_d)
        if sep:
            targets = tuple(ops.mixup(t, perm_d, lam_d) for t in targets)
        else:
            targets = ops.mixup(targets, perm_d, lam_d)
    if mel.shape[1] != 96:
        raise NotImplementedError(f"the B200 patch kernels take 96 mel bands, got {mel.shape[1]}")
    Fp, Tp = (mel.shape[1] - 16) // 10 + 1, (mel.shape[2] - 16) // 10 + 1
    t_offset, keep_f, keep_t, keep_seq = model._draw_patchout(Fp, Tp)
    keep_ft = ops.keep_ft_tensor(keep_f, keep_t, Fp, Tp, keep_seq, dev)
    params = [p for _, p in _flat_params(model)]
    return MaestTrainStep.apply(model, mel, targets, keep_ft, t_offset, *params)
