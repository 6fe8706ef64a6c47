"""GPU bring-up diagnostics: each check runs in its own process (a trapped kernel must not poison the rest).

    python tools/gpu_diag.py all            # run every check, one subprocess each, JSON lines -> gpurun_out/diag.jsonl
    python tools/gpu_diag.py gemm|attn|logmel|rowops|tokens|e2e|perf

Compares the CUDA kernels with torch fp32/fp64 math on the same inputs (unit level) and with the CPU oracle.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")

CHECKS = ["logmel", "gemm2", "rowops", "gemm", "attn", "tokens", "e2e", "perf", "bwdunits", "attnbwd", "train"]


def rel(a, b):
    import torch
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def emit(**kw):
    print("DIAG " + json.dumps(kw), flush=True)


def describe_mismatch(got, exp, name, tol):
    """Print where and how a 2-D result differs (helps to tell a layout/descriptor bug from a numeric one)."""
    import torch
    err = (got.double() - exp.double()).abs()
    bad = err > tol * exp.abs().max().double()
    nbad = int(bad.sum())
    info = dict(check=name, nbad=nbad, total=bad.numel(), max_err=float(err.max()), exp_absmax=float(exp.abs().max()),
                nan=int(torch.isnan(got.float()).sum()))
    if nbad:
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        info["bad_rows"] = rows[:16].tolist() + (["..."] if len(rows) > 16 else [])
        info["n_bad_rows"] = len(rows)
        info["bad_cols"] = cols[:16].tolist() + (["..."] if len(cols) > 16 else [])
        info["n_bad_cols"] = len(cols)
        r, c = int(rows[0]), int(cols[0])
        info["sample_got"] = got[r, c: c + 8].float().tolist()
        info["sample_exp"] = exp[r, c: c + 8].float().tolist()
    emit(**info)
    return nbad == 0


# ------------------------------------------------------------------------------------------------
def check_logmel():
    import torch
    from maest_b200 import ops, synth
    from oracle import maest_oracle as O
    for name, x in [("waveA_2x160000", synth.wave_a(2, 160000)), ("waveB_160000", synth.wave_b(160000)[None]),
                    ("waveA_3x48123", synth.wave_a(3, 48123, seed=3)), ("waveA_1x480000", synth.wave_a(1, 480000))]:
        ref = O.logmel(x, torch.float64)
        got = ops.logmel(x.cuda()).cpu()
        emit(check="logmel", case=name, shape=list(got.shape), max_abs=float((got.double() - ref).abs().max()),
             ok=bool((got.double() - ref).abs().max() < 2e-5))
    x1 = synth.wave_a(1, 160000)[0]
    got = ops.logmel(x1.cuda()).cpu()
    emit(check="logmel", case="1d", shape=list(got.shape),
         max_abs=float((got.double() - O.logmel(x1, torch.float64)).abs().max()))


def check_rowops():
    import torch
    from maest_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 768, generator=g) * 3 + 0.5
    w = torch.randn(768, generator=g)
    b = torch.randn(768, generator=g)
    ref = torch.nn.functional.layer_norm(x.double(), (768,), w.double(), b.double(), 1e-6)
    for dt in ("fp16", "bf16"):
        y, mean, rstd = ops.layernorm16(x.cuda(), w.cuda(), b.cuda(), 1e-6, dt, save_stats=True)
        emit(check="layernorm", dt=dt, rel=rel(y.cpu(), ref), mean_err=float((mean.cpu().double() - x.double().mean(1)).abs().max()))
    # block embedding
    B, N = 3, 77
    xs = torch.randn(B, N, 768, generator=g)
    emb = ops.block_embedding(xs.cuda(), B, N).cpu()
    ref = torch.cat([xs[:, 0], xs[:, 1], xs[:, 2:].double().mean(1).float()], 1)
    emit(check="block_embedding", max_abs=float((emb - ref).abs().max()))
    # pool/head
    from oracle import maest_oracle as O
    from maest_b200 import synth
    sd = synth.synth_state_dict(62, 400, depth=1)
    for sep in (False, True):
        lo, ld, ft = ops.pool_head(xs.cuda(), B, N, sd["norm.weight"].cuda(), sd["norm.bias"].cuda(), sd["head.0.weight"].cuda(),
                                   sd["head.0.bias"].cuda(), sd["head.1.weight"].cuda(), sd["head.1.bias"].cuda(),
                                   sd["head_dist.weight"].cuda(), sd["head_dist.bias"].cuda(), separated=sep)
        xn = O.layer_norm(xs.double(), sd["norm.weight"].double(), sd["norm.bias"].double(), 1e-6)
        outs = O.head(xn[:, 0], xn[:, 1], sd, "separated" if sep else "mean")
        if sep:
            emit(check="pool_head", sep=sep, rel_cls=rel(lo.cpu(), outs[0]), rel_dist=rel(ld.cpu(), outs[1]), rel_feats=rel(ft.cpu(), outs[2]))
        else:
            emit(check="pool_head", sep=sep, rel_logits=rel(lo.cpu(), outs[0]), rel_feats=rel(ft.cpu(), outs[1]))
    c = ops.cast16(torch.arange(1003, dtype=torch.float32).cuda() * 0.37, "bf16")
    emit(check="cast16", ok=bool(torch.equal(c.cpu(), (torch.arange(1003, dtype=torch.float32) * 0.37).bfloat16())))


def check_gemm():
    import torch
    from maest_b200 import _lib, ops
    g = torch.Generator().manual_seed(1)
    # probe: W = identity -> C must equal A (reveals swizzle / descriptor / lane-mapping problems exactly)
    for dt in (torch.float16, torch.bfloat16):
        A = (torch.randn(256, 256, generator=g)).to(dt).cuda()
        W = torch.eye(256).to(dt).cuda()
        C = ops.linear(A, W, None, _lib.EPI_STORE32)
        torch.cuda.synchronize()
        describe_mismatch(C.cpu(), A.float().cpu(), f"gemm_identity_{dt}", 1e-6)
    cases = [(300, 256, 64), (128, 256, 768), (1000, 2304, 768), (257, 768, 3072), (4000, 3072, 768), (558 * 3, 768, 256)]
    for dt in (torch.float16, torch.bfloat16):
        for (M, N, K) in cases:
            A = (torch.randn(M, K, generator=g) * 0.5).to(dt).cuda()
            W = (torch.randn(N, K, generator=g) * 0.05).to(dt).cuda()
            bias = torch.randn(N, generator=g).cuda()
            ref = A.double() @ W.double().t() + bias.double()
            for epi, nm in ((_lib.EPI_STORE32, "store32"), (_lib.EPI_STORE16, "store16"), (_lib.EPI_GELU16, "gelu16"),
                            (_lib.EPI_RESID32, "resid32")):
                if epi == _lib.EPI_RESID32:
                    x0 = torch.randn(M, N, generator=g).cuda()
                    out = x0.clone()
                    ops.linear(A, W, bias, epi, resid=out, out=out)
                    exp = x0.double() + ref
                elif epi == _lib.EPI_GELU16:
                    out = ops.linear(A, W, bias, epi)
                    exp = torch.nn.functional.gelu(ref)
                else:
                    out = ops.linear(A, W, bias, epi)
                    exp = ref
                torch.cuda.synchronize()
                r = rel(out, exp)
                ok = r < (2e-3 if dt == torch.float16 else 8e-3) if epi in (_lib.EPI_STORE16, _lib.EPI_GELU16) else r < 1e-5
                emit(check="gemm", dt=str(dt), M=M, N=N, K=K, epi=nm, rel=r, ok=bool(ok))
                if not ok and epi == _lib.EPI_STORE32:
                    describe_mismatch(out.cpu(), exp.float().cpu(), f"gemm_{M}x{N}x{K}", 1e-3)
    # row remap + addend (patch-embed epilogue)
    M, N, K, P = 6 * 50, 768, 256, 50
    A = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).half().cuda()
    add = torch.randn(P, N, generator=g).cuda()
    out = torch.zeros(6, 2 + P, N).cuda()
    ops.linear(A, W, None, _lib.EPI_STORE32, out=out.view(-1, N), addend=add, rows_per_group=P, group_stride=2 + P, row_offset=2)
    exp = (A.double() @ W.double().t()).view(6, P, N) + add.double()
    emit(check="gemm_remap", rel=rel(out[:, 2:], exp), rows01_untouched=bool((out[:, :2] == 0).all()))


def check_gemm2():
    """CTA-pair (cta_group::2) GEMM: correctness vs fp64 and speed vs the 1-CTA kernel."""
    import torch
    from maest_b200 import _lib, ops
    ops.set_gemm_mode(True)
    check_gemm()
    M = 64 * 1685
    for (Nn, K, epi, nm) in [(2304, 768, _lib.EPI_STORE16, "qkv"), (768, 768, _lib.EPI_RESID32, "proj"),
                             (3072, 768, _lib.EPI_GELU16, "fc1"), (768, 3072, _lib.EPI_RESID32, "fc2")]:
        A = (torch.randn(M, K, device="cuda") * 0.5).half()
        W = (torch.randn(Nn, K, device="cuda") * 0.05).half()
        bias = torch.randn(Nn, device="cuda")
        out = torch.empty(M, Nn, device="cuda", dtype=torch.float16 if epi in (_lib.EPI_STORE16, _lib.EPI_GELU16) else torch.float32)
        res = {}
        for mode in (False, True):
            ops.set_gemm_mode(mode)
            fn = lambda: ops.linear(A, W, bias, epi, out=out, resid=out if epi == _lib.EPI_RESID32 else None)  # noqa: E731
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(6):
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            res["pair" if mode else "single"] = min(ts)
        emit(check="gemm2_perf", kernel=nm, ms_single=res["single"], ms_pair=res["pair"], tflops_pair=2.0 * M * Nn * K / res["pair"] / 1e9)
    ops.set_gemm_mode(None)


def check_attn():
    import torch
    from maest_b200 import ops
    g = torch.Generator().manual_seed(2)
    for dt in (torch.float16, torch.bfloat16):
        for (B, N) in [(1, 128), (2, 100), (2, 560), (1, 1685), (3, 866)]:
            qkv = (torch.randn(B * N, 2304, generator=g)).to(dt).cuda()
            q, k, v = qkv.view(B, N, 3, 12, 64).permute(2, 0, 3, 1, 4).double()
            s = (q @ k.transpose(-1, -2)) * 0.125
            ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, 768)
            for variant in (0, 1, 2):
                try:
                    o = ops.attention(qkv, B, N, 12, variant)
                    torch.cuda.synchronize()
                    r = rel(o, ref)
                    emit(check="attn", dt=str(dt), B=B, N=N, variant=variant, rel=r, ok=bool(r < (3e-3 if dt == torch.float16 else 1.5e-2)),
                         nan=int(torch.isnan(o.float()).sum()))
                    if r > 0.05:
                        describe_mismatch(o.float().cpu(), ref.float().cpu(), f"attn_v{variant}_{B}x{N}", 2e-2)
                except Exception as e:  # noqa: BLE001
                    emit(check="attn", dt=str(dt), B=B, N=N, variant=variant, error=str(e)[:300])
                    raise
    # large-magnitude scores: exercises the lazy rescale path
    B, N = 1, 700
    qkv = (torch.randn(B * N, 2304, generator=g) * 4).half().cuda()
    q, k, v = qkv.view(B, N, 3, 12, 64).permute(2, 0, 3, 1, 4).double()
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).transpose(1, 2).reshape(B * N, 768)
    for variant in (0, 1, 2):
        o = ops.attention(qkv, B, N, 12, variant)
        emit(check="attn_sharp", variant=variant, rel=rel(o, ref))


def check_tokens():
    import torch
    from maest_b200 import get_maest, synth
    from oracle import maest_oracle as O
    sd = synth.synth_state_dict(62, 400, depth=12)
    m = get_maest(arch="discogs-maest-10s-pw-129e", pretrained=False)
    m.load_state_dict(sd, strict=False)
    m = m.cuda().eval()
    mel = O.logmel(synth.wave_a(2, 160000), torch.float32)
    ref = O.patch_tokens(mel.double(), sd)
    got = m.tokens_from_mel(mel.cuda()).cpu()
    emit(check="tokens", shape=list(got.shape), rel=rel(got, ref), rel_rows01=rel(got[:, :2], ref[:, :2]),
         max_abs=float((got.double() - ref).abs().max()))
    # patchout (kept columns/rows passed explicitly through the ops layer)
    from maest_b200 import ops
    keep_t = list(range(0, 62, 3))
    keep_f = [0, 2, 3, 7]
    kft = ops.keep_ft_tensor(keep_f, keep_t, 9, 62, None, "cuda")
    got = ops.patch_tokens(mel.contiguous().cuda(), m._weight16("patch_embed.proj", m.patch_embed.proj.weight), m.patch_embed.proj.bias.detach(),
                           m.freq_new_pos_embed.detach().reshape(768, -1).contiguous(), m.time_new_pos_embed.detach().reshape(768, -1).contiguous(),
                           m.cls_token.detach().reshape(-1), m.dist_token.detach().reshape(-1), m.new_pos_embed.detach().reshape(2, 768),
                           keep_ft=kft, t_offset=0).cpu()
    ref = O.patch_tokens(mel.double(), sd, keep_t=keep_t, keep_f=keep_f)
    emit(check="tokens_patchout", shape=list(got.shape), rel=rel(got, ref))


def check_e2e():
    import numpy as np
    import torch
    from maest_b200 import get_maest, synth
    from oracle import maest_oracle as O
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "c2.npz")))
    sd = synth.synth_state_dict(62, 400)
    x = synth.wave_a(2, 160000)
    for dt in ("fp16", "bf16"):
        for variant in (0, 1, 2):
            m = get_maest(arch="discogs-maest-10s-pw-129e", pretrained=False, op_dtype=dt)
            m.attn_variant = variant
            m.load_state_dict(sd, strict=False)
            m = m.cuda().eval()
            with torch.no_grad():
                lo, em = m(x.cuda())
                e6 = m(x.cuda(), transformer_block=6)[1]
            emit(check="e2e", dt=dt, variant=variant, rel_logits=rel(lo.cpu(), torch.tensor(gold["logits"])),
                 rel_emb=rel(em.cpu(), torch.tensor(gold["emb"])), rel_emb6=rel(e6.cpu(), torch.tensor(gold["emb_block6"])),
                 max_abs_logits=float((lo.cpu() - torch.tensor(gold["logits"])).abs().max()))
    # block-by-block drift (fp16, variant 0)
    m = get_maest(arch="discogs-maest-10s-pw-129e", pretrained=False)
    m.load_state_dict(sd, strict=False)
    m = m.cuda().eval()
    rows = list(gold["row_probe"])
    from maest_b200 import ops
    mel = ops.logmel(x.cuda())
    tok = m.tokens_from_mel(mel)
    B, N, _ = tok.shape
    table = m._blocks_ctypes()
    emit(check="e2e_tokens", rel=rel(tok[:, rows].cpu(), torch.tensor(gold["tokens_probe"])))
    for i in range(12):
        ops.encoder(tok.view(B * N, 768), B, N, (type(table[0]) * 1)(table[i]), 1, False, "fp16", 0)
        emit(check="e2e_block", block=i, rel=rel(tok[:, rows].cpu(), torch.tensor(gold[f"block{i}_probe"])))


def check_perf():
    """First timing pass (not the bench): per-kernel CUDA-event times at the config-3 shapes."""
    import torch
    from maest_b200 import _lib, ops
    dev = "cuda"
    B, N = 64, 1685
    M = B * N

    def timeit(fn, n=5):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return min(ts)

    x = torch.randn(M, 768, device=dev)
    w = torch.ones(768, device=dev)
    bz = torch.zeros(768, device=dev)
    t = timeit(lambda: ops.layernorm16(x, w, bz, 1e-6, "fp16"))
    emit(check="perf", kernel="layernorm", ms=t, gbs=M * 768 * 6 / t / 1e6)
    for dt in (torch.float16, torch.bfloat16):
        for (Nn, K, epi, nm) in [(2304, 768, _lib.EPI_STORE16, "qkv"), (768, 768, _lib.EPI_RESID32, "proj"),
                                 (3072, 768, _lib.EPI_GELU16, "fc1"), (768, 3072, _lib.EPI_RESID32, "fc2")]:
            A = (torch.randn(M, K, device=dev) * 0.5).to(dt)
            W = (torch.randn(Nn, K, device=dev) * 0.05).to(dt)
            bias = torch.randn(Nn, device=dev)
            out = torch.empty(M, Nn, device=dev, dtype=dt if epi in (_lib.EPI_STORE16, _lib.EPI_GELU16) else torch.float32)
            t = timeit(lambda: ops.linear(A, W, bias, epi, out=out, resid=out if epi == _lib.EPI_RESID32 else None))
            emit(check="perf", kernel="gemm_" + nm, dt=str(dt), ms=t, tflops=2.0 * M * Nn * K / t / 1e9)
            tt = timeit(lambda: torch.matmul(A, W.t()))
            emit(check="perf", kernel="cublas_" + nm, dt=str(dt), ms=tt, tflops=2.0 * M * Nn * K / tt / 1e9)
            del A, W, out
        qkv = torch.randn(M, 2304, device=dev).to(dt)
        for variant in (0, 1, 2):
            t = timeit(lambda: ops.attention(qkv, B, N, 12, variant))
            emit(check="perf", kernel="attention", dt=str(dt), variant=variant, ms=t, tflops=4.0 * B * 12 * N * N * 64 / t / 1e9)
        del qkv
    wav = torch.rand(64, 480000, device=dev) * 2 - 1
    t = timeit(lambda: ops.logmel(wav))
    emit(check="perf", kernel="logmel", ms=t, gbs=(64 * 480000 * 4 + 64 * 96 * 1876 * 4) / t / 1e6)



def check_bwdunits():
    """Backward building blocks vs torch autograd / fp64 math."""
    import torch
    from maest_b200 import _lib, ops
    g = torch.Generator().manual_seed(5)
    for dt in (torch.float16, torch.bfloat16):
        tol = 2e-3 if dt == torch.float16 else 1e-2
        Mt, Nout, Kin = 1732, 768, 3072          # tokens, out features, in features (fc2-like)
        dY = (torch.randn(Mt, Nout, generator=g) * 0.1).to(dt).cuda()
        X = (torch.randn(Mt, Kin, generator=g) * 0.5).to(dt).cuda()
        W = (torch.randn(Nout, Kin, generator=g) * 0.05).to(dt).cuda()
        # input gradient: dX = dY @ W  (B MN-major)
        dX = ops.gemm(dY, W, _lib.EPI_STORE32, Mt, Kin, Nout, b_mn=True)
        emit(check="dgrad_store32", dt=str(dt), rel=rel(dX, dY.double() @ W.double()))
        dX16 = ops.gemm(dY, W, _lib.EPI_STORE16, Mt, Kin, Nout, b_mn=True)
        emit(check="dgrad_store16", dt=str(dt), rel=rel(dX16, dY.double() @ W.double()), ok=bool(rel(dX16, dY.double() @ W.double()) < tol))
        # GELU backward epilogue
        upre = (torch.randn(Mt, Kin, generator=g) * 1.5).to(dt).cuda()
        up = upre.double().requires_grad_(True)
        torch.nn.functional.gelu(up).backward(dY.double() @ W.double())
        dU = ops.gemm(dY, W, _lib.EPI_GELUBWD16, Mt, Kin, Nout, b_mn=True, aux16=upre)
        emit(check="dgrad_gelubwd16", dt=str(dt), rel=rel(dU, up.grad))
        # weight gradient: dW = dY^T @ X  (both MN-major, split-K atomics)
        for splits in (1, 3, ops.wgrad_splits(Nout, Kin, Mt)):
            dW = torch.zeros(Nout, Kin, device="cuda")
            ops.gemm(dY, X, _lib.EPI_ATOMIC32, Nout, Kin, Mt, a_mn=True, b_mn=True, out=dW, k_splits=splits)
            r = rel(dW, dY.double().t() @ X.double())
            emit(check="wgrad_atomic32", dt=str(dt), splits=splits, rel=r, ok=bool(r < 1e-5))
            if r > 1e-3:
                describe_mismatch(dW.cpu(), (dY.double().t() @ X.double()).float().cpu(), "wgrad", 1e-3)
        # GELU16 with saved pre-activation
        A = (torch.randn(300, 768, generator=g) * 0.5).to(dt).cuda()
        W1 = (torch.randn(3072, 768, generator=g) * 0.05).to(dt).cuda()
        b1 = torch.randn(3072, generator=g).cuda()
        pre = torch.empty(300, 3072, device="cuda", dtype=dt)
        u = ops.gemm(A, W1, _lib.EPI_GELU16, 300, 3072, 768, bias=b1, aux16=pre)
        ref = A.double() @ W1.double().t() + b1.double()
        emit(check="gelu16_aux", dt=str(dt), rel_pre=rel(pre, ref), rel_u=rel(u, torch.nn.functional.gelu(ref)))
    # LayerNorm backward
    rows = 1000
    x = (torch.randn(rows, 768, generator=g) * 2 + 0.3)
    gam = torch.randn(768, generator=g)
    dy = torch.randn(rows, 768, generator=g)
    dx0 = torch.randn(rows, 768, generator=g)
    xd = x.double().requires_grad_(True)
    gd = gam.double().requires_grad_(True)
    bd = torch.zeros(768, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.layer_norm(xd, (768,), gd, bd, 1e-6).backward(dy.double())
    _, mean, rstd = ops.layernorm16(x.cuda(), gam.cuda(), torch.zeros(768).cuda(), 1e-6, "fp16", save_stats=True)
    dx = dx0.clone().cuda()
    dg, db = torch.zeros(768).cuda(), torch.zeros(768).cuda()
    dx16 = torch.empty(rows, 768, device="cuda", dtype=torch.float16)
    ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, gam.cuda(), dx, dg, db, "fp16", dx16=dx16)
    emit(check="layernorm_bwd", rel_dx=rel(dx.cpu(), dx0.double() + xd.grad), rel_dgamma=rel(dg.cpu(), gd.grad), rel_dbeta=rel(db.cpu(), bd.grad),
         rel_dx16=rel(dx16.cpu(), dx0.double() + xd.grad))
    # colsum / cast_rows16 / mixup / bce
    z = torch.randn(777, 2304, generator=g)
    out = torch.zeros(2304).cuda()
    ops.colsum(z.half().cuda(), out)
    emit(check="colsum", rel=rel(out.cpu(), z.half().double().sum(0)))
    src = torch.randn(3 * 52, 768, generator=g)
    c = ops.cast_rows16(src.cuda(), 3 * 50, "bf16", rows_per_group=50, group_stride=52, row_offset=2)
    emit(check="cast_rows16", ok=bool(torch.equal(c.cpu(), src.view(3, 52, 768)[:, 2:].reshape(150, 768).bfloat16())))
    xm = torch.randn(4, 1, 96, 100, generator=g).half()
    perm = torch.tensor([2, 0, 3, 1])
    lam = torch.tensor([0.9, 0.6, 0.75, 0.51])
    mx = ops.mixup(xm.cuda(), perm.cuda(), lam.cuda())
    refm = xm.float() * lam.view(4, 1, 1, 1) + xm.float()[perm] * (1 - lam.view(4, 1, 1, 1))
    emit(check="mixup", max_abs=float((mx.cpu() - refm).abs().max()))
    lz = torch.randn(8, 400, generator=g) * 3
    ly = (torch.rand(8, 400, generator=g) > 0.9).float()
    loss, dz = ops.bce_logits(lz.cuda(), ly.cuda())
    lzd = lz.double().requires_grad_(True)
    lref = torch.nn.functional.binary_cross_entropy_with_logits(lzd, ly.double())
    lref.backward()
    emit(check="bce", loss_err=abs(float(loss) - float(lref)), rel_dz=rel(dz.cpu(), lzd.grad))


def check_attnbwd():
    import torch
    from maest_b200 import ops
    g = torch.Generator().manual_seed(6)
    for dt in (torch.float16, torch.bfloat16):
        for (B, N) in [(1, 128), (2, 100), (2, 866), (1, 300)]:
            qkv = torch.randn(B * N, 2304, generator=g).to(dt).cuda()
            d_o = (torch.randn(B * N, 768, generator=g) * 0.1).to(dt).cuda()
            o, lse = ops.attention(qkv, B, N, 12, 0, save_lse=True)
            qd = qkv.double().requires_grad_(True)
            q, k, v = qd.view(B, N, 3, 12, 64).permute(2, 0, 3, 1, 4)
            s = (q @ k.transpose(-1, -2)) * 0.125
            ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, 768)
            ref.backward(d_o.double())
            lse_ref = (torch.logsumexp(s, -1) * 1.4426950408889634)          # [B,12,N] in log2 units
            dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, N, 12)
            torch.cuda.synchronize()
            gq = qd.grad
            emit(check="attn_bwd", dt=str(dt), B=B, N=N, lse_err=float((lse.double() - lse_ref).abs().max()),
                 rel_dq=rel(dqkv[:, :768], gq[:, :768]), rel_dk=rel(dqkv[:, 768:1536], gq[:, 768:1536]),
                 rel_dv=rel(dqkv[:, 1536:], gq[:, 1536:]), nan=int(torch.isnan(dqkv.float()).sum()))
            if rel(dqkv, gq) > 0.05:
                describe_mismatch(dqkv.float().cpu(), gq.float().cpu(), f"attn_bwd_{B}x{N}", 5e-2)


def check_train():
    """Full training step (mixup -> forward with patchout -> BCE -> backward) vs the reference's golden c4."""
    import numpy as np
    import torch
    from maest_b200 import get_maest, synth
    from maest_b200.train import training_forward
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "c4.npz")))
    for dt in ("fp16", "bf16"):
        m = get_maest(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=400, input_f=96, input_t=1875, s_patchout_t=90, op_dtype=dt)
        m.load_state_dict(synth.synth_state_dict(187, 400, seed=0), strict=False)
        m = m.cuda().train()
        x, y = synth.train_batch(2)
        torch.manual_seed(1)
        np.random.seed(1)
        from maest_b200.module import my_mixup
        mix = my_mixup(2, 0.3)
        loss, logits = training_forward(m, x.cuda(), y.cuda(), mix)
        loss.backward()
        torch.cuda.synchronize()
        res = dict(check="train", dt=dt, loss=float(loss), loss_ref=float(gold["loss"]), loss_err=abs(float(loss) - float(gold["loss"])))
        grads = {n: p.grad for n, p in m.named_parameters()}
        for k in ["cls_token", "time_new_pos_embed", "freq_new_pos_embed", "patch_embed.proj.bias", "blocks.0.norm1.weight",
                  "blocks.0.attn.qkv.bias", "blocks.5.mlp.fc1.bias", "blocks.11.attn.proj.bias", "norm.weight", "head.0.bias", "head.1.bias"]:
            res["g." + k] = round(rel(grads[k].cpu(), torch.tensor(gold["grad." + k])), 6)
        for k in ["patch_embed.proj.weight", "blocks.0.attn.qkv.weight", "blocks.5.mlp.fc1.weight", "blocks.11.mlp.fc2.weight", "head.1.weight"]:
            gk = grads[k]
            res["gs." + k] = round(rel(gk.reshape(gk.shape[0], -1)[::37, ::29].cpu(), torch.tensor(gold["grad." + k + ".sub"])), 6)
            res["gn." + k] = round(float(gk.double().norm()) / float(gold["gnorm." + k]), 6)
        res["head_dist_none"] = grads["head_dist.weight"] is None
        emit(**res)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        os.makedirs(OUT, exist_ok=True)
        names = sys.argv[2:] or CHECKS
        with open(os.path.join(OUT, "diag.jsonl"), "a") as f:
            for name in names:
                t0 = time.time()
                try:
                    res = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True, timeout=420)
                    rc, so, se = res.returncode, res.stdout, res.stderr
                except subprocess.TimeoutExpired as e:
                    rc, so, se = -999, (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT"
                print(f"=== {name}: rc={rc} ({time.time() - t0:.1f}s)")
                for line in so.splitlines():
                    if line.startswith("DIAG "):
                        print("  " + line[5:])
                        f.write(line[5:] + "\n")
                    elif line.strip():
                        print("  | " + line[:300])
                if rc != 0:
                    tail = se.strip().splitlines()[-25:]
                    print("  stderr tail:\n    " + "\n    ".join(tail))
                    f.write(json.dumps(dict(check=name, rc=rc, stderr=tail[-8:])) + "\n")
                f.flush()
        return
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    globals()["check_" + what]()


if __name__ == "__main__":
    main()
