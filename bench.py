#!/usr/bin/env python
"""bench.py — clips/sec of the MAEST hot path (waveform -> log-mel -> tokens -> 12 blocks -> logits) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--arch ...] [--batch B]

A "step" is one forward pass over one batch of synthetic clips (BASELINE.json configs[2]:
discogs-maest-30s-pw-129e, batch 64, [64, 480000] 16 kHz waveforms -> 96x1876 mel -> 1685 tokens).  One process per
GPU (torchrun for N>1); clips shard across ranks with no data-path collective (weak scaling: 64 clips per GPU).
Rank 0 prints ONE JSON line.  `value` = whole-job clips/s with inputs resident in HBM; `e2e` = the same metric through
the public API `model(waveform)` with the waveform in pinned host memory (H2D inside the timed region, logits read
back to the host every step).  `roofline` describes the dominant kernel (the tcgen05 GEMM), timed with CUDA events.

`--impl reference` times the CPU restatement of the reference (oracle/maest_oracle.py, torch CPU ops, all host
threads) on a bounded sample of the same workload; the reference itself is pure Python under /root/reference, which does
not exist on the GPU box, and the oracle is pinned against it by tests/golden.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH_T = {"discogs-maest-30s-pw-129e": (480000, 187), "discogs-maest-10s-pw-129e": (160000, 62)}
EMBED, DEPTH, HEADS, MLP = 768, 12, 12, 3072


def flops_per_clip(N: int, P: int, C: int = 400):
    lin = DEPTH * 2 * N * (EMBED * 3 * EMBED + EMBED * EMBED + 2 * EMBED * MLP)
    att = DEPTH * 4 * N * N * EMBED
    patch = 2 * P * 256 * EMBED
    return dict(linear=lin, attention=att, patch=patch, head=2 * EMBED * C, total=lin + att + patch + 2 * EMBED * C)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm))


def cpu_forward_fn(arch: str):
    """The CPU implementation of the path that the baseline legs time, and what it is.

    Preferred: the UNMODIFIED reference (palonso/MAEST `maest` package pip-installed into baseline/_ref, see DESIGN.md section 1;
    its two absent pure-Python imports, sacred and timm, are stubbed in memory by tests/golden/ref_loader.py) -> kind "reference".
    Fallback when baseline/_ref did not travel: oracle/maest_oracle.py, the torch-CPU restatement pinned to it -> kind "port"."""
    import torch
    from maest_b200 import synth

    S, grid_t = ARCH_T[arch]
    sd = synth.synth_state_dict(grid_t, 400, seed=0)
    ref_dir = os.environ.get("MAEST_REF_INSTALL", os.path.join(ROOT, "baseline", "_ref"))
    if os.path.isfile(os.path.join(ref_dir, "maest", "maest.py")) and not os.environ.get("MAEST_BENCH_FORCE_PORT"):
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
            sys.path.insert(0, ref_dir)
            import ref_loader
            ref_loader.install_stubs(with_lightning=True)
            import maest as ref_pkg
            net = ref_pkg.get_maest(arch=arch, pretrained=False)
            net.load_state_dict(sd, strict=False)
            net.eval()
            return (lambda x: net(x)[0]), "reference", "unmodified reference (baseline/_ref: maest.get_maest(arch).forward, torch CPU fp32)"
        except Exception as e:  # noqa: BLE001
            print(f"bench.py: baseline/_ref unusable ({type(e).__name__}: {e}); timing the oracle port instead", file=sys.stderr)
    from oracle import maest_oracle as O
    return (lambda x: O.forward(x, sd, img_t=(S // 256), dtype=torch.float32)), "port", \
        "oracle/maest_oracle.py (torch-CPU fp32 restatement of the reference)"


def cpu_baseline_clips_per_s(arch: str, clips: int, repeats: int = 1):
    """Time the reference's CPU path on `clips` clips of the workload (fp32, all host threads)."""
    import torch
    from maest_b200 import synth

    S, _ = ARCH_T[arch]
    torch.set_num_threads(os.cpu_count() or 1)
    fwd, kind, what = cpu_forward_fn(arch)
    x = synth.wave_a(clips, S)
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            fwd(x.clone())
            best = min(best, time.perf_counter() - t0)
    return clips / best, torch.get_num_threads(), kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)   # torchrun pins this to 1; the CPU arm uses every host core
    os.environ["MKL_NUM_THREADS"] = str(os.cpu_count() or 1)
    import torch
    from maest_b200 import synth

    S, grid_t = ARCH_T[args.arch]
    clips = args.ref_clips
    torch.set_num_threads(os.cpu_count() or 1)
    fwd, kind, what = cpu_forward_fn(args.arch)
    x = synth.wave_a(clips, S)
    with torch.no_grad():
        for _ in range(args.warmup):
            fwd(x.clone())           # (the reference unsqueezes its input in place)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fwd(x.clone())
        dt = time.perf_counter() - t0
    val = clips * args.steps / dt
    N = 2 + 9 * ((S // 256 + 1 - 16) // 10 + 1)
    line = dict(metric="clips/sec", value=val, unit="clips/s", impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"{args.arch} inference, waveform [{clips},{S}] -> logits, N={N} tokens (bounded CPU sample of the batch-64 workload)"),
                cpu_baseline=dict(value=val, unit="clips/s", cores=torch.get_num_threads(), kind=kind,
                                  sample=f"{clips} clips/step x {args.steps} steps, {what}"),
                e2e=dict(value=val, unit="clips/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def kernel_breakdown(model, wav, iters: int = 3):
    """Per-kernel CUDA-event times of one step, composed at the ops level exactly like MAEST.forward.  Every activation buffer is
    allocated ONCE before the timed iterations (an allocation inside an event pair can stall the stream and would be billed to
    the kernel that follows it)."""
    import torch
    from maest_b200 import _lib, ops
    acc = {}

    def timed(name, fn):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        out = fn()
        b.record()
        acc.setdefault(name, []).append((a, b))
        return out

    dt = model.op_dtype
    t16 = torch.float16 if ops.op_dtype_code(dt) == ops.F16 else torch.bfloat16
    dev = wav.device
    fuse = bool(getattr(model, "fuse_ln", False))
    nb = len(model.blocks)
    bufs = None
    warm = 2
    for it in range(iters + warm):
        if it == warm:
            torch.cuda.synchronize()
            acc.clear()                      # the first iterations warm torch's caching allocator and create the reusable buffers
        mel = timed("logmel", lambda: ops.logmel(wav))
        tok = timed("patch_tokens", lambda: model.tokens_from_mel(mel))
        B, N, _ = tok.shape
        M = B * N
        x = tok.view(M, EMBED)
        if bufs is None:
            bufs = dict(h=torch.empty((M, EMBED), device=dev, dtype=t16), qkv=torch.empty((M, 3 * EMBED), device=dev, dtype=t16),
                        o=torch.empty((M, EMBED), device=dev, dtype=t16), u=torch.empty((M, MLP), device=dev, dtype=t16))
        h, qkv, o, u = bufs["h"], bufs["qkv"], bufs["o"], bufs["u"]
        stats = parts = None
        if fuse:
            parts = torch.empty((EMBED // 128, M, 4), device=dev, dtype=torch.float32)
        for i, blk in enumerate(model.blocks):
            w = {k: model._weight16(f"blocks.{i}.{k}", p) for k, p in (("qkv", blk.attn.qkv.weight), ("proj", blk.attn.proj.weight),
                                                                       ("fc1", blk.mlp.fc1.weight), ("fc2", blk.mlp.fc2.weight))}
            if fuse and i > 0:      # norm1 was folded into the previous block's fc2 epilogue (h, stats) and finishes in this one
                wg, bf = model._ln_fold(f"blocks.{i}.qkv", w["qkv"], blk.norm1, blk.attn.qkv.bias)
                timed("gemm_qkv", lambda: ops.linear_ln(h, w["qkv"], bf, _lib.EPI_STORE16_LN, stats, wg, out=qkv))
            else:
                timed("layernorm", lambda: ops.layernorm16(x, blk.norm1.weight.detach(), blk.norm1.bias.detach(), 1e-6, dt, out=h))
                timed("gemm_qkv", lambda: ops.linear(h, w["qkv"], blk.attn.qkv.bias.detach(), _lib.EPI_STORE16, out=qkv))
            timed("attention", lambda: ops.attention(qkv, B, N, HEADS, model.attn_variant, out=o))
            if fuse:
                wg, bf = model._ln_fold(f"blocks.{i}.fc1", w["fc1"], blk.norm2, blk.mlp.fc1.bias)
                timed("gemm_proj", lambda: ops.linear_ln(o, w["proj"], blk.attn.proj.bias.detach(), _lib.EPI_RESID32_LN, parts, blk.norm2.weight.detach(),
                                                         out=x, resid=x, out16b=h))
                stats = timed("layernorm", lambda: ops.ln_finalize(parts))
                timed("gemm_fc1", lambda: ops.linear_ln(h, w["fc1"], bf, _lib.EPI_GELU16_LN, stats, wg, out=u))
            else:
                timed("gemm_proj", lambda: ops.linear(o, w["proj"], blk.attn.proj.bias.detach(), _lib.EPI_RESID32, resid=x, out=x))
                timed("layernorm", lambda: ops.layernorm16(x, blk.norm2.weight.detach(), blk.norm2.bias.detach(), 1e-6, dt, out=h))
                timed("gemm_fc1", lambda: ops.linear(h, w["fc1"], blk.mlp.fc1.bias.detach(), _lib.EPI_GELU16, out=u))
            if fuse and i + 1 < nb:
                timed("gemm_fc2", lambda: ops.linear_ln(u, w["fc2"], blk.mlp.fc2.bias.detach(), _lib.EPI_RESID32_LN, parts,
                                                        model.blocks[i + 1].norm1.weight.detach(), out=x, resid=x, out16b=h))
                stats = timed("layernorm", lambda: ops.ln_finalize(parts))
            else:
                timed("gemm_fc2", lambda: ops.linear(u, w["fc2"], blk.mlp.fc2.bias.detach(), _lib.EPI_RESID32, resid=x, out=x))
        timed("pool_head", lambda: ops.pool_head(tok, B, N, model.norm.weight.detach(), model.norm.bias.detach(), model.head[0].weight.detach(),
                                                 model.head[0].bias.detach(), model.head[1].weight.detach(), model.head[1].bias.detach()))
    torch.cuda.synchronize()
    out = {}
    for k, evs in acc.items():
        ms = [a.elapsed_time(b) for a, b in evs]
        per_iter = len(evs) // iters
        out[k] = dict(launches_per_step=per_iter, ms_per_launch=sum(ms) / len(ms), ms_per_step=sum(ms) / iters)
    return out


def run_train(args):
    """--mode train: only the training leg (see train_leg), printed as its own JSON line."""
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = train_leg(args, args.steps, args.warmup, args.batch)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def train_leg(args, steps: int, warmup: int, B: int):
    """configs[3]: maest_30s_from_passt_pretrain training step (mel [B,1,96,1875] fp16 in, s_patchout_t=90 -> 866 tokens,
    mixup 0.3, BCE), bf16 operands, fwd + bwd + gradient all-reduce (N>1) + AdamW step inside the timed region.
    The process group (N>1) is the caller's.  Returns the JSON line (rank 0) / None."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from maest_b200 import get_maest, synth
    from maest_b200.module import Module

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    op = "bf16" if args.op_dtype == "fp16" and not os.environ.get("MAEST_TRAIN_FP16") else args.op_dtype
    net = get_maest(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=400, input_f=96, input_t=1875, s_patchout_t=90, op_dtype=op)
    net.load_state_dict(synth.synth_state_dict(187, 400, seed=0), strict=False)
    if world > 1:
        # ONE flat NCCL all-reduce of the gradient buffer after the last backward kernel: fp32 (the reference's DDP semantics) unless
        # MAEST_ALLREDUCE=bf16 asks for the compressed variant
        net.grad_allreduce = "bf16" if os.environ.get("MAEST_ALLREDUCE") == "bf16" else True
    net.allreduce_events = []              # (start, end) CUDA events around that all-reduce, one pair per step (train.py)
    mod = Module(net=net, mixup_alpha=0.3, do_swa=False).to(dev).train()
    from maest_b200.optim import FusedAdamW
    opt = FusedAdamW(mod.parameters(), lr=2e-5, weight_decay=1e-4)      # one launch for all 152 parameter tensors
    torch.manual_seed(1 + rank)
    np.random.seed(1 + rank)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    x = (0.5 * torch.randn(B, 1, 96, 1875, generator=g, device=dev)).half()
    y = (torch.rand(B, 400, generator=g, device=dev) > 0.99).half()
    batch = (x, ["clip"] * B, y)
    n_grad = [0]

    def step():
        opt.zero_grad(set_to_none=True)
        loss = mod.training_step(batch, 0)
        loss.backward()
        if not n_grad[0]:
            n_grad[0] = sum(p.numel() for p in mod.parameters() if p.grad is not None)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    net.allreduce_events.clear()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ar_mode = getattr(net, "grad_allreduce", None)
    ar_ms = [a.elapsed_time(b) for a, b in net.allreduce_events]
    ar_ms = sum(ar_ms) / len(ar_ms) if ar_ms else 0.0
    if world > 1:
        t = torch.tensor([ar_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ar_ms = float(t.item())
    line = None
    if rank == 0:
        N, P = 866, 864
        fl = flops_per_clip(N, P)
        peaks = measured_peaks()
        value = B * world * steps / (ms / 1e3)
        tf = value / world * 3 * fl["total"] / 1e12
        line = dict(metric="clips/sec (training step)", value=value, unit="clips/s", n_gpus=world, steps=steps, warmup=warmup,
                    ms_per_step=ms / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=op + " operands, fp32 master weights/accumulate/residual/softmax", data="synthetic", mode="train",
                    config=dict(workload=f"maest_30s_from_passt_pretrain training step: mel [{B},1,96,1875] fp16 per GPU, s_patchout_t=90 -> {N} tokens, "
                                         "mixup 0.3, BCE, fwd+bwd" + (" + NCCL gradient all-reduce" if world > 1 else "") + " + AdamW step",
                                batch_per_gpu=B, tokens=N, gflop_per_clip_fwd=fl["total"] / 1e9),
                    loss=float(loss.detach()), grad_elements_allreduced=n_grad[0], allreduce_bytes=(2 if ar_mode == "bf16" else 4) * n_grad[0],
                    allreduce_ms_exposed=ar_ms, allreduce=("one flat " + ("bf16-compressed" if ar_mode == "bf16" else "fp32") +
                               " dist.all_reduce (NCCL) after the last backward kernel, not overlapped") if world > 1 else "none (1 GPU)",
                    clocks=clocks,
                    model_tflops=tf, model_frac_of_bf16_sustained=tf / peaks["bf16_tflops_sustained"],
                    gpu_launches=steps * 333)      # mb:: kernels per step (profiles/r02d_launches_train_step.txt)
    del mod, net, opt
    torch.cuda.empty_cache()
    return line


def run_ingest(args):
    """--mode ingest: one step = one training batch of raw float16 .mmap windows
    (B clips x 1875 frames x 96 bands) -> [B, 1, 96, 1875] float16 model input.  `value`: kernel only, windows resident
    in HBM; `e2e`: MelWindowBatcher from files on the box's disk (window reads + pinned H2D + kernel)."""
    import tempfile
    import numpy as np
    import torch
    from maest_b200 import ingest, ops
    from oracle import ingest_oracle as IO

    torch.cuda.set_device(0)
    B, T = args.batch, 1875
    rng = np.random.RandomState(0)
    raw = torch.from_numpy((rng.rand(B, T, 96) * 5).astype(np.float16)).cuda()
    nread = torch.full((B,), T, dtype=torch.int32, device="cuda")
    shift = torch.from_numpy(rng.randint(-50, 51, B).astype(np.int32)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    for _ in range(max(args.warmup, 3)):
        ops.mel_ingest(raw, nread, shift, ingest.NORM_MEAN, ingest.NORM_STD)
    sampler = ClockSampler(0)
    sampler.start()
    ts = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        ops.mel_ingest(raw, nread, shift, ingest.NORM_MEAN, ingest.NORM_STD)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    with tempfile.TemporaryDirectory() as d:
        files = []
        for i in range(B):
            f = os.path.join(d, f"{i}.mmap")
            (rng.rand(4000, 96) * 5).astype(np.float16).tofile(f)
            files.append(f)
        bt = ingest.MelWindowBatcher(B, 30, roll=True)
        for _ in range(3):
            out = bt(files)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = bt(files)
            out[0, 0, 0, 0].item()          # D2H read of the step's result
        e2e_s = (time.perf_counter() - t0) / args.steps
        clocks = sampler.stop()
        # CPU baseline: the numpy restatement of the reference's loader on the same files, one thread (as a DataLoader worker would)
        t0 = time.perf_counter()
        n_cpu = min(B, 32)
        for f in files[:n_cpu]:
            whole = np.fromfile(f, dtype=np.float16).reshape(-1, 96)
            IO.ingest(whole, T, 100, ingest.NORM_MEAN, ingest.NORM_STD, 7)
        cpu_s = time.perf_counter() - t0
    peaks = measured_peaks()
    by = B * T * 96 * 2 * 2                      # algorithmic bytes: read the window once, write the batch once
    line = dict(metric="clips/sec (loader ingest)", value=B / (ms / 1e3), unit="clips/s", n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16", data="synthetic", mode="ingest",
                config=dict(workload=f"raw float16 [frames,96] windows -> [{B},1,96,{T}] float16: zero-pad centring, transpose, fp16 normalisation, time roll",
                            batch_per_gpu=B, l2="256 MB buffer zeroed between timed launches (L2 flush)"),
                e2e=dict(value=B / e2e_s, unit="clips/s", h2d_bytes_per_step=B * T * 96 * 2 + 8 * B, d2h_bytes_per_step=2,
                         note="MelWindowBatcher: np.memmap window reads into a pinned buffer + H2D + kernel, files on local disk (page cache warm)"),
                gpu_launches=args.steps, clocks=clocks,
                roofline=dict(bound="hbm", kernel="mel_ingest_kernel", achieved=by / (ms / 1e3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                              frac=by / (ms / 1e3) / 1e9 / peaks["hbm_gbs"], traffic=None, algorithmic_bytes_per_launch=by,
                              peak_source=peaks["source"] + " hbm_gbs"),
                cpu_baseline=dict(value=n_cpu / cpu_s, unit="clips/s", cores=1, kind="port",
                                  sample=f"{n_cpu} clips, oracle/ingest_oracle.py (numpy restatement of discogs/dataset.py + datamodule norm/roll), one thread"))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="discogs-maest-30s-pw-129e")
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--op-dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--attn-variant", type=int, default=8)
    ap.add_argument("--fuse-ln", action="store_true", help="fold the LayerNorms into the GEMM epilogues around them (A/B; default off)")
    ap.add_argument("--ref-clips", type=int, default=2, help="--impl reference: clips per step (bounded CPU sample)")
    ap.add_argument("--cpu-baseline-clips", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--train-steps", type=int, default=5, help="default (infer) mode: timed steps of the configs[3] training leg reported "
                                                               "under the `train` key (0 = skip)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train", "ingest"],
                    help="infer: BASELINE.json configs[2] (headline).  train: configs[3], one optimisation step per 'step'.  "
                         "ingest: loader -> device ingest kernel (SURVEY.md section 8(f) row 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "train":
        return run_train(args)
    if args.mode == "ingest":
        return run_ingest(args)

    import torch
    import torch.distributed as dist
    from maest_b200 import get_maest, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    S, grid_t = ARCH_T[args.arch]
    B = args.batch
    model = get_maest(arch=args.arch, pretrained=False, op_dtype=args.op_dtype, fuse_ln=args.fuse_ln)
    model.attn_variant = args.attn_variant
    model_attn_variant = args.attn_variant
    model.load_state_dict(synth.synth_state_dict(grid_t, 400, seed=0), strict=False)   # random-init weights (no checkpoints offline)
    model = model.to(dev).eval()

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    wav_dev = torch.rand(B, S, generator=g, device=dev) * 2 - 1          # synthetic full-scale noise, resident in HBM
    host = [torch.empty(B, S, dtype=torch.float32).pin_memory() for _ in range(2)]
    for h in host:
        h.copy_(wav_dev.cpu())
    T = 1 + S // 256
    P = 9 * ((T - 16) // 10 + 1)
    N = 2 + P

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        # ---------------- device-resident throughput ----------------
        for _ in range(args.warmup):
            model(wav_dev)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(args.steps):
            logits, _ = model(wav_dev)
        e1.record()
        barrier()
        ms_dev = max_over_ranks(e0.elapsed_time(e1))

        # ---------------- end-to-end through the public API, host buffers ----------------
        copy_stream = torch.cuda.Stream(device=dev)
        stage = [torch.empty(B, S, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        out_host = torch.empty(B, 400, dtype=torch.float32).pin_memory()

        def h2d(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])
                stage[i % 2].copy_(host[i % 2], non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_loop(steps):
            for c in consumed:
                c.record()
            h2d(0)
            for i in range(steps):
                if i + 1 < steps:
                    h2d(i + 1)
                torch.cuda.current_stream().wait_event(ready[i % 2])
                lo, _ = model(stage[i % 2])
                consumed[i % 2].record()
                out_host.copy_(lo, non_blocking=True)
            torch.cuda.synchronize()

        e2e_loop(2)
        barrier()
        e2, e3 = torch.cuda.Event(True), torch.cuda.Event(True)
        e2.record()
        e2e_loop(args.steps)
        e3.record()
        barrier()
        ms_e2e = max_over_ranks(e2.elapsed_time(e3))
        clocks = sampler.stop() if rank == 0 else None

        breakdown = None
        if rank == 0 and not args.no_breakdown:
            breakdown = kernel_breakdown(model, wav_dev)

    # ---------------- configs[3] training leg (every rank; gradient all-reduce over NCCL at N > 1) ----------------
    train_line = None
    if args.train_steps > 0:
        del model, stage, wav_dev
        torch.cuda.empty_cache()
        train_line = train_leg(args, args.train_steps, 3, B)
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    fl = flops_per_clip(N, P)
    clips = B * world * args.steps
    value = clips / (ms_dev / 1e3)
    e2e_val = clips / (ms_e2e / 1e3)
    # logmel, pos table, patch gather, patch GEMM, 12 x (LN | LN-finalise, qkv, attn, proj, LN | LN-finalise, fc1, fc2), pool/head: the
    # folded path replaces 23 LayerNorm kernels by 23 (tiny) statistics-finalise kernels, the launch count is the same
    per_step_launches = 5 + DEPTH * 7
    line = dict(metric="clips/sec", value=value, unit="clips/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=args.op_dtype + " operands, fp32 accumulate/residual/softmax/mel", data="synthetic",
                config=dict(workload=f"{args.arch} inference, waveform [{B},{S}] per GPU -> 96x{T} log-mel -> {N} tokens -> 12 blocks -> logits[{B},400]",
                            batch_per_gpu=B, tokens=N, weights="random-init (seeded, fp16/bf16-representable)",
                            l2="per-step working set (>1.4 GB activations) exceeds the 126 MB L2; no explicit flush",
                            gflop_per_clip=fl["total"] / 1e9),
                e2e=dict(value=e2e_val, unit="clips/s", h2d_bytes_per_step=B * S * 4, d2h_bytes_per_step=B * 400 * 4,
                         ms_per_step=ms_e2e / args.steps, note="pinned host waveform, double-buffered H2D on a copy stream, logits D2H every step"),
                gpu_launches=per_step_launches * args.steps, clocks=clocks,
                model_tflops=value / world * fl["total"] / 1e12, model_frac_of_bf16_sustained=value / world * fl["total"] / 1e12 / peaks["bf16_tflops_sustained"])
    if breakdown:
        gemm_ms = sum(v["ms_per_step"] for k, v in breakdown.items() if k.startswith("gemm_"))
        gemm_launches = sum(v["launches_per_step"] for k, v in breakdown.items() if k.startswith("gemm_"))
        gemm_flops = B * fl["linear"]
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
        total_ms = sum(v["ms_per_step"] for v in breakdown.values())
        traffic = None
        tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic.json")))      # newest round's ncu capture
        tpath = tfiles[-1] if tfiles else ""
        if os.path.exists(tpath) and B == 64 and args.arch == "discogs-maest-30s-pw-129e":
            with open(tpath) as tf:
                traffic = json.load(tf)["gemm_family_bytes_per_step"] / gemm_launches    # measured DRAM bytes per launch (ncu --set full)
        line["gemm_family"] = dict(bound="tensor", kernel="gemm_tn_kernel / gemm2_tn_kernel (qkv/proj/fc1/fc2, 48 launches per step)", achieved=achieved,
                                peak=peaks["bf16_tflops_sustained"], unit="TFLOP/s", frac=achieved / peaks["bf16_tflops_sustained"],
                                traffic=traffic, algorithmic_flops_per_launch=gemm_flops / gemm_launches,
                                algorithmic_bytes_per_launch=B * N * (2 * (768 + 2304) + 2 * 768 + 8 * 768 + 2 * (768 + 3072) + 2 * 3072 + 8 * 768) / 4, peak_source=peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                                share_of_step=gemm_ms / total_ms, launches_per_step=gemm_launches)
        att = breakdown.get("attention")
        if att:
            # top-level roofline = the DOMINANT kernel: the attention forward is the largest single kernel of the step (12 launches)
            # and the one furthest below its roofline; the GEMM family's numbers stay under `gemm_family`.
            a_tf = B * fl["attention"] / (att["ms_per_step"] / 1e3) / 1e12
            a_traffic = None
            afiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_attention_traffic.json")))
            if afiles and B == 64 and args.arch == "discogs-maest-30s-pw-129e":
                with open(afiles[-1]) as tf:
                    a_traffic = json.load(tf).get("dram_bytes_per_launch")
            kname = {0: "attention_fwd_spec_kernel", 3: "attention_fwd_chain_kernel<3 x 128>", 4: "attention_fwd_chain_kernel<4 x 96>",
                     5: "attention_fwd_chain_kernel<3 x 128, split columns>", 8: "attention_fwd_chain_kernel<3 x 128 + epilogue warpgroup>"}.get(model_attn_variant, f"attention variant {model_attn_variant}")
            line["roofline"] = dict(bound="tensor", kernel=f"{kname} (12 launches per step)", achieved=a_tf,
                                    peak=peaks["bf16_tflops_sustained"], unit="TFLOP/s", frac=a_tf / peaks["bf16_tflops_sustained"],
                                    traffic=a_traffic, algorithmic_flops_per_launch=B * fl["attention"] / DEPTH,
                                    algorithmic_bytes_per_launch=B * N * (3 * EMBED + EMBED) * 2,
                                    ms_per_launch=att["ms_per_launch"], share_of_step=att["ms_per_step"] / total_ms, launches_per_step=DEPTH,
                                    peak_source=peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)")
            line["attention"] = dict(achieved=a_tf, unit="TFLOP/s", frac=a_tf / peaks["bf16_tflops_sustained"], share_of_step=att["ms_per_step"] / total_ms)
        lm = breakdown.get("logmel")
        pt = breakdown.get("patch_tokens")
        if lm and pt:
            by = B * (4 * S + 2 * N * EMBED)
            gbs = by / ((lm["ms_per_step"] + pt["ms_per_step"]) / 1e3) / 1e9
            line["mel_patch_stage"] = dict(bound="hbm", achieved=gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=gbs / peaks["hbm_gbs"],
                                           algorithmic_bytes_per_clip=4 * S + 2 * N * EMBED)
        line["breakdown_ms_per_step"] = {k: round(v["ms_per_step"], 4) for k, v in breakdown.items()}
    if train_line:
        line["train"] = {k: train_line[k] for k in ("value", "unit", "ms_per_step", "steps", "dtype", "allreduce_bytes", "allreduce_ms_exposed",
                                                    "allreduce", "grad_elements_allreduced", "loss", "model_frac_of_bf16_sustained", "gpu_launches")}
        line["train"]["workload"] = train_line["config"]["workload"]
    if not args.no_cpu_baseline:
        v, cores, kind, what = cpu_baseline_clips_per_s(args.arch, args.cpu_baseline_clips)
        line["cpu_baseline"] = dict(value=v, unit="clips/s", cores=cores, kind=kind,
                                    sample=f"{args.cpu_baseline_clips} clips of the same workload, {what}, 1 run")
    ifiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_incumbent.json")))
    if ifiles:      # torch eager / library FMHA on the same B200, measured by tools/incumbent.py in a separate session (not timed here)
        with open(ifiles[-1]) as f:
            inc = json.load(f)
        line["incumbent_gpu"] = dict(source=os.path.relpath(ifiles[-1], ROOT), eager_impl=inc.get("eager_impl"),
                                     eager_clips_per_s={k: round(v["clips_per_s"], 1) for k, v in inc.get("eager_gpu", {}).items()},
                                     attention_ms={k: round(v["ms"], 4) for k, v in inc.get("attention_B64_H12_N1685_d64", {}).items() if "ms" in v})
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
