"""Incumbent numbers on the same B200 (SURVEY.md section 2a / BASELINE.md section 3: "the existing sm_100 path to beat is torch eager on
the same B200, fp32 and autocast bf16"), recorded beside the CPU baseline -- NOT part of the product path or of bench.py's timed regions.

    python tools/incumbent.py [--batch 64] [--out gpurun_out/r02_incumbent.json]

Measures, at configs[2] (discogs-maest-30s-pw-129e, waveform [B, 480000] -> logits):
  * the UNMODIFIED reference (baseline/_ref, `maest.get_maest(arch)`) moved to cuda: fp32 eager, TF32 matmuls, autocast(bf16);
    when baseline/_ref is absent the torch restatement oracle/maest_oracle.py takes its place (stated in the output);
  * torch.nn.functional.scaled_dot_product_attention (cuDNN / flash sm_100 FMHA) and flash_attn at the attention shape of the model
    (B x 12 heads x 1685 tokens x d 64, fp16 and bf16), the bar for csrc/attention_chain.cuh;
  * our attention forward and whole model on the same box for a like-for-like ratio.
Every number: CUDA events, 3 warm-up + 10 timed calls, inputs resident in HBM.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, warm=3, iters=10):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_incumbent.json"))
    args = ap.parse_args()
    import torch
    import torch.nn.functional as F
    from maest_b200 import get_maest, ops, synth

    B, S, N, H, D = args.batch, 480000, 1685, 12, 64
    dev = torch.device("cuda", 0)
    out = dict(batch=B, workload="discogs-maest-30s-pw-129e inference, waveform [B,480000] -> logits (1685 tokens)", gpu=torch.cuda.get_device_name(0))
    sd = synth.synth_state_dict(187, 400, seed=0)
    g = torch.Generator(device=dev).manual_seed(1234)
    wav = torch.rand(B, S, generator=g, device=dev) * 2 - 1

    # ---- whole model: the reference's own eager path on this GPU
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    fwd, what = None, None
    if os.path.isfile(os.path.join(ref_dir, "maest", "maest.py")):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        sys.path.insert(0, ref_dir)
        import ref_loader
        ref_loader.install_stubs(with_lightning=True)
        import maest as ref_pkg
        net = ref_pkg.get_maest(arch="discogs-maest-30s-pw-129e", pretrained=False)
        net.load_state_dict(sd, strict=False)
        net = net.to(dev).eval()
        fwd, what = (lambda x: net(x.clone())[0]), "unmodified reference (baseline/_ref) on cuda"
    else:
        from oracle import maest_oracle as O
        sd_dev = {k: v.to(dev) for k, v in sd.items()}
        fwd, what = (lambda x: O.forward(x, sd_dev, img_t=S // 256, dtype=torch.float32)[0]), "oracle/maest_oracle.py (torch restatement) on cuda"
    out["eager_impl"] = what
    eager = {}
    with torch.no_grad():
        for name, setup in (("fp32", dict(tf32=False, autocast=None)), ("tf32", dict(tf32=True, autocast=None)),
                            ("autocast_bf16", dict(tf32=True, autocast=torch.bfloat16)), ("autocast_fp16", dict(tf32=True, autocast=torch.float16))):
            torch.backends.cuda.matmul.allow_tf32 = setup["tf32"]
            torch.backends.cudnn.allow_tf32 = setup["tf32"]
            bsz = B
            while bsz >= 1:
                try:
                    x = wav[:bsz]
                    if setup["autocast"] is None:
                        ms = timed(lambda: fwd(x), warm=2, iters=3)
                    else:
                        def run():
                            with torch.autocast("cuda", dtype=setup["autocast"]):
                                return fwd(x)
                        ms = timed(run, warm=2, iters=3)
                    eager[name] = dict(batch=bsz, ms_per_step=ms, clips_per_s=bsz / ms * 1e3)
                    break
                except torch.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    bsz //= 2
    out["eager_gpu"] = eager
    del fwd
    torch.cuda.empty_cache()

    # ---- attention at the model's shape: library FMHA kernels vs ours
    att = {}
    fl = 4.0 * N * N * D * H * B
    for dtname, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        qkv = torch.randn(B * N, 3 * H * D, generator=g, device=dev).to(dt)
        q, k, v = qkv.view(B, N, 3, H, D).permute(2, 0, 3, 1, 4)          # [B, H, N, D] strided views of the packed activation
        qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
        with torch.no_grad():
            for backend in ("CUDNN_ATTENTION", "FLASH_ATTENTION", "EFFICIENT_ATTENTION"):
                try:
                    from torch.nn.attention import SDPBackend, sdpa_kernel
                    with sdpa_kernel([getattr(SDPBackend, backend)]):
                        ms = timed(lambda: F.scaled_dot_product_attention(qc, kc, vc))
                    att[f"sdpa_{backend.lower()}_{dtname}"] = dict(ms=ms, tflops=fl / ms / 1e9)
                except Exception as e:  # noqa: BLE001
                    att[f"sdpa_{backend.lower()}_{dtname}"] = dict(error=f"{type(e).__name__}: {str(e)[:120]}")
            try:
                from flash_attn import flash_attn_func
                qf, kf, vf = (t.transpose(1, 2).contiguous() for t in (qc, kc, vc))      # [B, N, H, D]
                ms = timed(lambda: flash_attn_func(qf, kf, vf))
                att[f"flash_attn_{dtname}"] = dict(ms=ms, tflops=fl / ms / 1e9)
            except Exception as e:  # noqa: BLE001
                att[f"flash_attn_{dtname}"] = dict(error=f"{type(e).__name__}: {str(e)[:120]}")
            for variant in (0, 3, 5):
                try:
                    ms = timed(lambda: ops.attention(qkv, B, N, H, variant))
                    att[f"ours_variant{variant}_{dtname}"] = dict(ms=ms, tflops=fl / ms / 1e9)
                except Exception as e:  # noqa: BLE001
                    att[f"ours_variant{variant}_{dtname}"] = dict(error=f"{type(e).__name__}: {str(e)[:120]}")
    out["attention_B64_H12_N1685_d64"] = att

    # ---- our model on the same box
    model = get_maest(arch="discogs-maest-30s-pw-129e", pretrained=False)
    model.load_state_dict(sd, strict=False)
    model = model.to(dev).eval()
    with torch.no_grad():
        ms = timed(lambda: model(wav), warm=3, iters=10)
    out["ours"] = dict(batch=B, ms_per_step=ms, clips_per_s=B / ms * 1e3)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
