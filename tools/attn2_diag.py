"""Round-2 attention forward diagnostics (GPU box): correctness of a variant against float64 math, the exact-redo path,
log-sum-exp, and A/B timing of variants / experiment libraries.  Every case runs in its own subprocess with a timeout, so a
protocol bug (hang, trap) in a new kernel costs one case, not the session.

    python tools/attn2_diag.py check 3        # correctness of variant 3 (uses libmaest_b200_safe.so when present)
    python tools/attn2_diag.py time 0 3       # timing, config 3 / config 4 / config 2 shapes
    ATT_LIBS="a.so b.so" python tools/attn2_diag.py time 3      # the same variant from several experiment builds
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref_attention(qkv, B, N):
    import torch
    q, k, v = qkv.view(B, N, 3, 12, 64).permute(2, 0, 3, 1, 4).double()
    s = (q @ k.transpose(-1, -2)) * 0.125
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, 768)
    lse = torch.logsumexp(s, -1) * 1.4426950408889634        # [B,12,N] in log2 units
    return o, lse


def child():
    import torch
    from maest_b200 import ops
    mode, variant = os.environ["ATT_MODE"], int(os.environ["ATT_VARIANT"])
    B, N = int(os.environ.get("ATT_B", 64)), int(os.environ.get("ATT_N", 1685))
    dt = torch.bfloat16 if os.environ.get("ATT_DT") == "bf16" else torch.float16
    out = dict(lib=os.path.basename(os.environ.get("MAEST_B200_LIB", "default")), mode=mode, variant=variant, B=B, N=N, dt=str(dt)[6:])
    g = torch.Generator().manual_seed(B * 1000 + N)
    if mode == "check":
        qkv = torch.randn(B * N, 2304, generator=g).to(dt).cuda()
        ref, lse_ref = ref_attention(qkv, B, N)
        o, lse = ops.attention(qkv, B, N, 12, variant, save_lse=True)
        torch.cuda.synchronize()
        out["rel"] = float((o.double() - ref).norm() / ref.norm())
        out["lse_maxabs"] = float((lse.double() - lse_ref).abs().max())
        out["nan"] = bool(torch.isnan(o.float()).any())
        o2 = ops.attention(qkv, B, N, 12, variant)
        out["deterministic"] = bool(torch.equal(o, o2))
    elif mode == "sharp":
        qkv = (torch.randn(B * N, 2304, generator=g) * 4).to(dt).cuda()
        qkv.view(B * N, 3, 12, 64)[N // 2:, 1] *= 3
        ref, lse_ref = ref_attention(qkv, B, N)
        o, lse = ops.attention(qkv, B, N, 12, variant, save_lse=True)
        torch.cuda.synchronize()
        out["rel"] = float((o.double() - ref).norm() / ref.norm())
        out["lse_maxabs"] = float((lse.double() - lse_ref).abs().max())
        out["nan"] = bool(torch.isnan(o.float()).any())
    elif mode == "time":
        qkv = torch.randn(B * N, 2304, generator=g).to(dt).cuda()
        fn = lambda: ops.attention(qkv, B, N, 12, variant)
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(8):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        # back-to-back (sustained clocks): 30 launches
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(30):
            fn()
        b.record(); torch.cuda.synchronize()
        out.update(ms_best=round(min(ts), 4), ms_sustained=round(a.elapsed_time(b) / 30, 4),
                   tflops_best=round(4 * N * N * 64 * 12 * B / min(ts) / 1e9, 1))
    elif mode == "clocks":      # ATC_DIAG build: per-role wait clocks of the chain kernel (written over the lse buffer)
        qkv = torch.randn(B * N, 2304, generator=g).to(dt).cuda()
        ops.attention(qkv, B, N, 12, variant, save_lse=True)
        _, lse = ops.attention(qkv, B, N, 12, variant, save_lse=True)
        torch.cuda.synchronize()
        nch, bkv = (4, 96) if variant == 4 else (3, 64) if variant == 6 else (3, 128)
        nsm = 4 * nch * (2 if variant == 5 else 1)
        d = lse.reshape(-1)[: 148 * 512].view(148, 512).double()
        nq, nkv = (N + 127) // 128, (N + bkv - 1) // bkv
        tiles = B * 12 * nq * nkv / 148.0
        items = B * 12 * nq / 148.0
        sm = d[:, :16 * nsm].view(148, nsm, 16).mean((0, 1))
        names = ["wait_s", "wait_mref", "flush", "epilogue", "exp_to_arrive", "leader_max", "total"]
        out["tiles_per_sm"] = round(tiles)
        out["softmax_per_tile"] = {k: round(float(sm[i]) / (tiles / nch)) for i, k in enumerate(names)}
        out["bad_rows"] = float(sm[10])
        t = d[:, 400:404].mean(0) / tiles
        out["tma_per_tile"] = dict(zip(["wait_kv_empty", "-", "wait_q_empty", "total"], [round(float(x)) for x in t]))
        q = d[:, 420:425].mean(0) / tiles
        out["qk_issuer_per_tile"] = dict(zip(["wait_q", "wait_kv", "wait_pv_done", "issue", "total"], [round(float(x)) for x in q]))
        m = d[:, 408:417].mean(0) / tiles
        out["pv_issuer_per_tile"] = dict(zip(["-", "wait_kv", "wait_o_empty", "wait_p", "-", "total", "pv_issue", "commit(lean)", "next|mma(lean)"], [round(float(x)) for x in m]))
    print("ATTN2 " + json.dumps(out), flush=True)


def run(env_extra, timeout=240):
    env = dict(os.environ, ATT_CHILD="1", **{k: str(v) for k, v in env_extra.items()})
    try:
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=timeout)
        lines = [l for l in r.stdout.splitlines() if l.startswith("ATTN2 ")]
        print(lines[-1] if lines else f"ATTN2-FAIL {env_extra}: rc={r.returncode} {r.stdout[-300:]} {r.stderr[-600:]}", flush=True)
        return bool(lines)
    except subprocess.TimeoutExpired:
        print(f"ATTN2-TIMEOUT {env_extra}", flush=True)
        return False


if __name__ == "__main__":
    if os.environ.get("ATT_CHILD"):
        child()
        sys.exit(0)
    mode = sys.argv[1]
    variants = [int(v) for v in sys.argv[2:]] or [3]
    libs = os.environ.get("ATT_LIBS", "").split() or [None]
    safe = os.path.join(ROOT, "maest_b200", "lib", "libmaest_b200_safe.so")
    for lib in libs:
        for v in variants:
            e = {"ATT_VARIANT": v}
            if lib:
                e["MAEST_B200_LIB"] = lib if os.path.isabs(lib) else os.path.join(ROOT, "maest_b200", "lib", lib)
            if mode == "check":
                if not lib and os.path.exists(safe) and os.environ.get("ATT_SAFE", "1") == "1":
                    e["MAEST_B200_LIB"] = safe
                ok = True
                for (B, N, dt) in [(1, 300, "f16"), (2, 560, "f16"), (1, 1685, "f16"), (3, 866, "bf16"), (1, 129, "f16"), (2, 256, "bf16"),
                                   (1, 3, "f16"), (1, 128, "f16"), (20, 1685, "f16"), (64, 866, "bf16")]:
                    ok = run(dict(e, ATT_MODE="check", ATT_B=B, ATT_N=N, ATT_DT=dt), 120) and ok
                    if not ok:
                        break
                if ok:
                    run(dict(e, ATT_MODE="sharp", ATT_B=1, ATT_N=700, ATT_DT="f16"), 120)
                    run(dict(e, ATT_MODE="sharp", ATT_B=2, ATT_N=1685, ATT_DT="bf16"), 120)
            elif mode == "clocks":
                run(dict(e, ATT_MODE="clocks", ATT_B=64, ATT_N=1685, ATT_DT="f16"), 180)
            else:
                shapes = [(64, 1685, "f16"), (64, 1685, "bf16"), (64, 866, "bf16"), (64, 560, "f16")]
                for (B, N, dt) in shapes[:1] if os.environ.get("ATT_QUICK") else shapes:
                    run(dict(e, ATT_MODE="time", ATT_B=B, ATT_N=N, ATT_DT=dt), 180)
