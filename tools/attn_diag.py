"""Attention forward timing diagnostics (GPU box): A/B of experiment builds, occupancy, per-phase clocks.

    python tools/attn_diag.py            # prints one JSON line per (library, variant)
Experiment libraries are built on the dev box with maest_b200.build.build_variant(name, defines)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from maest_b200 import ops
    B, N = int(os.environ.get("ATT_B", 64)), int(os.environ.get("ATT_N", 1685))
    variant = int(os.environ.get("ATT_VARIANT", 0))
    torch.manual_seed(0)
    qkv = torch.randn(B * N, 2304, device="cuda").half()

    def timeit(fn, n=6):
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        return min(ts)

    q, k, v = qkv[:N].view(1, N, 3, 12, 64).permute(2, 0, 3, 1, 4).double()
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).transpose(1, 2).reshape(N, 768)
    o = ops.attention(qkv[:N].contiguous(), 1, N, 12, variant)
    rel = float((o.double() - ref).norm() / ref.norm())
    ms = timeit(lambda: ops.attention(qkv, B, N, 12, variant))
    out = dict(lib=os.path.basename(os.environ.get("MAEST_B200_LIB", "default")), variant=variant, B=B, N=N, ms=round(ms, 4), rel=rel,
               tflops=round(4 * N * N * 64 * 12 * B / ms / 1e9, 1))
    if "clocks" in os.environ.get("MAEST_B200_LIB", ""):
        _, lse = ops.attention(qkv, B, N, 12, variant, save_lse=True)
        torch.cuda.synchronize()
        nq = (N + 127) // 128
        idx = torch.arange(nq - 1, device="cuda") * 128
        d = torch.stack([lse[:, :, idx + i] for i in range(8)], -1).double()    # [B,H,nq-1,8]
        nkv = nq
        m = d.mean((0, 1, 2))
        out["clk_per_tile"] = dict(wait_s=round(float(m[0]) / nkv), ld=round(float(m[1]) / nkv), max=round(float(m[2]) / nkv),
                                   rescale=round(float(m[3]) / nkv), exp=round(float(m[4]) / nkv), wait_o=round(float(m[5]) / nkv),
                                   st_arrive=round(float(m[6]) / nkv), total_cta=round(float(m[7])))
    print("ATTN " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if os.environ.get("ATT_CHILD"):
        child()
        sys.exit(0)
    libs = [None] + (sorted(glob.glob(os.path.join(ROOT, "maest_b200", "lib", "libmaest_b200_x*.so"))) if os.environ.get("ATT_XLIBS") else [])
    for lib in libs:
        for variant in [0, 2]:
            env = dict(os.environ, ATT_CHILD="1", ATT_VARIANT=str(variant))
            if lib:
                env["MAEST_B200_LIB"] = lib
            r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=300)
            lines = [l for l in r.stdout.splitlines() if l.startswith("ATTN ")]
            print(lines[-1] if lines else f"ATTN-FAIL {lib} {variant}: {r.stderr[-400:]}", flush=True)
