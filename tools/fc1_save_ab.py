"""A/B of the training-forward fc1 GEMM (GELU16_SAVE: two 16-bit outputs) on the 1-CTA and the CTA-pair kernel, M = 64 x 866."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maest_b200 import _lib, ops
M, K, N = 64 * 866, 768, 3072
for dt in (torch.bfloat16, torch.float16):
    A = (torch.randn(M, K, device="cuda") * 0.5).to(dt); W = (torch.randn(N, K, device="cuda") * 0.05).to(dt); b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=dt); pre = torch.empty_like(out)
    def t(fn, n=8):
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(n):
            a, c = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); fn(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
        return min(ts)
    ref = A[:256].double() @ W.double().t() + b.double()
    for mode in (0, 1, 0, 1):
        ops.set_gemm_mode(mode)
        ms = t(lambda: ops.gemm(A, W, _lib.EPI_GELU16, M, N, K, bias=b, out=out, aux16=pre))
        r1 = float((out[:256].double() - torch.nn.functional.gelu(ref)).norm() / torch.nn.functional.gelu(ref).norm())
        r2 = float((pre[:256].double() - ref).norm() / ref.norm())
        print("FC1SAVE", str(dt)[6:], "pair" if mode else "1cta", round(ms, 4), f"rel {r1:.1e} {r2:.1e}", flush=True)
ops.set_gemm_mode(None)
