"""A/B of the fc1 + GELU GEMM (M = 64 x 1685, K = 768, N = 3072): 1-CTA kernel vs CTA-pair kernel (maest_set_gemm_mode)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maest_b200 import _lib, ops
M, K, N = 64 * 1685, 768, 3072
for dt in (torch.float16, torch.bfloat16):
    A = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
    W = (torch.randn(N, K, device="cuda") * 0.05).to(dt)
    b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=dt)
    ref = torch.nn.functional.gelu(A[:256].double() @ W.double().t() + b.double())
    def t(fn, n=6):
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(n):
            a, c = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); fn(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
        return min(ts)
    for mode in (0, 1, 0, 1):
        ops.set_gemm_mode(mode)
        ms = t(lambda: ops.linear(A, W, b, _lib.EPI_GELU16, out=out))
        rel = float((out[:256].double() - ref).norm() / ref.norm())
        print("fc1 gelu", str(dt)[6:], "pair_mode", mode, round(ms, 4), "rel", f"{rel:.2e}")
ops.set_gemm_mode(None)
