"""A/B of the in-place residual GEMMs of the inference encoder (proj: K = 768, fc2: K = 3072; M = 64 x 1685, N = 768):
MAEST_RESID_REDUCE=0 -> load-add-store epilogue, default -> TMA reduce-add into the residual stream."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maest_b200 import _lib, ops
M, N = 64 * 1685, 768
for K in (768, 3072):
    for dt in (torch.bfloat16, torch.float16):
        A = (torch.randn(M, K, device="cuda") * 0.5).to(dt); W = (torch.randn(N, K, device="cuda") * 0.05).to(dt); b = torch.randn(N, device="cuda")
        x = torch.randn(M, N, device="cuda")
        fn = lambda: ops.linear(A, W, b, _lib.EPI_RESID32, resid=x, out=x)
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(10):
            a, c = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); fn(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
        print("RESID", "reduce" if os.environ.get("MAEST_RESID_REDUCE", "1") != "0" else "ldst", K, str(dt)[6:], round(min(ts), 4), round(sorted(ts)[5], 4), flush=True)
