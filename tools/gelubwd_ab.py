"""Timing of the fc2 input-gradient GEMM (GELUBWD16: dupre = (dx @ W_fc2) * gelu'(upre), column sums into the fc1 bias gradient)
at the training shape M = 64 x 866, N = 3072, K = 768.  MAEST_B200_LIB selects an experiment build."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maest_b200 import _lib, ops
M, N, K = 64 * 866, 3072, 768
for dt in (torch.bfloat16, torch.float16):
    dx = (torch.randn(M, K, device="cuda") * 0.5).to(dt); W = (torch.randn(K, N, device="cuda") * 0.05).to(dt)
    upre = torch.randn(M, N, device="cuda").to(dt); out = torch.empty_like(upre); cs = torch.zeros(N, device="cuda")
    fn = lambda: ops.gemm(dx, W, _lib.EPI_GELUBWD16, M, N, K, b_mn=True, out=out, aux16=upre, colsum_out=cs)
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(10):
        a, c = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); fn(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
    u = upre[:256].double().requires_grad_(True)
    torch.nn.functional.gelu(u).sum().backward()
    ref = (dx[:256].double() @ W.double()) * u.grad
    print("GELUBWD", os.path.basename(os.environ.get("MAEST_B200_LIB", "default")), str(dt)[6:], round(min(ts), 4), round(sorted(ts)[5], 4),
          f"rel {float((out[:256].double() - ref).norm() / ref.norm()):.1e}", flush=True)
