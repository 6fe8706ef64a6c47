"""Mainloop-only timing of the four encoder GEMMs (experiment build with -DGEMM_DIAG_NOEPI skips the epilogue): how much of each
kernel is epilogue-limited.   MAEST_B200_LIB=maest_b200/lib/libmaest_b200_xnoepi.so python tools/gemm_diag.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maest_b200 import _lib, ops
M = 64 * 1685
def t(fn, n=6):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, c = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); fn(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
    return min(ts)
for (N, K, epi, nm) in [(2304, 768, _lib.EPI_STORE16, "qkv"), (768, 768, _lib.EPI_RESID32, "proj"), (3072, 768, _lib.EPI_GELU16, "fc1"), (768, 3072, _lib.EPI_RESID32, "fc2")]:
    A = (torch.randn(M, K, device="cuda") * 0.5).half(); W = (torch.randn(N, K, device="cuda") * 0.05).half(); b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float16 if epi in (_lib.EPI_STORE16, _lib.EPI_GELU16) else torch.float32)
    for mode in (0, 1):
        ops.set_gemm_mode(mode)
        ms = t(lambda: ops.linear(A, W, b, epi, out=out, resid=out if epi == _lib.EPI_RESID32 else None))
        print("GEMMDIAG", os.path.basename(os.environ.get("MAEST_B200_LIB", "default")), nm, "pair" if mode else "1cta", round(ms, 4), round(2.0 * M * N * K / ms / 1e9, 1), "TFLOP/s")
