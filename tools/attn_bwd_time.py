"""Isolated timing of the attention backward (maest_attention_bwd: delta + tile kernel + dq cast) at the training shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maest_b200 import ops
for (B, N, dt) in [(64, 866, torch.bfloat16), (64, 1685, torch.bfloat16), (64, 866, torch.float16)]:
    g = torch.Generator(device="cuda").manual_seed(B + N)
    qkv = torch.randn(B * N, 2304, generator=g, device="cuda").to(dt)
    o, lse = ops.attention(qkv, B, N, 12, 8, save_lse=True)
    d_o = torch.randn(B * N, 768, generator=g, device="cuda").to(dt)
    out = torch.empty_like(qkv)
    ops.attention_bwd(qkv, o, d_o, lse, B, N, 12, out=out); torch.cuda.synchronize(); ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); ops.attention_bwd(qkv, o, d_o, lse, B, N, 12, out=out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    fl = 2.5 * 4 * N * N * 64 * 12 * B
    print("ATTNBWD", B, N, str(dt)[6:], "ms", round(min(ts), 4), "TFLOP/s", round(fl / min(ts) / 1e9, 1), flush=True)
