"""Host (enqueue) time vs device time of one training step: is the step launch-bound?"""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maest_b200 import get_maest, synth
from maest_b200.module import Module
from maest_b200.optim import FusedAdamW
B = int(os.environ.get("B", 64))
net = get_maest(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=400, input_f=96, input_t=1875, s_patchout_t=90, op_dtype="bf16")
net.load_state_dict(synth.synth_state_dict(187, 400, seed=0), strict=False)
mod = Module(net=net, mixup_alpha=0.3, do_swa=False).cuda().train()
opt = FusedAdamW(mod.parameters(), lr=2e-5, weight_decay=1e-4)
g = torch.Generator(device="cuda").manual_seed(7)
x = (0.5 * torch.randn(B, 1, 96, 1875, generator=g, device="cuda")).half()
y = (torch.rand(B, 400, generator=g, device="cuda") > 0.99).half()
def step():
    opt.zero_grad(set_to_none=True)
    loss = mod.training_step((x, ["c"] * B, y), 0)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"TRAINHOST B={B} host enqueue {1e3 * (t1 - t0) / 10:.2f} ms/step, wall incl. device {1e3 * (t2 - t0) / 10:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
