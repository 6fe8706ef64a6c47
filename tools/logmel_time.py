import sys, os, torch
sys.path.insert(0, os.getcwd())
from maest_b200 import ops, synth
x = (torch.rand(64, 480000, device="cuda") * 2 - 1)
ops.logmel(x); torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); m = ops.logmel(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ref = None
print("LOGMEL", os.path.basename(os.environ.get("MAEST_B200_LIB", "default")), round(min(ts), 4), round(sorted(ts)[5], 4), float(m.double().sum()))
