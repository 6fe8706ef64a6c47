"""Golden values of the reference's learning-rate lambdas (helpers/ramp.py) -- dev container only.
    python tests/golden/make_golden_sched.py  ->  tests/golden/c7_sched.npz"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

ref_loader.install_stubs(with_lightning=False)
spec = importlib.util.spec_from_file_location("ref_ramp", os.path.join(ref_loader.REF_ROOT, "helpers", "ramp.py"))
ramp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ramp)
ep = np.arange(0, 140)
out = dict(epochs=ep,
           exp_lin=np.array([ramp.exp_warmup_linear_down(5, 50, 50, 0.01)(int(e)) for e in ep]),      # module defaults (models/module.py:27-40)
           exp_lin_b=np.array([ramp.exp_warmup_linear_down(20, 100, 50, 0.001)(int(e)) for e in ep]),
           cos_cyc=np.array([ramp.cosine_cycle(5, 50, 0.01)(int(e)) for e in ep]))
np.savez_compressed(os.path.join(HERE, "c7_sched.npz"), **out)
print({k: v[:8] for k, v in out.items()})
