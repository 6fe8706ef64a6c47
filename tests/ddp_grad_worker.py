"""torchrun worker of tests/test_gpu_train.py::test_grad_allreduce_two_ranks_match_single_rank_mean (2 GPUs, NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maest_b200 import get_maest, synth  # noqa: E402
from maest_b200.module import my_mixup  # noqa: E402
from maest_b200.train import training_forward  # noqa: E402


def grads(rank_seed, allreduce):
    m = get_maest(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=400, input_f=96, input_t=1875, s_patchout_t=90, op_dtype="bf16")
    m.load_state_dict(synth.synth_state_dict(187, 400, seed=0), strict=False)
    m = m.cuda().train()
    if allreduce:
        m.grad_allreduce = True
    x, y = synth.train_batch(2, seed_x=100 + rank_seed, seed_y=200 + rank_seed)
    torch.manual_seed(1 + rank_seed)
    np.random.seed(1 + rank_seed)
    loss, _ = training_forward(m, x.cuda(), y.cuda(), my_mixup(2, 0.3))
    loss.backward()
    return {n: p.grad.double().clone() for n, p in m.named_parameters() if p.grad is not None}


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    reduced = grads(rank, True)                       # this rank's batch, gradients averaged over the 2 ranks by train.py
    g0, g1 = grads(0, False), grads(1, False)         # both single-rank gradients, recomputed locally without communication
    # The backward is not bit-reproducible run to run (fp32 reduce-adds in dQ and the split-K weight gradients: ~1e-7) and bf16
    # roundings downstream amplify that by ~2.5x per block (see test_grad_allreduce_world1_equals_plain_path): tight where no
    # amplification has happened yet, the parity tolerance elsewhere.
    err, worst = 0.0, ""
    good = True
    for n, g in reduced.items():
        want = (g0[n] + g1[n]) / 2
        e = float((g - want).norm() / want.norm().clamp_min(1e-30))
        tol = 1e-5 if n.startswith(("head.", "norm.", "blocks.11.")) else 1e-2
        if e > err:
            err, worst = e, n
        good = good and e < tol
    ok = torch.tensor([1.0 if good else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"largest rel err {err:.3e} ({worst})")
        print("DDP_GRAD_OK" if float(ok) == 1.0 else "DDP_GRAD_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if float(ok) == 1.0 else 1)


if __name__ == "__main__":
    main()
