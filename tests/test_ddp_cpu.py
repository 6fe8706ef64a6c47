"""CPU, world_size 2, gloo: host-side logic of the data-parallel path (one process per GPU in production; NCCL there).
Covers (a) gradient averaging with parameters that receive no gradient (`head_dist.*`, the reason the reference needs
find_unused_parameters=True, ex_maest.py:57), (b) clip sharding without a data-path collective, (c) max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from maest_b200.module import allreduce_gradients
    torch.manual_seed(0)
    net = torch.nn.ModuleDict(dict(a=torch.nn.Linear(8, 4), head_dist=torch.nn.Linear(4, 3)))   # head_dist never used
    x = torch.full((2, 8), float(rank + 1))
    net["a"](x).sum().backward()
    n = allreduce_gradients(net)
    g = net["a"].weight.grad.clone()
    # clip sharding: clip i -> rank i % world, no communication; union covers every clip exactly once
    clips = list(range(10))
    mine = clips[rank::world]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # max-over-ranks timing as in bench.py
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save(dict(n=n, grad=g, unused=net["head_dist"].weight.grad is None, shards=gathered, tmax=float(t)), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_sharding(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["n"] == 8 * 4 + 4 and r["unused"]
    assert torch.allclose(r["grad"], torch.full((4, 8), 2 * 1.5))      # mean of rank grads: 2 rows * (1 + 2) / 2
    assert sorted(sum(r["shards"], [])) == list(range(10)) and r["tmax"] == 11.0
