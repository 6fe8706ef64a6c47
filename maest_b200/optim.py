"""Optimiser side of the training step: one-launch AdamW over all parameters with the SWA running
average fused in, and the reference's learning-rate lambdas.

Mirrors `Module.get_optimizer` / `get_scheduler_lambda` / `get_lr_scheduler` (models/module.py:207-243) and the running average
that `StochasticWeightAveragingAndCopy` (helpers/swa_callback.py) transfers into `net_swa` (torch.optim.swa_utils:
avg += (p - avg) / (n_averaged + 1)).

The four closed-form ramps below (`exp_rampup`, `linear_rampdown`, `exp_warmup_linear_down`, `cosine_cycle`) are TRANSCRIBED
from helpers/ramp.py:21-33,47-62,102-109,124-137 (same names, same `np.clip(epoch, 0.5, ...)` bodies): ~35 lines of host-side
schedule arithmetic that must be value-identical to the reference (pinned by tests/golden/c7_sched.npz); they are not a
re-design target."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib

CHUNK = 8192


# ------------------------------------------------------------------------------------------------ LR lambdas (helpers/ramp.py)
def exp_rampup(rampup_length):
    def wrapper(epoch):
        if epoch < rampup_length:
            epoch = np.clip(epoch, 0.5, rampup_length)
            phase = 1.0 - epoch / rampup_length
            return float(np.exp(-5.0 * phase * phase))
        return 1.0
    return wrapper


def linear_rampdown(rampdown_length, start=0, last_value=0):
    def wrapper(epoch):
        if epoch <= start:
            return 1.0
        if epoch - start < rampdown_length:
            return last_value + (1.0 - last_value) * (rampdown_length - epoch + start) / rampdown_length
        return last_value
    return wrapper


def exp_warmup_linear_down(warmup, rampdown_length, start_rampdown, last_value):
    up, down = exp_rampup(warmup), linear_rampdown(rampdown_length, start_rampdown, last_value)
    return lambda epoch: up(epoch) * down(epoch)


def cosine_cycle(cycle_len=20, ramp_down_start=100, last_lr_value=0.01):
    ramp_down_start = cycle_len + (ramp_down_start - 1) // cycle_len * cycle_len

    def wrapper(epoch):
        ep = (epoch + cycle_len // 2.0) / (1.0 * cycle_len)
        if epoch > ramp_down_start:
            return last_lr_value
        return float(last_lr_value + (1.0 - last_lr_value) * 0.5 * (np.cos(2.0 * np.pi * ep) + 1))
    return wrapper


def get_scheduler_lambda(warm_up_len, ramp_down_start, ramp_down_len, last_lr_value, schedule_mode):
    """models/module.py:207-221."""
    if schedule_mode == "exp_lin":
        return exp_warmup_linear_down(warm_up_len, ramp_down_len, ramp_down_start, last_lr_value)
    if schedule_mode == "cos_cyc":
        return cosine_cycle(warm_up_len, ramp_down_start, last_lr_value)
    raise RuntimeError(f"schedule_mode={schedule_mode} Unknown for a lambda funtion.")


# ------------------------------------------------------------------------------------------------ fused AdamW
class _OptTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("swa", C.c_void_p), ("n", C.c_int64)]


class _OptChunk(C.Structure):
    _fields_ = [("tensor", C.c_int32), ("pad", C.c_int32), ("start", C.c_int64)]


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (same defaults, same per-step arithmetic in fp32, no amsgrad / maximize), one kernel launch
    per step for ALL parameters that have a gradient.  `swa_params`: optional iterable of tensors parallel to the parameters
    (e.g. `net_swa.parameters()`); `step(update_swa=True)` also folds the new weights into them."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, swa_params: Optional[Iterable] = None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._swa = None if swa_params is None else list(swa_params)
        self.n_averaged = 0
        self._tables = None

    def _build(self, group, plist):
        dev = plist[0].device
        all_params = [p for g in self.param_groups for p in g["params"]]
        rows, chunks, keep = [], [], []
        for ti, p in enumerate(plist):
            st = self.state[p]
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            swa = None
            if self._swa is not None:
                swa = self._swa[next(i for i, q in enumerate(all_params) if q is p)]
            assert p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous() and p.grad.dtype == torch.float32
            rows.append(_OptTensor(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                   None if swa is None else swa.data_ptr(), p.numel()))
            chunks += [_OptChunk(ti, 0, s) for s in range(0, p.numel(), CHUNK)]
        tb = bytes((_OptTensor * len(rows))(*rows))
        cb = bytes((_OptChunk * len(chunks))(*chunks))
        tt = torch.frombuffer(bytearray(tb), dtype=torch.uint8).to(dev)
        ct = torch.frombuffer(bytearray(cb), dtype=torch.uint8).to(dev)
        sig = tuple((r.p, r.g, r.m, r.v, r.swa, r.n) for r in rows)
        return sig, tt, ct, len(chunks)

    @torch.no_grad()
    def step(self, closure=None, update_swa: bool = False, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            if not plist[0].is_cuda:
                raise RuntimeError("FusedAdamW runs on CUDA (sm_100a) only; there is no CPU fallback")
            sig = tuple((p.data_ptr(), p.grad.data_ptr()) for p in plist)
            cache = (self._tables or {}).get(gi)
            if cache is None or cache[0] != sig:
                _, tt, ct, n = self._build(group, plist)
                self._tables = dict(self._tables or {})
                self._tables[gi] = cache = (sig, tt, ct, n)
            for p in plist:
                self.state[p]["step"] += 1
            step = self.state[plist[0]]["step"]
            swa_inv = 0.0
            if update_swa and self._swa is not None:
                swa_inv = 1.0 / (self.n_averaged + 1)
            with torch.cuda.device(plist[0].device):
                lib = _lib.init(plist[0].device.index if plist[0].device.index is not None else torch.cuda.current_device())
                _lib.check(lib.maest_adamw_step(cache[1].data_ptr(), cache[2].data_ptr(), cache[3], float(group["lr"]),
                                                float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                                                float(group["weight_decay"]), int(step), float(grad_scale), float(swa_inv),
                                                torch.cuda.current_stream().cuda_stream), "adamw_step")
            # The kernel wrote the parameters (and the SWA copies) through raw pointers: tell autograd / every cache keyed on
            # `tensor._version` (MAEST._weight16's 16-bit operand copies, MAEST._ln_fold) that the values changed.
            torch._C._increment_version(plist)
            if swa_inv != 0.0:
                all_params = [p for g in self.param_groups for p in g["params"]]
                pid = {id(p) for p in plist}
                torch._C._increment_version([s for p, s in zip(all_params, self._swa) if id(p) in pid])
        if update_swa and self._swa is not None:
            self.n_averaged += 1
        return loss

    @torch.no_grad()
    def fold_into_swa(self, params, swa_params, n_averaged: int):
        """swa += (p - swa) / (n_averaged + 1) for every (p, swa) pair in ONE launch (maest_swa_fold): the epoch-end update of
        the SWA callback (maest_b200.module.StochasticWeightAveragingAndCopy), which makes the averaged model reachable from
        Lightning without AveragedModel's second copy."""
        params, swa_params = list(params), list(swa_params)
        if not params:
            return
        dev = params[0].device
        sig = tuple((p.data_ptr(), s.data_ptr(), p.numel()) for p, s in zip(params, swa_params))
        cache = getattr(self, "_swa_tables", None)
        if cache is None or cache[0] != sig:
            rows, chunks = [], []
            for ti, (p, s) in enumerate(zip(params, swa_params)):
                assert p.dtype == s.dtype == torch.float32 and p.is_contiguous() and s.is_contiguous() and p.numel() == s.numel()
                rows.append(_OptTensor(p.data_ptr(), None, None, None, s.data_ptr(), p.numel()))
                chunks += [_OptChunk(ti, 0, st) for st in range(0, p.numel(), CHUNK)]
            tt = torch.frombuffer(bytearray(bytes((_OptTensor * len(rows))(*rows))), dtype=torch.uint8).to(dev)
            ct = torch.frombuffer(bytearray(bytes((_OptChunk * len(chunks))(*chunks))), dtype=torch.uint8).to(dev)
            self._swa_tables = cache = (sig, tt, ct, len(chunks))
        with torch.cuda.device(dev):
            lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
            _lib.check(lib.maest_swa_fold(cache[1].data_ptr(), cache[2].data_ptr(), cache[3], 1.0 / (n_averaged + 1),
                                          torch.cuda.current_stream().cuda_stream), "swa_fold")
        torch._C._increment_version(swa_params)
