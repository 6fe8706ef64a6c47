"""Lightning-facing training module: the `Module` surface of the reference (models/module.py:44-276) on the B200 path.

`lightning` is optional: when it is not importable (this image) a minimal `LightningModule` stand-in with no-op
`log` / `log_dict` keeps the class usable as a plain nn.Module (training_step, predict_step, configure_optimizers).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .maest import _MiniIngredient, get_maest

try:  # pragma: no cover
    import lightning.pytorch as pl
    _LightningModule = pl.LightningModule
except Exception:  # noqa: BLE001
    class _LightningModule(torch.nn.Module):
        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

        def all_gather(self, x, *a, **k):
            return x

try:  # pragma: no cover
    from sacred import Ingredient as _SacredIngredient
    module_ing = _SacredIngredient("module")
except Exception:  # noqa: BLE001
    module_ing = _MiniIngredient("module")

MODULE_DEFAULT_CONF = dict(      # models/module.py:22-41
    do_swa=True, swa_epoch_start=50, swa_lrs=2e-5, swa_freq=5, mixup_alpha=0.3,
    optimizer=dict(lr=0.00002, adamw=True, weight_decay=0.0001, warm_up_len=5, ramp_down_start=50, ramp_down_len=50,
                   last_lr_value=0.01, schedule_mode="exp_lin", reaload_dataloaders_every_n_epochs=1),
)
module_ing.add_config(MODULE_DEFAULT_CONF)


def my_mixup(size, alpha):
    """helpers/mixup.py:5-12 — same host RNG calls in the same order (torch.randperm, then np.random.beta)."""
    rn_indices = torch.randperm(size)
    lambd = np.random.beta(alpha, alpha, size).astype(np.float32)
    lambd = np.concatenate([lambd[:, None], 1 - lambd[:, None]], 1).max(1)
    return rn_indices, torch.FloatTensor(lambd)
