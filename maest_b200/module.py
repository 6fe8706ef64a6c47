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


class Module(_LightningModule):
    """Drop-in for the reference's Lightning `Module` (models/module.py:44-276) for the hot path: `net`, `training_step`,
    `forward`, `predict_step`, `configure_optimizers`, and the validation / test path (`validation_step`, `test_step`,
    `on_validation_epoch_end`, `on_test_epoch_end`: twin-net evaluation with `net_swa` when the SWA callback created it, macro
    AP / ROC AUC computed on the device instead of scikit-learn on host copies)."""

    @module_ing.capture
    def __init__(self, do_swa=True, swa_epoch_start=50, swa_lrs=2e-5, swa_freq=5, mixup_alpha=0.3, distributed_mode=False,
                 optimizer=None, net=None):
        super().__init__()
        self.mixup_alpha = mixup_alpha
        self.do_swa = do_swa
        self.swa_freq = swa_freq
        self.swa_epoch_start = swa_epoch_start
        self.swa_lrs = swa_lrs
        self.distributed_mode = distributed_mode
        self.optimizer_cfg = dict(MODULE_DEFAULT_CONF["optimizer"], **(optimizer or {}))
        self.net = net if net is not None else get_maest()      # models/module.py:63 (arguments come from the Sacred ingredient)
        self.validation_outputs = []
        self.test_outputs = []
        self.transformer_block = -1

    def forward(self, batch, transformer_block=-1):
        # the reference ignores `transformer_block` here and always passes -1 (models/module.py:68-71); kept as is
        return self.net.forward(batch, transformer_block=-1, return_self_attention=False)

    def training_step(self, batch, batch_idx):
        from .train import training_forward
        x, f, y = batch
        mix = None
        if self.mixup_alpha > 0:
            mix = my_mixup(len(y), self.mixup_alpha)             # host RNG: torch.randperm, then np.random.beta
        loss, _ = training_forward(self.net, x, y, mix)
        self.log("train_loss", loss, on_step=True, on_epoch=True, prog_bar=True, logger=True, batch_size=len(y), sync_dist=True)
        return loss

    def predict_step(self, batch, batch_idx: int, dataloader_idx: int = None):
        x, f, y = batch
        logits, embed = self.forward(x, transformer_block=self.transformer_block)
        return {"logits": logits.detach().cpu(), "embeddings": embed.detach().cpu(), "filename": f}

    # ---- validation / test (models/module.py:121-216) -----------------------------------------------------------------
    @staticmethod
    def _join(strings):
        return "_".join(filter(lambda x: x, strings))

    def _net_map(self):
        net_map = [(None, self.net)]
        if self.do_swa and hasattr(self, "net_swa"):       # helpers/swa_callback.py:43-44 creates net_swa on fit start
            net_map.append(("swa", self.net_swa))
        return net_map

    def test_validation_step(self, batch, batch_idx, output_buffer, stage):
        from . import ops
        x, f, y = batch
        outputs = {"y": y.detach()}
        for name, net in self._net_map():
            with torch.no_grad():
                logits, _ = net(x)
            loss, _ = ops.bce_logits(logits, y.to(logits.device))       # F.binary_cross_entropy_with_logits(...).mean()
            outputs[self._join((name, "loss"))] = loss
            outputs[self._join((name, "y_hat"))] = torch.sigmoid(logits.detach())
            self.log(self._join((stage, "loss", name)), loss, batch_size=len(y), sync_dist=True)
        output_buffer.append(outputs)
        return outputs

    def validation_step(self, batch, batch_idx):
        return self.test_validation_step(batch, batch_idx, self.validation_outputs, "val")

    def test_step(self, batch, batch_idx):
        return self.test_validation_step(batch, batch_idx, self.test_outputs, "test")

    def on_test_validation_epoch_end(self, outputs, stage):
        """models/module.py:155-205 with the predictions kept on the GPU: per-class AP / ROC AUC by `ops.ap_roc` (sort +
        threshold-scan kernel), macro-averaged.  Like scikit-learn (>= 1.6), a class with a single label value makes the macro
        ROC AUC nan (with a warning)."""
        from . import ops
        result = {}
        if not outputs:
            return result
        y = torch.cat([o["y"] for o in outputs], dim=0)
        if self.distributed_mode:
            y = self.all_gather(y).reshape(-1, y.shape[-1])
        for name, _ in self._net_map():
            loss = torch.stack([o[self._join((name, "loss"))] for o in outputs]).mean()
            y_hat = torch.cat([o[self._join((name, "y_hat"))] for o in outputs], dim=0)
            if self.distributed_mode:
                loss = self.all_gather(loss).mean()
                y_hat = self.all_gather(y_hat).reshape(-1, y_hat.shape[-1])
            ap, auc, _ = ops.ap_roc(y.to(y_hat.device), y_hat)
            if bool(torch.isnan(auc).any()):
                import warnings
                warnings.warn("Only one class is present in y_true for some label. ROC AUC score is not defined in that case.")
            result.update({self._join((stage, "loss", name)): float(loss), self._join((stage, "ap", name)): float(ap.mean()),
                           self._join((stage, "roc", name)): float(auc.mean())})
        self.log_dict(result, sync_dist=True)
        outputs.clear()
        return result

    def on_validation_epoch_end(self):
        return self.on_test_validation_epoch_end(self.validation_outputs, "val")

    def on_test_epoch_end(self):
        return self.on_test_validation_epoch_end(self.test_outputs, "test")

    def set_prediction_tranformer_block(self, transformer_block):
        self.transformer_block = transformer_block

    def get_optimizer(self, params):
        """models/module.py:236-243.  On CUDA parameters the AdamW step is one fused launch (maest_b200.optim.FusedAdamW, same
        arithmetic); `fused_optimizer=False` in the optimizer config keeps torch.optim.AdamW."""
        cfg = self.optimizer_cfg
        params = list(params)
        if not cfg["adamw"]:
            return torch.optim.Adam(params, lr=cfg["lr"])
        if cfg.get("fused_optimizer", True) and params and params[0].is_cuda:
            from .optim import FusedAdamW
            return FusedAdamW(params, lr=cfg["lr"], weight_decay=cfg["weight_decay"])
        return torch.optim.AdamW(params, lr=cfg["lr"], weight_decay=cfg["weight_decay"])

    def get_scheduler_lambda(self):
        from .optim import get_scheduler_lambda
        c = self.optimizer_cfg
        return get_scheduler_lambda(c["warm_up_len"], c["ramp_down_start"], c["ramp_down_len"], c["last_lr_value"], c["schedule_mode"])

    def get_lr_scheduler(self, optimizer):
        if self.optimizer_cfg["schedule_mode"] in {"exp_lin", "cos_cyc"}:       # models/module.py:225-231
            return torch.optim.lr_scheduler.LambdaLR(optimizer, self.get_scheduler_lambda())
        raise RuntimeError(f"schedule_mode={self.optimizer_cfg['schedule_mode']} Unknown.")

    def configure_optimizers(self):
        # models/module.py:245-254: {"optimizer": AdamW over all parameters, "lr_scheduler": LambdaLR(epoch lambda)}
        optimizer = self.get_optimizer(self.parameters())
        return {"optimizer": optimizer, "lr_scheduler": self.get_lr_scheduler(optimizer)}


def allreduce_gradients(module: torch.nn.Module, group=None) -> int:
    """Data-parallel gradient averaging for one process per GPU WITHOUT the DDP wrapper: one flat NCCL all-reduce
    (NVLink/NVSwitch) over every parameter that received a gradient; parameters without one (`head_dist.*` in "mean"
    mode) are skipped on every rank, which is what `find_unused_parameters=True` does in the reference
    (ex_maest.py:57).  Returns the number of gradient elements reduced."""
    import torch.distributed as dist

    ps = [p for p in module.parameters() if p.grad is not None]
    if not ps:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for p in ps:
        n = p.numel()
        p.grad.copy_(flat[off: off + n].view_as(p.grad))
        off += n
    return off
