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
    from lightning.pytorch.callbacks import Callback as _Callback, ModelCheckpoint
    _LightningModule = pl.LightningModule
    _HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _HAVE_LIGHTNING = False

    class _Callback:            # stand-ins with the constructor surface `configure_callbacks` uses (models/module.py:256-276)
        pass

    class ModelCheckpoint(_Callback):
        def __init__(self, monitor=None, mode="min", filename=None, every_n_epochs=None, **kw):
            self.monitor, self.mode, self.filename, self.every_n_epochs = monitor, mode, filename, every_n_epochs

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
        # models/module.py:63 (arguments come from the Sacred ingredient).  The module trains, so its net defaults to bf16
        # operands (what the north star names; no loss scaling needed); fp16 operands are opt-in: pass net=get_maest(op_dtype="fp16").
        self.net = net if net is not None else get_maest(op_dtype="bf16")
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

    def configure_callbacks(self):
        """models/module.py:256-276: best-val-loss checkpoint, per-epoch checkpoint, and (do_swa) the weight-averaging callback
        that creates `net_swa` on fit start and refreshes it at every epoch end."""
        callbacks = [ModelCheckpoint(monitor="val_loss", mode="min", filename="{epoch}-{val_loss:.2f}-best"),
                     ModelCheckpoint(filename="{epoch}", every_n_epochs=1)]
        if self.do_swa:
            callbacks.append(StochasticWeightAveragingAndCopy(swa_lrs=self.swa_lrs, swa_epoch_start=self.swa_epoch_start))
        return callbacks


class StochasticWeightAveragingAndCopy(_Callback):
    """The reference's SWA callback (helpers/swa_callback.py:11-45 on top of Lightning's StochasticWeightAveraging) without
    the second full-model copy: `net_swa` IS the running average.  From `swa_epoch_start` on, once per epoch (after the last
    optimiser step of the epoch, which is when Lightning's callback calls `update_parameters`), the fused optimiser folds the
    current weights into `net_swa` inside its own kernel (`FusedAdamW.step(update_swa=True)`: avg += (p - avg) / (n + 1),
    torch.optim.swa_utils semantics); before that epoch `net_swa` tracks `net` exactly, as `transfer_weights` of a freshly
    deep-copied average model does in the reference.  With a non-fused optimiser the same update runs as torch ops.

    Not reproduced (host-side training-loop policy, out of the hot path): the SWA learning-rate annealing (`swa_lrs`,
    `annealing_epochs`) and the final swap of the averaged weights into `net` at fit end (the reference keeps both nets and
    validates both, models/module.py:121-146, which is what this callback serves)."""

    def __init__(self, swa_lrs=2e-5, swa_epoch_start=50, **_):
        self.swa_lrs, self._swa_epoch_start = swa_lrs, swa_epoch_start
        self.n_averaged = 0

    def on_fit_start(self, trainer, pl_module):
        import copy
        max_epochs = getattr(trainer, "max_epochs", None)
        if isinstance(self._swa_epoch_start, float) and max_epochs:
            self._swa_epoch_start = int(max_epochs * self._swa_epoch_start)      # helpers/swa_callback.py:31-32
        if not hasattr(pl_module, "net_swa"):
            pl_module.net_swa = copy.deepcopy(pl_module.net)                     # helpers/swa_callback.py:43-44
        for p in pl_module.net_swa.parameters():
            p.requires_grad_(False)

    @staticmethod
    def _optimizer(trainer):
        opts = getattr(trainer, "optimizers", None) or []
        return opts[0] if opts else None

    def on_train_epoch_start(self, trainer, pl_module):
        """Arms the fused path: the LAST optimiser step of an averaging epoch also updates net_swa."""
        self._epoch = int(getattr(trainer, "current_epoch", 0))

    def averaging(self):
        return getattr(self, "_epoch", 0) >= int(self._swa_epoch_start)

    @torch.no_grad()
    def on_train_epoch_end(self, trainer, pl_module):
        from .optim import FusedAdamW
        if not hasattr(pl_module, "net_swa"):
            self.on_fit_start(trainer, pl_module)
        src, dst = list(pl_module.net.parameters()), list(pl_module.net_swa.parameters())
        if not self.averaging():
            for d, s_ in zip(dst, src):              # before swa_epoch_start the average model is a copy of the net
                d.copy_(s_)
            return
        opt = self._optimizer(trainer)
        if isinstance(opt, FusedAdamW) and src and src[0].is_cuda:
            opt.fold_into_swa(src, dst, self.n_averaged)          # one launch over all tensors
        else:
            for d, s_ in zip(dst, src):
                d.add_((s_ - d) / (self.n_averaged + 1))
        self.n_averaged += 1


class TeacherStudentModule(Module):
    """models/module.py:279-349: the student has two heads (`distilled_type="separated"`): `head` on the CLS token is trained on
    the ground truth, `head_dist` on the DIST token on the teacher's soft labels; loss = (BCE + BCE_teacher) / 2."""

    def training_step(self, batch, batch_idx):
        from .train import training_forward
        x, f, y, y_teacher = batch
        mix = None
        if self.mixup_alpha > 0:
            mix = my_mixup(len(y), self.mixup_alpha)             # one draw blends x, y and y_teacher (models/module.py:284-295)
        loss, _, _, loss_standard, loss_teacher = training_forward(self.net, x, (y, y_teacher), mix)
        self.log_dict({"train_loss": loss, "train_loss_standard": loss_standard, "tran_loss_teacher": loss_teacher},
                      on_step=True, on_epoch=True, prog_bar=True, logger=True)
        return loss

    def test_validation_step(self, batch, batch_idx, output_buffer, stage):
        # models/module.py:317-349.  Like the reference, `logits, _ = net(x)` requires a net whose forward returns two values;
        # a "separated" net returns three, so the first (CLS) logits are used and the unpacking mismatch of the reference
        # (ValueError: too many values to unpack) is not reproduced.
        from . import ops
        x, f, y, y_teacher = batch
        outputs = {"y": y.detach(), "y_teacher": y_teacher.detach()}
        for name, net in self._net_map():
            with torch.no_grad():
                logits = net(x)[0]
            loss_standard, _ = ops.bce_logits(logits, y.to(logits.device))
            loss_teacher, _ = ops.bce_logits(logits, y_teacher.to(logits.device))
            loss = (loss_standard + loss_teacher) / 2
            outputs[self._join((name, "loss_standard"))] = loss_standard
            outputs[self._join((name, "loss_teacher"))] = loss_teacher
            outputs[self._join((name, "loss"))] = loss
            outputs[self._join((name, "y_hat"))] = torch.sigmoid(logits.detach())
            self.log_dict({self._join((stage, "loss_standard", name)): loss_standard,
                           self._join((stage, "loss_teacher", name)): loss_teacher, self._join((stage, "loss", name)): loss})
        output_buffer.append(outputs)
        return outputs


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
