"""Import the UNMODIFIED reference (palonso/MAEST) from $MAEST_REF (default /root/reference).

Container-only helper: the reference tree does not exist on the GPU box, so nothing in
`pytest -m gpu`, `smoke()` or `bench.py` may import this file.  It is used by
`tests/golden/make_golden.py` (fixture generation) and by the optional live-reference
tests, which skip when the tree is absent.

The reference needs two pure-Python packages that are not installed here:
  * `sacred.Ingredient`            (models/maest.py:21, models/module.py:8)
  * `timm.models.load_pretrained`  (models/helpers/vit_helpers.py:11, only called for pretrained=True)
and, for the Lightning module, `lightning.pytorch`.  We register tiny in-memory stand-ins
for them in `sys.modules`; no reference source is copied or modified.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("MAEST_REF", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "maest.py"))


class _Ingredient:
    """sacred.Ingredient stand-in: decorators are identities, config fns are recorded."""

    def __init__(self, name, *a, **k):
        self.name = name
        self.configs = []
        self.named = {}

    def config(self, f):
        self.configs.append(f)
        return f

    def named_config(self, f):
        self.named[f.__name__] = f
        return f

    def capture(self, f=None, prefix=None):
        if f is None:
            return lambda g: g
        return f

    def command(self, f=None, **k):
        if f is None:
            return lambda g: g
        return f

    def add_config(self, *a, **k):
        pass


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs(with_lightning: bool = True) -> None:
    import torch

    if "sacred" not in sys.modules:
        _mod("sacred", Ingredient=_Ingredient, Experiment=_Ingredient)
    if "timm" not in sys.modules:
        def load_pretrained(*a, **k):
            raise RuntimeError("timm.load_pretrained: no network / no weights in this sandbox")

        t = _mod("timm")
        t.models = _mod("timm.models", load_pretrained=load_pretrained)
    if with_lightning and "lightning" not in sys.modules:
        class LightningModule(torch.nn.Module):
            def log(self, *a, **k):
                pass

            def log_dict(self, *a, **k):
                pass

            def all_gather(self, x, *a, **k):
                return x

        class _CB:
            def __init__(self, *a, **k):
                pass

        class MisconfigurationException(Exception):
            pass

        pl = _mod("lightning.pytorch", LightningModule=LightningModule, Trainer=object)
        cb = _mod("lightning.pytorch.callbacks", ModelCheckpoint=_CB, StochasticWeightAveraging=_CB,
                  Callback=_CB)
        ut = _mod("lightning.pytorch.utilities")
        ex = _mod("lightning.pytorch.utilities.exceptions",
                  MisconfigurationException=MisconfigurationException)
        rz = _mod("lightning.pytorch.utilities.rank_zero", rank_zero_warn=lambda *a, **k: None,
                  rank_zero_info=lambda *a, **k: None)
        lg = _mod("lightning.pytorch.loggers", TensorBoardLogger=_CB)
        ut.exceptions = ex
        ut.rank_zero = rz
        pl.callbacks = cb
        pl.utilities = ut
        pl.loggers = lg
        top = _mod("lightning")
        top.pytorch = pl
        _mod("pytorch_lightning")


def load_reference_maest():
    """Return the reference `models` package imported under the name `maest` (pyproject.toml:34-38)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    install_stubs(with_lightning=False)
    if "maest" in sys.modules and getattr(sys.modules["maest"], "__ref__", False):
        return sys.modules["maest"]
    pkg_dir = os.path.join(REF_ROOT, "models")
    spec = importlib.util.spec_from_file_location(
        "maest", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules["maest"] = mod
    spec.loader.exec_module(mod)
    mod.__ref__ = True
    return mod


def load_reference_module():
    """Return the reference `models.module` (Lightning `Module`), `get_maest` left for the caller to patch."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    install_stubs(with_lightning=True)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("models.module")
