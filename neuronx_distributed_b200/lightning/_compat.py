"""Resolve Lightning base classes if the package exists; otherwise minimal stand-ins."""
try:  # pragma: no cover - lightning is not installed in the offline image
    import lightning.pytorch as pl
    from lightning.pytorch.callbacks import Callback
    from lightning.pytorch.strategies import DDPStrategy as _BaseStrategy
    from lightning.pytorch.plugins.io import CheckpointIO

    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:
    HAVE_LIGHTNING = False

    import torch

    class LightningModule(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.trainer = None
            self._logged = {}

        def log(self, name, value, **kw):
            self._logged[name] = value

        def optimizers(self):
            return getattr(self, "_optimizers", None)

        def lr_schedulers(self):
            return getattr(self, "_schedulers", None)

    class Callback:
        pass

    class _BaseStrategy:
        def __init__(self, *a, **k):
            pass

    class CheckpointIO:
        pass
