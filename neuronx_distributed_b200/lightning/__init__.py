"""PyTorch-Lightning integration (reference ``lightning/__init__.py:1-8``): strategy, module, checkpoint IO, logger,
progress bar, hooks callback.  ``lightning`` is an optional dependency (not in this image): when it is importable the
classes derive from the Lightning base classes, otherwise from small stand-ins with the same method surface so that
the :class:`NeuronLTModule` training logic can also be driven by the plain loop in ``examples/``."""
from .module import NeuronLTModule  # noqa: F401
from .strategy import NeuronXLAStrategy, NxDStrategy  # noqa: F401
from .checkpoint_io import NeuronCheckpointIO  # noqa: F401
from .logger import NeuronTensorBoardLogger  # noqa: F401
from .callbacks import NeuronHooksCallback, NeuronTQDMProgressBar  # noqa: F401
from .accelerator import NeuronXLAAccelerator  # noqa: F401
from .precision_plugin import NeuronXLAPrecisionPlugin  # noqa: F401
from .trainer import Trainer  # noqa: F401
