"""Process launching (reference ``lightning/launcher.py`` subclasses Lightning's XLA spawner).  Ranks are created by
``torchrun`` (one process per GPU, ``RANK`` / ``LOCAL_RANK`` / ``WORLD_SIZE`` in the environment); this launcher just runs
the function in the current process, which is what Lightning's ``_SubprocessScriptLauncher`` does for externally launched
jobs."""
from __future__ import annotations

from typing import Any, Callable


class _NeuronXLALauncher:
    is_interactive_compatible = False

    def __init__(self, strategy=None) -> None:
        self._strategy = strategy

    def launch(self, function: Callable, *args: Any, trainer=None, **kwargs: Any) -> Any:
        return function(*args, **kwargs)
