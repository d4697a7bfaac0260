"""Callbacks of the Lightning integration; the classes live in the modules the reference names them after."""
from .neuron_hooks_callback import NeuronHooksCallback  # noqa: F401
from .progress_bar import NeuronTQDMProgressBar  # noqa: F401
