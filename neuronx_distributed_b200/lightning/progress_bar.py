from .callbacks import NeuronTQDMProgressBar  # noqa: F401  (reference module name)
