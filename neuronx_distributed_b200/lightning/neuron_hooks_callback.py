from .callbacks import NeuronHooksCallback  # noqa: F401  (reference module name)
