from .zero_redundancy_optimizer import NeuronEPZero1Optimizer, NeuronZero1Optimizer, Zero1Optimizer  # noqa: F401
