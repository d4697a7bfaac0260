"""Flatten arbitrary python containers into (tensor list, skeleton) and back — role of reference
``utils/serialization.py:36-254`` (``SerializationManager``, ``TensorMeta``, ``find_loss_from_output_and_spec``).
Built on ``torch.utils._pytree`` instead of a hand-written recursive walker."""
from __future__ import annotations

from typing import Any, List, NamedTuple, Tuple

import torch
from torch.utils import _pytree as pytree


class TensorMeta(NamedTuple):
    """What a receiver needs to allocate a tensor before its payload arrives (reference :73-83; the pipeline's p2p metadata,
    ``pipeline/comm.py``, carries these)."""
    tensor_index: int
    dtype: torch.dtype
    shape: torch.Size
    requires_grad: bool
    device: torch.device


class TensorStub:
    """Placeholder left in the skeleton where a tensor was."""

    def __init__(self, index: int):
        self.index = index

    def __repr__(self) -> str:
        return f"TensorStub({self.index})"


class SerializationManager:
    def serialize(self, obj: Any, return_stub_list: bool = False):
        """``(skeleton, tensors)``; with ``return_stub_list=True`` also the :class:`TensorMeta` of every tensor — the
        reference's three-tuple (:103-129; it defaults to returning it, the two-tuple is this package's default because the
        pipeline engine keeps the metadata in its own cache)."""
        if return_stub_list:
            skeleton, tensors = self.serialize(obj)
            return skeleton, tensors, self.tensor_metas(tensors)
        leaves, spec = pytree.tree_flatten(obj)
        tensors: List[torch.Tensor] = []
        stubs = []
        for leaf in leaves:
            if isinstance(leaf, torch.Tensor):
                stubs.append(TensorStub(len(tensors)))
                tensors.append(leaf)
            else:
                stubs.append(leaf)
        return (stubs, spec), tensors

    def tensor_metas(self, tensors: List[torch.Tensor]) -> List[TensorMeta]:
        """The third element of the reference's ``serialize`` result: one :class:`TensorMeta` per extracted tensor."""
        return [TensorMeta(i, t.dtype, t.shape, t.requires_grad, t.device) for i, t in enumerate(tensors)]

    def extract_stubs(self, stubbed_obj: Any) -> List[TensorStub]:
        """The tensor placeholders of a serialised skeleton, in tensor order (what a receiver must allocate)."""
        stubs, _ = stubbed_obj
        return [s for s in stubs if isinstance(s, TensorStub)]

    @staticmethod
    def catch_and_raise_for_large_object(obj: Any):
        """Context manager turning a ``RecursionError`` while walking a very deep object into a ``RuntimeError`` naming
        the object's class (reference :96-101)."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            try:
                yield
            except RecursionError:
                raise RuntimeError(obj.__class__.__name__) from None

        return ctx()

    def deserialize(self, stubbed_obj: Any, tensors: List[torch.Tensor]) -> Any:
        stubs, spec = stubbed_obj
        leaves = [tensors[s.index] if isinstance(s, TensorStub) else s for s in stubs]
        return pytree.tree_unflatten(leaves, spec)


def find_loss_from_output_and_spec(output_val: Any, spec_val: Any) -> torch.Tensor:
    """Pick the loss out of a model output using a spec of the same structure whose leaves are
    booleans (exactly one True), e.g. ``(True, False)`` or ``{"loss": True}``; ``True`` alone means
    the output itself (reference :212-254)."""
    output, spec = output_val, spec_val
    if spec is True or spec is None:
        if isinstance(output, torch.Tensor):
            return output
        if hasattr(output, "loss"):
            return output.loss
        if isinstance(output, (tuple, list)):
            return output[0]
        raise ValueError("cannot infer loss from output")
    if isinstance(spec, dict):
        for k, v in spec.items():
            if v is not False and v is not None:
                sub = output[k] if isinstance(output, dict) else getattr(output, k)
                return find_loss_from_output_and_spec(sub, v)
    if isinstance(spec, (tuple, list)):
        for i, v in enumerate(spec):
            if v is not False and v is not None:
                return find_loss_from_output_and_spec(output[i], v)
    raise ValueError(f"no loss selected by spec {spec!r}")


def compress_to_string(obj: Any) -> str:
    """Pickle → base64 text (pipeline metadata travels between stages as strings; reference :14-20)."""
    import base64
    import pickle

    return base64.b64encode(pickle.dumps(obj)).decode("utf-8")


def uncompress_from_string(serialized_data_string: str) -> Any:
    import base64
    import pickle

    return pickle.loads(base64.b64decode(serialized_data_string.encode("utf-8")))


def is_instance_namedtuple(iterable: Any) -> bool:
    return isinstance(iterable, tuple) and hasattr(iterable, "_fields") and type(iterable).__base__ is tuple
