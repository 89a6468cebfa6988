"""Deferred producers for the two tensors of the path whose only consumer is normally ``cam2mask``.

``denormalize_img`` (main.py:117) writes a [B,3,H,W] image that ``cam2mask`` reads once, and ``cam_validation``
(main.py:137) writes B*(C-1) planes of which ``cam2mask`` reads the two or three present classes.  Both functions
therefore return a :class:`LazyTensor`: a real ``torch.Tensor`` subclass with the right shape / dtype / device that
records how to compute itself.  ``cam2mask`` / ``cam_to_label`` of this package recognise it and fold the producer into
their first kernel (``cosa_cam2mask_ex``), so the intermediate never exists in HBM; ANY other use - a torch op, an
index, ``.cpu()``, printing, another function of this package - materialises it first through the stand-alone kernel
(``__torch_dispatch__``), after which it behaves as the plain tensor the reference would have produced.
The value is computed from the source tensor as it is when first needed: do not modify the source in place in between.
"""
import torch
from torch.utils._pytree import tree_map


class LazyTensor(torch.Tensor):
    """A tensor that is computed on first use.  ``kind`` names the producer ("denormalize_img" / "cam_validation"),
    ``sources`` the tensors (and constants) it is computed from, ``producer()`` the eager computation."""

    @staticmethod
    def __new__(cls, kind, sources, producer, like):
        r = torch.Tensor._make_wrapper_subclass(cls, like.shape, dtype=like.dtype, device=like.device,
                                                requires_grad=False)
        r._kind, r._sources, r._producer, r._value = kind, sources, producer, None
        return r

    def __init__(self, kind, sources, producer, like):
        pass

    @property
    def is_materialized(self):
        return self._value is not None

    def materialize(self):
        if self._value is None:
            self._value = self._producer()
            self._producer = None
        return self._value

    def __repr__(self):
        state = "materialized" if self._value is not None else "pending"
        return "LazyTensor(%s, %s, shape=%s)" % (self._kind, state, tuple(self.shape))

    __torch_function__ = torch._C._disabled_torch_function_impl

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        unwrap = lambda t: t.materialize() if isinstance(t, LazyTensor) else t
        return func(*tree_map(unwrap, args), **tree_map(unwrap, kwargs or {}))


def pending(t, kind):
    """The sources of ``t`` if it is a not yet materialised LazyTensor of the given kind, else None."""
    if isinstance(t, LazyTensor) and t._kind == kind and t._value is None:
        return t._sources
    return None


def plain(t):
    """``t`` itself, or the materialised value of a LazyTensor."""
    return t.materialize() if isinstance(t, LazyTensor) else t
