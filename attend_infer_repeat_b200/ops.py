"""Mirror of the reference's ops.py: Loss accumulator (ops.py:5-43) and clip_preserve (ops.py:67-76)."""
from __future__ import annotations

import torch


class Loss:
    """Keeps a scalar ``value`` and a per-sample ``[B]`` vector, both accumulated with weights (ops.py:5-43)."""

    def __init__(self):
        self._value = None
        self._per_sample = None

    def add(self, loss=None, per_sample=None, weight=1.0):
        if isinstance(loss, Loss):
            per_sample = loss.per_sample
            loss = loss.value
        self._update("_value", loss, weight)
        self._update("_per_sample", per_sample, weight)

    def _update(self, name, expr, weight):
        value = getattr(self, name)
        expr = expr * weight
        if value is None:
            value = expr
        else:
            assert tuple(value.shape) == tuple(expr.shape), \
                "Shape should be {} but is {}".format(tuple(value.shape), tuple(expr.shape))
            value = value + expr
        setattr(self, name, value)

    def _get_value(self, name):
        v = getattr(self, name)
        if v is None:
            v = torch.zeros([])
        return v

    @property
    def value(self):
        return self._get_value("_value")

    @property
    def per_sample(self):
        return self._get_value("_per_sample")


def clip_preserve(expr, min, max):
    """Clip in the forward pass, identity in the backward pass (ops.py:67-76)."""
    clipped = torch.clamp(expr, min, max)
    return (clipped - expr).detach() + expr
