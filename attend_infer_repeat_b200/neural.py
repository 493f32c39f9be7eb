"""Mirror of the reference's neural.py (neural.py:42-102): Affine / MLP.

In the reference these are Sonnet modules that add ops to a TF graph.  Here they are *layer descriptors*: they
record the layer widths, and -- once an owner (AIRCell) has bound them to slices of its flat parameter buffer -- they
can also be called stand-alone, in which case each layer is one ``air_linear`` launch of the CUDA library.  On the hot
path the modules are never called one by one: AIRCell hands the whole layer table to the fused C-ABI call.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Union

import torch

from . import functional as F


def _flatten(x) -> List[int]:
    if isinstance(x, Iterable):
        return [int(i) for i in x]
    return [int(x)]


class MLP:
    """neural.py:63-102.  ``n_hiddens`` int or iterable; ELU hidden layers; optional linear output of size n_out."""

    def __init__(self, n_hiddens, n_out: Optional[int] = None, name: str = "mlp"):
        self._n_hiddens = _flatten(n_hiddens)
        self._n_out = n_out
        self.name = name
        self._views: Optional[Dict[str, torch.Tensor]] = None
        self._prefix: Optional[str] = None

    @property
    def n_hiddens(self) -> List[int]:
        return list(self._n_hiddens)

    @property
    def output_size(self) -> int:
        return self._n_out if self._n_out is not None else self._n_hiddens[-1]

    def bind(self, views: Dict[str, torch.Tensor], prefix: str):
        """Attach the parameter views (``prefix.i.w`` [in,out], ``prefix.i.b``, ``prefix.out.w/b``)."""
        self._views, self._prefix = views, prefix
        return self

    def __call__(self, inpt: torch.Tensor) -> torch.Tensor:
        if self._views is None:
            raise RuntimeError("MLP is not bound to parameters (build it through AIRCell / AIRModel)")
        x = inpt.reshape(inpt.shape[0], -1)
        self._saved = [x]          # layer inputs, kept for backward()
        for i in range(len(self._n_hiddens)):
            x = F.linear(x, self._views[f"{self._prefix}.{i}.w"], self._views[f"{self._prefix}.{i}.b"], F.ACT_ELU)
            self._saved.append(x)
        if self._n_out is not None:
            x = F.linear(x, self._views[f"{self._prefix}.out.w"], self._views[f"{self._prefix}.out.b"], F.ACT_NONE)
        return x

    def backward(self, d_out: torch.Tensor, grad_views: Dict[str, torch.Tensor]):
        """Accumulate d loss / d parameters of the LAST call into ``grad_views`` (same names as the parameter views, zeroed
        by the caller); d_out is the gradient at the MLP's output (requires the linear output layer, n_out)."""
        if self._n_out is None:
            raise NotImplementedError("backward() is built for MLPs with a linear output layer (BaselineMLP)")
        names = [f"{self._prefix}.{i}" for i in range(len(self._n_hiddens))] + [f"{self._prefix}.out"]
        dy = d_out
        for li in range(len(names) - 1, -1, -1):
            x = self._saved[li]
            dy = F.linear_backward(x, self._views[names[li] + ".w"], dy, grad_views[names[li] + ".w"],
                                   grad_views[names[li] + ".b"], need_dx=li > 0, x_is_elu_output=li > 0)
