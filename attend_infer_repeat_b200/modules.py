"""Mirror of the reference's modules.py (modules.py:11-143) plus the transition core (snt.LSTM, mnist_model.py:35).

Same class names, constructor arguments and meaning as the reference.  The classes are layer descriptors consumed by
AIRCell, which lowers the whole cell into one air_config for the fused CUDA path; bound to parameters they can also be
called one by one (each call = stand-alone C-ABI launches), which is what the unit-parity tests do.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import functional as F
from .neural import MLP, _flatten


class LSTM:
    """snt.LSTM(hidden_size) [upstream Sonnet v1.1]: gates = [x,h] @ W + b, order (i, j, f, o), forget_bias 1.0,
    no peepholes.  Only the attributes AIRCell reads are kept: output_size, state_size, initial_state."""

    def __init__(self, hidden_size: int, forget_bias: float = 1.0):
        self._hidden_size = int(hidden_size)
        self.forget_bias = float(forget_bias)
        self._views = None

    @property
    def output_size(self):
        return (self._hidden_size,)          # cell.py:44 reads output_size[0]

    @property
    def state_size(self):
        return (self._hidden_size, self._hidden_size)

    def bind(self, views, prefix="lstm"):
        self._views, self._prefix = views, prefix
        return self

    def initial_state(self, batch_size, dtype=torch.float32, trainable=True):
        """Trainable (h0, c0) of shape [1, nh] tiled to the batch (cell.py:103)."""
        h0 = self._views[f"{self._prefix}.h0"].reshape(1, -1).expand(batch_size, -1).contiguous()
        c0 = self._views[f"{self._prefix}.c0"].reshape(1, -1).expand(batch_size, -1).contiguous()
        return h0, c0

    def __call__(self, inpt, state):
        h, c = F.lstm_step(inpt, state[0], state[1], self._views[f"{self._prefix}.w"],
                           self._views[f"{self._prefix}.b"], self.forget_bias)
        return h, (h, c)


class ParametrisedGaussian:
    """modules.py:11-24: Linear(2 n) -> loc, softplus(scale_raw + scale_offset)."""

    def __init__(self, n_params, scale_offset=0.0):
        self._n_params = int(n_params)
        self._scale_offset = float(scale_offset)


class TransformParam:
    """modules.py:27-50.  (The reference base class recurses forever in _build -- modules.py:49; as there, only the
    stochastic subclass is usable.)"""

    def __init__(self, n_hidden, n_param, max_crop_size=1.0):
        self._n_hidden = _flatten(n_hidden)
        self._n_param = int(n_param)
        self._max_crop_size = float(max_crop_size)


class StochasticTransformParam(TransformParam):
    """modules.py:53-63: MLP -> 2*n_param; loc = (sig, tanh, sig, tanh) squash; scale_raw + scale_bias."""

    def __init__(self, n_hidden, n_param, max_crop_size=1.0, scale_bias=-2.0):
        super().__init__(n_hidden, n_param * 2, max_crop_size)
        self._scale_bias = scale_bias


class Encoder:
    """modules.py:66-76: flatten -> MLP(n_hidden), ELU everywhere."""

    def __init__(self, n_hidden):
        self._n_hidden = _flatten(n_hidden)
        self.mlp = MLP(self._n_hidden)

    def bind(self, views, prefix):
        self.mlp.bind(views, prefix)
        return self

    def __call__(self, inpt):
        return self.mlp(inpt)


class Decoder:
    """modules.py:79-91: MLP(n_hidden, n_out=prod(output_size)) with a linear output layer, reshaped."""

    def __init__(self, n_hidden, output_size):
        self._n_hidden = _flatten(n_hidden)
        self._output_size = tuple(int(i) for i in output_size)
        self.mlp = MLP(self._n_hidden, n_out=int(np.prod(self._output_size)))

    def bind(self, views, prefix):
        self.mlp.bind(views, prefix)
        return self

    def __call__(self, inpt):
        return self.mlp(inpt).reshape((inpt.shape[0],) + self._output_size)


class SpatialTransformer:
    """modules.py:94-109: snt.AffineGridWarper(img_size, crop_size, no_shear_2d) [+ .inverse()] + snt.resampler.
    transform_params = (sx, tx, sy, ty)."""

    def __init__(self, img_size, crop_size, constraints=None, inverse=False):
        self._img_size = tuple(int(i) for i in img_size)
        self._crop_size = tuple(int(i) for i in crop_size)
        self._inverse = bool(inverse)

    def __call__(self, img, transform_params):
        if img.dim() == 4:
            img = img[..., 0]
        if self._inverse:
            return F.stn_paint(img, transform_params, self._img_size)
        return F.stn_read(img, transform_params, self._crop_size)


class StepsPredictor:
    """modules.py:112-122: sigmoid(MLP(n_hidden, n_out=1) + steps_bias)."""

    def __init__(self, n_hidden, steps_bias=0.0):
        self._n_hidden = _flatten(n_hidden)
        self._steps_bias = steps_bias


class BaselineMLP:
    """modules.py:125-143: concat[img, what, where, presence (batch-major), state] -> MLP(n_hidden, n_out=1)."""

    def __init__(self, n_hidden):
        self._n_hidden = _flatten(n_hidden)
        self.mlp: Optional[MLP] = None
        self.params: Optional[torch.Tensor] = None
        self.views = None

    def _build_params(self, n_in, device, seed=0):
        from .cell import _init_flat
        spec, d = [], n_in
        for i, n in enumerate(self._n_hidden):
            spec += [(f"baseline.{i}.w", (d, n)), (f"baseline.{i}.b", (1, n))]
            d = n
        spec += [("baseline.out.w", (d, 1)), ("baseline.out.b", (1, 1))]
        self._spec = spec
        self.params, self.views = _init_flat(spec, device, seed)
        self.mlp = MLP(self._n_hidden, n_out=1).bind(self.views, "baseline")
        # gradient buffer + optimiser slots of the baseline's own optimiser (model.py:362-367; RMSProp slots: ms = 1)
        from .engine import make_views
        self.grad = torch.zeros_like(self.params)
        self.grad_views = make_views(spec, self.grad)
        self.slots = dict(mg=torch.zeros_like(self.params), ms=torch.ones_like(self.params),
                          mom=torch.zeros_like(self.params))

    def attach(self, engine):
        """Run on `engine` (Engine.baseline_attach, before train_enable): the input rows are gathered from the engine's own
        cell outputs in one pass and the big first layer and its weight gradient run on the engine's GEMM path."""
        n_in, n_params = engine.baseline_attach(self._n_hidden)
        if self.params is None:
            self._build_params(n_in, engine.device)
        assert self.params.numel() == n_params, (self.params.numel(), n_params)
        self._engine = engine
        return self

    def backward(self, target, baseline, target_mean=None, inv_batch=None, defer_join=False):
        """Gradient of baseline_loss = .5 * mean((stop_gradient(target) - baseline)^2) (model.py:253-259; target [B],
        baseline [B,1] -> [B,B] broadcast) with respect to the baseline's parameters, for the LAST
        call; left in ``self.grad`` (flat).  ``target_mean`` / ``inv_batch`` are the global-batch values under sharding."""
        from . import functional as F
        B = baseline.shape[0]
        ib = 1.0 / B if inv_batch is None else inv_batch
        if isinstance(target_mean, torch.Tensor):      # a device scalar (e.g. scalars[mean_iw]): no host round trip
            d_out = F.baseline_grad_dev(baseline, target_mean, ib)
        else:
            tm = float(target.mean()) if target_mean is None else float(target_mean)
            d_out = F.baseline_grad(target, baseline, tm, ib)
        self._d_out = d_out     # (kept alive: with defer_join the side streams still read it after this call returns)
        if getattr(self, "_engine", None) is not None:
            # defer_join: self.grad is complete after the engine's next backward() (Engine.baseline_backward)
            return self._engine.baseline_backward(self.params, d_out, self.grad, defer_join=defer_join)
        self.grad.zero_()
        self.mlp.backward(d_out, self.grad_views)
        return self.grad

    def __call__(self, img, what, where, presence_prob, state=None):
        if getattr(self, "_engine", None) is not None:
            # (what / where / presence / state are the engine's own output buffers of the last forward)
            return self._engine.baseline_forward(self.params, img.reshape(img.shape[0], -1))
        B = img.shape[0]
        parts = [t.transpose(0, 1).reshape(B, -1) for t in (what, where, presence_prob)]
        if state is not None:
            parts += list(state)
        x = torch.cat([img.reshape(B, -1)] + parts, -1).contiguous()
        if self.mlp is None:
            self._build_params(x.shape[1], img.device)
        return self.mlp(x)
