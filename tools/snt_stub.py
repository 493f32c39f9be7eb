"""A tiny eager stand-in for the slice of Sonnet v1 (3fd7d9d) that the reference's cell.py / modules.py / neural.py /
mnist_model.py touch, backed by torch on the CPU.  Companion of tools/tf_stub.py, used only by tools/make_golden.py.

With it the reference's OWN AIRCell / AIRModel / AIRonMNIST source executes in this container, so everything those files
decide -- the (sx, tx, sy, ty) order, which bias goes where, the explore-eps mix, the presence product, what feeds the
LSTM, the canvas accumulation, the post-processing of model.py:83-104 -- comes from the reference itself.  What this file
supplies is the arithmetic INSIDE the Sonnet modules, restated from their documented semantics [upstream]:

  snt.Linear              y = x @ w + b, w [in, out]
  snt.LSTM                gates = [x, h] @ w_gates + b_gates; i, j, f, o = split(gates, 4);
                          c' = sigmoid(f + forget_bias) * c + sigmoid(i) * tanh(j); h' = tanh(c') * sigmoid(o);
                          state = (h, c); initial_state(trainable=True) tiles two [1, n] variables
  snt.AffineGridWarper    output grid linspace(-1, 1) per axis, source = A @ grid + t in normalised coordinates, pixel =
                          (coordinate + 1) * (size - 1) / 2; no_shear_2d constraints: parameters (sx, tx, sy, ty);
                          .inverse(): the inverse affine map, sampled on the source-sized grid
  snt.resampler           bilinear, each of the four taps is zero outside the image
  (the last two through torch's affine_grid / grid_sample with align_corners=True -- an implementation independent of
  the oracle's hand-written gather)

Variables live in one store keyed by the module path, like tf.make_template: a module built a second time (every RNN
step re-runs _build and re-creates the Affine layers inside it) gets the same names and therefore the same variables.
`VARIABLE_SOURCE(path, shape)` supplies their values, so the generator can feed the weights it stores in the vectors.
"""
import contextlib
import types

import torch
import torch.nn.functional as F

VARIABLES = {}            # path -> tensor, in creation order
VARIABLE_SOURCE = None    # callable(path, shape) -> tensor


class _Frame:
    def __init__(self, path):
        self.path = path
        self.counters = {}


_STACK = [_Frame("")]


def reset():
    VARIABLES.clear()
    del _STACK[1:]
    _STACK[0].counters.clear()


def get_variable(name, shape):
    path = _STACK[-1].path + "/" + name
    if path not in VARIABLES:
        v = VARIABLE_SOURCE(path, tuple(int(s) for s in shape))
        assert tuple(v.shape) == tuple(shape), (path, tuple(v.shape), tuple(shape))
        VARIABLES[path] = v
    return VARIABLES[path]


class AbstractModule:
    def __init__(self, name=None, **_):
        base = name or type(self).__name__
        parent = _STACK[-1]
        idx = parent.counters.get(base, 0)            # unique within the enclosing scope, as tf.variable_scope does
        parent.counters[base] = idx + 1
        self._scope = parent.path + "/" + base + ("" if idx == 0 else "_%d" % idx)

    @property
    def variable_scope(self):
        return types.SimpleNamespace(name=self._scope)

    @contextlib.contextmanager
    def _enter_variable_scope(self):
        _STACK.append(_Frame(self._scope))
        try:
            yield
        finally:
            _STACK.pop()

    def __call__(self, *args, **kwargs):
        _STACK.append(_Frame(self._scope))            # fresh counters: the same child names on every call
        try:
            return self._build(*args, **kwargs)
        finally:
            _STACK.pop()


class RNNCore(AbstractModule):
    pass


class Linear(AbstractModule):
    def __init__(self, output_size, use_bias=True, initializers=None, partitioners=None, regularizers=None,
                 custom_getter=None, name="linear"):
        # neural.py:49 passes its initializer dict in the `use_bias` position; it is truthy, so a bias is used and
        # Sonnet's default initialisers apply
        super().__init__(name)
        self._output_size = int(output_size)
        self._use_bias = bool(use_bias)

    def _build(self, inputs):
        w = get_variable("w", (inputs.shape[-1], self._output_size))
        out = inputs @ w
        if self._use_bias:
            out = out + get_variable("b", (self._output_size,))
        return out


class LSTM(RNNCore):
    def __init__(self, hidden_size, forget_bias=1.0, name="lstm"):
        super().__init__(name)
        self._hidden_size = int(hidden_size)
        self._forget_bias = float(forget_bias)

    @property
    def output_size(self):
        return [self._hidden_size]

    @property
    def state_size(self):
        return (self._hidden_size, self._hidden_size)

    def initial_state(self, batch_size, dtype=torch.float32, trainable=False, **_):
        if not trainable:
            z = torch.zeros(batch_size, self._hidden_size, dtype=dtype)
            return (z, z.clone())
        with self._enter_variable_scope():
            h0 = get_variable("initial_state_0", (1, self._hidden_size))
            c0 = get_variable("initial_state_1", (1, self._hidden_size))
        return (h0.repeat(batch_size, 1), c0.repeat(batch_size, 1))

    def _build(self, inputs, prev_state):
        prev_hidden, prev_cell = prev_state
        xh = torch.cat([inputs, prev_hidden], 1)
        w = get_variable("w_gates", (xh.shape[1], 4 * self._hidden_size))
        b = get_variable("b_gates", (4 * self._hidden_size,))
        gates = xh @ w + b
        i, j, f, o = torch.chunk(gates, 4, dim=1)
        next_cell = torch.sigmoid(f + self._forget_bias) * prev_cell + torch.sigmoid(i) * torch.tanh(j)
        next_hidden = torch.tanh(next_cell) * torch.sigmoid(o)
        return next_hidden, (next_hidden, next_cell)


class BatchFlatten(AbstractModule):
    def __init__(self, name="batch_flatten"):
        super().__init__(name)

    def _build(self, inputs):
        return inputs.reshape(inputs.shape[0], -1)


class BatchReshape(AbstractModule):
    def __init__(self, shape, name="batch_reshape"):
        super().__init__(name)
        self._shape = tuple(int(s) for s in shape)

    def _build(self, inputs):
        return inputs.reshape((inputs.shape[0],) + self._shape)


class Sequential(AbstractModule):
    def __init__(self, layers, name="sequential"):
        super().__init__(name)
        self._layers = list(layers)

    def _build(self, inputs):
        for layer in self._layers:
            inputs = layer(inputs)
        return inputs


class AffineWarpConstraints:
    def __init__(self, kind):
        self.kind = kind

    @classmethod
    def no_shear_2d(cls):
        return cls("no_shear_2d")


class _Grid:
    """What AffineGridWarper hands to snt.resampler: sampling positions in normalised [-1, 1] source coordinates."""
    def __init__(self, grid):
        self.grid = grid      # [B, out_h, out_w, 2] (x, y)


class AffineGridWarper(AbstractModule):
    def __init__(self, source_shape, output_shape, constraints=None, name="affine_grid_warper", _inverse=False):
        super().__init__(name)
        assert constraints is not None and constraints.kind == "no_shear_2d", "only the constraint the reference uses"
        self._source_shape = tuple(int(s) for s in source_shape)
        self._output_shape = tuple(int(s) for s in output_shape)
        self._constraints = constraints
        self._inverse = _inverse

    def inverse(self, name=None):
        # the inverse warper maps positions of the SOURCE-sized grid back into the output (glimpse) frame
        return AffineGridWarper(self._output_shape, self._source_shape, self._constraints,
                                name or "inverse_affine_grid_warper", _inverse=not self._inverse)

    def _build(self, inputs):
        sx, tx, sy, ty = (inputs[:, k] for k in range(4))
        zero = torch.zeros_like(sx)
        if self._inverse:
            # [[sx, 0], [0, sy]]^-1 and -A^-1 t, through the determinant as Sonnet computes it
            det = sx * sy
            a, d = sy / det, sx / det
            theta = torch.stack([torch.stack([a, zero, -(a * tx)], 1), torch.stack([zero, d, -(d * ty)], 1)], 1)
        else:
            theta = torch.stack([torch.stack([sx, zero, tx], 1), torch.stack([zero, sy, ty], 1)], 1)
        h, w = self._output_shape
        grid = F.affine_grid(theta, (inputs.shape[0], 1, h, w), align_corners=True)
        return _Grid(grid)


def resampler(data, warp, name="resampler"):
    """data [B, H, W, C], warp from AffineGridWarper -> [B, h, w, C]."""
    out = F.grid_sample(data.permute(0, 3, 1, 2), warp.grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.permute(0, 2, 3, 1)


def module():
    m = types.ModuleType("sonnet")
    for k in ("AbstractModule", "RNNCore", "Linear", "LSTM", "BatchFlatten", "BatchReshape", "Sequential",
              "AffineWarpConstraints", "AffineGridWarper", "resampler"):
        setattr(m, k, globals()[k])
    return m
