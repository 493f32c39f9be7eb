"""A tiny eager stand-in for the slice of TensorFlow 1.x / Sonnet v1 that the reference's prior.py, ops.py and the
loss half of model.py touch, backed by torch on the CPU.

Purpose (tools/make_golden.py): the reference cannot run here (no TensorFlow, no Sonnet, Python-2 sources), but those
three files are plain Python that only *compose* TF primitives.  Installing this module as ``tensorflow`` lets the
reference's OWN source files execute in this container, so the composition logic -- which is exactly what an oracle
could get wrong (argument order, axis conventions, masking, the [B]-[B,1] broadcast, float64 islands) -- comes from
the reference itself.  Each primitive below restates the documented TF 1.1 semantics it replaces [upstream].

Not product code, not imported by the package, the tests or the bench: only by tools/make_golden.py.
"""
import contextlib
import math
import sys
import types

import numpy as np
import torch

float32, float64, int32, int64, bool_ = torch.float32, torch.float64, torch.int32, torch.int64, torch.bool
Tensor = torch.Tensor


builtins_range = range      # tf.range shadows the builtin further down


def convert_to_tensor(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, (bool, np.bool_)):
        return torch.tensor(bool(x))
    if isinstance(x, (int, np.integer)):
        return torch.tensor(int(x), dtype=dtype or int32)
    if isinstance(x, (float, np.floating)):
        return torch.tensor(float(x), dtype=dtype or float32)        # python floats become float32 constants
    a = np.asarray(x)
    if a.dtype == np.float64 and dtype is None:
        dtype = float32
    t = torch.as_tensor(a)
    return t if dtype is None else t.to(dtype)


def _binary_args(a, b):
    """TF converts a python scalar operand to the dtype of the tensor operand."""
    if isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
        return a, torch.as_tensor(b, dtype=a.dtype)
    if isinstance(b, torch.Tensor) and not isinstance(a, torch.Tensor):
        return torch.as_tensor(a, dtype=b.dtype), b
    return convert_to_tensor(a), convert_to_tensor(b)


def cast(x, dtype):
    return convert_to_tensor(x).to(dtype)


def to_int32(x):
    return convert_to_tensor(x).to(int32)          # truncation toward zero


def to_float(x):
    return convert_to_tensor(x).to(float32)


def shape(x):
    return torch.tensor(list(convert_to_tensor(x).shape), dtype=int32)


def rank(x):
    return convert_to_tensor(x).dim()


def reshape(x, shp):
    if isinstance(shp, torch.Tensor):
        shp = [int(i) for i in shp.reshape(-1).tolist()]
    return convert_to_tensor(x).reshape(tuple(int(i) for i in shp))


def transpose(x, perm=None):
    x = convert_to_tensor(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return x.permute(*[int(p) for p in perm])


def squeeze(x, axis=None):
    x = convert_to_tensor(x)
    return x.squeeze() if axis is None else x.squeeze(axis)


def concat(values, axis):
    return torch.cat([convert_to_tensor(v) for v in values], dim=int(axis))


def split(value, num, axis=0):
    return list(torch.chunk(convert_to_tensor(value), int(num), dim=int(axis)))


def zeros(shp, dtype=float32, name=None):
    return torch.zeros(tuple(int(i) for i in shp), dtype=dtype)


def ones(shp, dtype=float32, name=None):
    return torch.ones(tuple(int(i) for i in shp), dtype=dtype)


def tile(x, multiples):
    return convert_to_tensor(x).repeat(*[int(m) for m in multiples])


newaxis = None


def Variable(initial_value, trainable=True, dtype=None, name=None):   # noqa: N802  (tf.Variable: eager, its value)
    return convert_to_tensor(initial_value, dtype)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
    return convert_to_tensor(initializer, dtype)


def zeros_like(x):
    return torch.zeros_like(convert_to_tensor(x))


def range(*args, **kw):   # noqa: A001  (tf.range)
    dtype = kw.get("dtype", int32)
    vals = [int(a) if not isinstance(a, torch.Tensor) else int(a.item()) for a in args]
    return torch.arange(*vals).to(dtype)


def reduce_sum(x, axis=None, keep_dims=False):
    x = convert_to_tensor(x)
    if axis is None:
        return x.sum()
    return x.sum(dim=axis, keepdim=keep_dims)


def reduce_mean(x, axis=None, keep_dims=False):
    x = convert_to_tensor(x)
    if axis is None:
        return x.mean()
    return x.mean(dim=axis, keepdim=keep_dims)


def log(x):
    return torch.log(convert_to_tensor(x))


def sqrt(x):
    return torch.sqrt(convert_to_tensor(x))


def pow(x, y):   # noqa: A001
    x, y = _binary_args(x, y)
    return torch.pow(x, y)


def maximum(x, y):
    x, y = _binary_args(x, y)
    return torch.maximum(x, y)


def greater(x, y):
    x, y = _binary_args(x, y)
    return x > y


def equal(x, y):
    x, y = _binary_args(x, y)
    return x == y


def logical_not(x):
    return ~convert_to_tensor(x)


def clip_by_value(x, lo, hi):
    # TF 1.1 clip_ops.clip_by_value: t_min = minimum(t, clip_value_max); return maximum(t_min, clip_value_min)
    # (the order matters for NumStepsDistribution.log_prob, which passes max = prob itself: prob 0 -> 1e-32)
    x = convert_to_tensor(x)
    return torch.maximum(torch.minimum(x, torch.as_tensor(hi, dtype=x.dtype)), torch.as_tensor(lo, dtype=x.dtype))


def stop_gradient(x):
    return convert_to_tensor(x).detach()


def where(cond, x=None, y=None):
    if x is None:
        return torch.nonzero(cond)                     # int64 coordinates of True entries, row-major
    return torch.where(cond, x, y)


def boolean_mask(tensor, mask):
    return convert_to_tensor(tensor)[mask]


def scatter_nd(indices, updates, shp):
    shp = [int(i) for i in (shp.tolist() if isinstance(shp, torch.Tensor) else shp)]
    out = torch.zeros(shp, dtype=updates.dtype)
    idx = tuple(indices[:, d].long() for d in np.arange(indices.shape[1]))
    out = out.index_put(idx, updates, accumulate=True)
    return out


def gather(params, indices):
    return convert_to_tensor(params)[convert_to_tensor(indices).long()]


def scan(fn, elems):
    acc, outs = elems[0], [elems[0]]
    for i in np.arange(1, elems.shape[0]):
        acc = fn(acc, elems[int(i)])
        outs.append(acc)
    return torch.stack(outs, 0)


def cumsum(x, axis=0, reverse=False):
    x = convert_to_tensor(x)
    if reverse:
        return torch.flip(torch.cumsum(torch.flip(x, [axis]), axis), [axis])
    return torch.cumsum(x, axis)


def cumprod(x, axis=0):
    return torch.cumprod(convert_to_tensor(x), int(axis))


@contextlib.contextmanager
def variable_scope(*a, **k):
    yield


control_dependencies = variable_scope


def group(*a, **k):
    return None


def square(x):
    x = convert_to_tensor(x)
    return x * x


def trainable_variables():
    """Every variable the Sonnet stand-in created, in creation order (nothing else in the reference is trainable)."""
    import snt_stub
    return list(snt_stub.VARIABLES.values())


name_scope = variable_scope


def get_collection(key=None, scope=None):
    """Only tf.get_collection(TRAINABLE_VARIABLES, scope=<module scope>) returns anything (model.py:228-229)."""
    if key == GraphKeys.TRAINABLE_VARIABLES and scope:
        import snt_stub
        return [v for path, v in snt_stub.VARIABLES.items() if path == scope or path.startswith(scope + "/")]
    return []


def add_to_collection(*a, **k):
    return None


class GraphKeys:
    UPDATE_OPS = "update_ops"
    TRAINABLE_VARIABLES = "trainable_variables"


class _Summary:
    @staticmethod
    def scalar(*a, **k):
        return None

    @staticmethod
    def histogram(*a, **k):
        return None


summary = _Summary()


class _NN:
    @staticmethod
    def moments(x, axes):
        x = convert_to_tensor(x)
        axes = list(axes)
        if not axes:
            return x, torch.zeros_like(x)
        mean = x.mean(dim=axes)
        var = ((x - x.mean(dim=axes, keepdim=True)) ** 2).mean(dim=axes)
        return mean, var

    @staticmethod
    def dynamic_rnn(cell, inputs, initial_state=None, time_major=False, **_):
        """Unrolls `cell` over the leading (time) axis and stacks every output over time  [upstream, time_major=True]."""
        assert time_major
        state, per_step = initial_state, []
        for t in builtins_range(int(inputs.shape[0])):
            out, state = cell(inputs[t], state)
            per_step.append(out)
        outputs = [torch.stack([o[i] for o in per_step], 0) for i in builtins_range(len(per_step[0]))]
        return outputs, state

    @staticmethod
    def l2_loss(t):
        return 0.5 * (convert_to_tensor(t) ** 2).sum()

    elu = staticmethod(torch.nn.functional.elu)
    sigmoid = staticmethod(torch.sigmoid)
    tanh = staticmethod(torch.tanh)
    relu = staticmethod(torch.relu)


nn = _NN()


class _Train:
    @staticmethod
    def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
        """decayed = learning_rate * decay_rate ^ (global_step / decay_steps)  [upstream]"""
        lr = convert_to_tensor(learning_rate)
        p = convert_to_tensor(global_step).to(lr.dtype) / convert_to_tensor(decay_steps).to(lr.dtype)
        return lr * torch.pow(convert_to_tensor(decay_rate).to(lr.dtype), p)

    class RMSPropOptimizer:   # only referenced as a default argument value
        pass

    @staticmethod
    def get_or_create_global_step():
        return torch.tensor(0, dtype=int64)


train = _Train()


def truncated_normal_initializer(*a, **k):
    return None


def zeros_initializer(*a, **k):
    return None


def constant_initializer(*a, **k):
    return None


def uniform_unit_scaling_initializer(*a, **k):
    return None


# ------------------------------------------------------------------------------------------------------------------
# tf.contrib.distributions  [upstream TF 1.1]
# ------------------------------------------------------------------------------------------------------------------
NOISE_NORMAL, NOISE_UNIFORM = [], []      # queued draws (tools/make_golden.py)


class Normal:
    def __init__(self, loc, scale, validate_args=False, allow_nan_stats=True, name=None):
        loc, scale = _binary_args(loc, scale)
        self.loc, self.scale = loc, scale
        self.dtype = scale.dtype

    def log_prob(self, x):
        z = (convert_to_tensor(x) - self.loc) / self.scale
        return -0.5 * z * z - (0.5 * math.log(2.0 * math.pi) + torch.log(self.scale))

    def sample(self):
        """sampled * scale + loc with sampled ~ N(0, 1)  [upstream Normal._sample_n]; the draws come from NOISE_NORMAL
        (first in, first out) when the generator queued any, so that the vectors can store them."""
        eps = NOISE_NORMAL.pop(0) if NOISE_NORMAL else torch.randn(self.loc.shape)
        assert tuple(eps.shape) == tuple(self.loc.shape), (tuple(eps.shape), tuple(self.loc.shape))
        return eps * self.scale + self.loc


def kl(a, b):
    """_kl_normal_normal"""
    s_a2, s_b2 = a.scale * a.scale, b.scale * b.scale
    ratio = s_a2 / s_b2
    return (a.loc - b.loc) ** 2 / (2.0 * s_b2) + 0.5 * (ratio - 1.0 - torch.log(ratio))


class Geometric:
    def __init__(self, probs):
        self.probs = convert_to_tensor(probs)
        self.dtype = self.probs.dtype

    def prob(self, counts):
        counts = convert_to_tensor(counts).to(self.dtype)
        return torch.exp(counts * torch.log1p(-self.probs) + torch.log(self.probs))


class Bernoulli:
    def __init__(self, probs=None, dtype=int32, **k):
        self.probs = convert_to_tensor(probs)
        self.dtype = dtype

    def sample(self, n=None):
        """cast(uniform < probs)  [upstream Bernoulli._sample_n]; uniforms from NOISE_UNIFORM when queued."""
        shp = tuple(self.probs.shape) if n is None else (int(n),) + tuple(self.probs.shape)
        u = NOISE_UNIFORM.pop(0) if NOISE_UNIFORM else torch.rand(shp)
        assert tuple(u.shape) == shp, (tuple(u.shape), shp)
        return (u < self.probs).to(self.dtype)


class NormalWithSoftplusScale(Normal):
    def __init__(self, loc, scale, **k):
        super().__init__(loc, torch.nn.functional.softplus(convert_to_tensor(scale)))


def install():
    """Register this module (and the sub-module paths the reference imports) under the TensorFlow / Sonnet names."""
    me = sys.modules[__name__]
    sys.modules["tensorflow"] = me
    contrib = types.ModuleType("tensorflow.contrib")
    dists = types.ModuleType("tensorflow.contrib.distributions")
    for n in ("Normal", "Geometric", "Bernoulli", "NormalWithSoftplusScale"):
        setattr(dists, n, getattr(me, n))
    contrib.distributions = dists
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.xavier_initializer = layers.variance_scaling_initializer = lambda *a, **k: None
    contrib.layers = layers
    me.contrib = contrib
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.distributions"] = dists
    path = "tensorflow.contrib.distributions"
    for part in ("python", "ops", "kullback_leibler"):
        path += "." + part
        sys.modules[path] = types.ModuleType(path)
    sys.modules[path].kl = kl
    for path in ("tensorflow.python", "tensorflow.python.training", "tensorflow.python.training.moving_averages",
                 "tensorflow.python.util", "tensorflow.python.util.nest"):
        sys.modules[path] = types.ModuleType(path)
    sys.modules["tensorflow.python.training"].moving_averages = sys.modules["tensorflow.python.training.moving_averages"]
    sys.modules["tensorflow.python.util"].nest = sys.modules["tensorflow.python.util.nest"]
    sys.modules["tensorflow.python.util.nest"].flatten = lambda x: list(x) if isinstance(x, (list, tuple)) else [x]
    import snt_stub                                       # the Sonnet modules cell.py / modules.py / neural.py use
    sys.modules["sonnet"] = snt_stub.module()
    ev = types.ModuleType("evaluation")                  # evaluation.py is Python-2 syntax; model.py only imports a name
    ev.gradient_summaries = lambda *a, **k: None
    sys.modules["evaluation"] = ev
