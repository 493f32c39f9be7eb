"""CPU oracle for the AIR hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under
``attend_infer_repeat_b200/`` imports it; the product path is CUDA-only.

What it is: a plain torch-CPU (float32, with float64 islands exactly where the
reference has them) restatement of the reference's algorithm for the path named
by BASELINE.json ``north_star``:

    cell.py:101-171       AIRCell.initial_state / _build         -> initial_state, cell_step
    modules.py:11-24      ParametrisedGaussian                   -> what_head
    modules.py:35-63      (Stochastic)TransformParam             -> where_head
    modules.py:66-91      Encoder / Decoder                      -> mlp
    modules.py:94-109     SpatialTransformer (+ Sonnet warper)   -> stn_read, stn_paint
    modules.py:112-122    StepsPredictor                         -> steps_head
    modules.py:125-143    BaselineMLP                            -> baseline_mlp
    neural.py:42-102      Affine / MLP                           -> mlp
    model.py:66-104       AIRModel._build                        -> unroll, postprocess
    model.py:106-124      _anneal_weight                         -> anneal_weight
    model.py:126-216      _prior_loss                            -> prior_loss
    model.py:218-251      _reinforce                             -> reinforce
    model.py:319-343      train_step (loss assembly)             -> elbo
    prior.py:26-32        geometric_prior                        -> geometric_prior
    prior.py:62-68        bernoulli_to_modified_geometric        -> bernoulli_to_modified_geometric
    prior.py:71-90        tabular_kl                             -> tabular_kl
    prior.py:103-151      sample_from_tensor/NumStepsDistribution-> num_steps_prob / num_steps_log_prob
    ops.py:5-43,67-76     Loss / clip_preserve                   -> Loss, clip_preserve

The arithmetic of the reference lives in third-party packages that are NOT under
/root/reference (TensorFlow 1.1.0rc1 and Sonnet v1.1 @ 3fd7d9d, README.md:14):
snt.LSTM, snt.Linear, snt.AffineGridWarper(+.inverse()), snt.resampler,
tf.contrib.distributions.{Normal, NormalWithSoftplusScale, Bernoulli, Geometric, kl}.
Their published algorithms are restated here (see each function).

PARITY PINNING STATUS
  * pinned by the reference's own tests (test/prior_test.py:15-24, 40-44, 86-120,
    160-205): geometric_prior, tabular_kl, bernoulli_to_modified_geometric and the
    KL(posterior || prior) stress properties  ->  tests/test_oracle_prior.py.
  * pinned by running the reference's own prior.py / ops.py / model.py source in
    this container over a torch-backed stand-in for the TF primitives
    (tools/make_golden.py -> tests/golden/*.npz): step-count algebra, _anneal_weight,
    _prior_loss, _reinforce.
  * pinned by running the reference's own cell.py / modules.py / neural.py / model.py /
    mnist_model.py source over tools/snt_stub.py + tools/tf_stub.py (make_golden.py:
    cell_vectors -> tests/golden/reference_cell_*.npz): AIRCell._build x T through
    dynamic_rnn, the post-processing of model.py:83-104 and the reconstruction loss --
    i.e. every decision those files make (parameter order (sx, tx, sy, ty), biases,
    explore-eps mix, presence product, LSTM wiring, canvas accumulation, output
    multiplier, variable inventory): script configuration through AIRonMNIST, a
    non-square canvas with odd widths, and the non-discrete mode; agreement <= 8e-6.
  * pinned likewise: AIRModel.train_step (model.py:261-376) run from the reference's
    source with compute_gradients = autograd through its own loss assembly -- every loss
    term and d opt_loss / d (every model variable), plus BaselineMLP output / loss /
    gradient through AIRonMNIST; autograd on this oracle agrees to <= 3.4e-5 of max |g|.
  * PARITY UNPINNED by the reference: the arithmetic INSIDE the Sonnet / TF modules
    (snt.Linear, snt.LSTM, AffineGridWarper, resampler, the distributions) -- those
    packages are absent, test/cell_test.py asserts nothing and there are no golden
    files; the stand-ins restate their documented semantics, with the warper and the
    resampler implemented through torch's affine_grid / grid_sample, independently of
    this file's hand-written gather.  Further cross-checks in
    tests/test_oracle_blocks.py: torch.nn.LSTMCell with gate permutation,
    torch.distributions Normal / kl_divergence.

Everything is written with differentiable torch ops so that autograd on this oracle is
the gradient oracle for the backward kernels (tf.gradients on the reference graph).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

F32 = torch.float32
F64 = torch.float64


# --------------------------------------------------------------------------------------
# configuration + canonical parameter layout (shared verbatim with the C-ABI library)
# --------------------------------------------------------------------------------------
@dataclass
class AirConfig:
    """Hyper-parameters of the path; defaults = scripts/multi_mnist.py:24-94 + mnist_model.py:13-44."""
    H: int = 50
    W: int = 50
    h: int = 20
    w: int = 20
    T: int = 3
    na: int = 50                      # n_appearance, mnist_model.py:34
    nh: int = 256                     # snt.LSTM(256), mnist_model.py:35
    enc_hidden: Sequence[int] = (256, 256)      # inpt_encoder_hidden, multi_mnist.py:85
    glenc_hidden: Sequence[int] = (256, 256)    # glimpse_encoder_hidden
    dec_hidden: Sequence[int] = (256, 256)      # glimpse_decoder_hidden
    where_hidden: Sequence[int] = (256, 256)    # transform_estimator_hidden
    steps_hidden: Sequence[int] = (128, 64)     # steps_pred_hidden, multi_mnist.py:89
    output_std: float = 0.3           # mnist_model.py:42
    output_multiplier: float = 0.5    # multi_mnist.py:57
    explore_eps: Optional[float] = 1e-3   # multi_mnist.py:59
    scale_bias: float = 0.5           # transform_var_bias, multi_mnist.py:56
    step_bias: float = 0.75           # multi_mnist.py:55
    what_scale_offset: float = 0.5    # cell.py:66
    forget_bias: float = 1.0          # snt.LSTM default
    max_crop_size: float = 1.0        # modules.py:29
    discrete_steps: bool = True

    @property
    def P(self) -> int:
        return self.H * self.W

    @property
    def G(self) -> int:
        return self.h * self.w


def param_spec(cfg: AirConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """Canonical flat parameter order.  Weights are [in, out] row-major like snt.Linear."""
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def mlp(prefix, n_in, hidden, n_out=None):
        d = n_in
        for i, n in enumerate(hidden):
            spec.append((f"{prefix}.{i}.w", (d, n)))
            spec.append((f"{prefix}.{i}.b", (n,)))
            d = n
        if n_out is not None:
            spec.append((f"{prefix}.out.w", (d, n_out)))
            spec.append((f"{prefix}.out.b", (n_out,)))
        return d

    n_enc = mlp("input_encoder", cfg.P, cfg.enc_hidden)
    spec.append(("lstm.w", (n_enc + cfg.nh, 4 * cfg.nh)))
    spec.append(("lstm.b", (4 * cfg.nh,)))
    spec.append(("lstm.h0", (cfg.nh,)))
    spec.append(("lstm.c0", (cfg.nh,)))
    mlp("transform_estimator", cfg.nh, cfg.where_hidden, 8)
    mlp("steps_predictor", cfg.nh, cfg.steps_hidden, 1)
    n_gl = mlp("glimpse_encoder", cfg.G, cfg.glenc_hidden)
    spec.append(("what.w", (n_gl, 2 * cfg.na)))
    spec.append(("what.b", (2 * cfg.na,)))
    mlp("glimpse_decoder", cfg.na, cfg.dec_hidden, cfg.G)
    return spec


def param_count(cfg: AirConfig) -> int:
    return sum(int(np.prod(s)) for _, s in param_spec(cfg))


def init_params(cfg: AirConfig, seed: int = 0, dtype=F32) -> Dict[str, torch.Tensor]:
    """Effective reference initialiser: truncated normal, sigma = 1/sqrt(fan_in),
    +-2 sigma, zero biases, zero trainable LSTM initial state."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_spec(cfg):
        if name.endswith(".w"):
            std = 1.0 / math.sqrt(shape[0])
            t = torch.empty(shape, dtype=F32)
            torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=g)
        else:
            t = torch.zeros(shape, dtype=F32)
        out[name] = t.to(dtype)
    return out


def flatten_params(cfg: AirConfig, params: Dict[str, torch.Tensor]) -> torch.Tensor:
    return torch.cat([params[n].reshape(-1) for n, _ in param_spec(cfg)])


def unflatten_params(cfg: AirConfig, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for name, shape in param_spec(cfg):
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].view(shape)
        off += n
    assert off == flat.numel()
    return out


def make_noise(cfg: AirConfig, B: int, seed: int = 0, dtype=F32):
    """Pre-drawn noise shared by oracle and kernel: eps_where[T,B,4], eps_what[T,B,na], u_pres[T,B,1]."""
    g = torch.Generator().manual_seed(1000 + seed)
    eps_where = torch.randn(cfg.T, B, 4, generator=g, dtype=F32).to(dtype)
    eps_what = torch.randn(cfg.T, B, cfg.na, generator=g, dtype=F32).to(dtype)
    u_pres = torch.rand(cfg.T, B, 1, generator=g, dtype=F32).to(dtype)
    return eps_where, eps_what, u_pres


def synthetic_multi_mnist(B: int, H: int = 50, W: int = 50, seed: int = 0, max_objects: int = 2):
    """Multi-MNIST-shaped synthetic canvases (data/data.py:35-107,116): 0..max_objects stroke-like blobs in tight
    boxes (~20x20 at 50x50), no overlap, background exactly 0, foreground uint8/255.  Returns (imgs[B,H,W] f32,
    nums[max_objects+1,B,1] f32 with the first n entries 1 (data.py:101-105))."""
    rng = np.random.default_rng(seed)
    imgs = np.zeros((B, H, W), dtype=np.uint8)
    nums = np.zeros((max_objects + 1, B, 1), dtype=np.float32)
    s = max(4, int(round(20 * H / 50)))
    yy, xx = np.mgrid[0:s, 0:s].astype(np.float32)
    for b in range(B):
        n = int(rng.integers(0, max_objects + 1))
        placed = []
        tries = 0
        while len(placed) < n and tries < 50:
            tries += 1
            y0 = int(rng.integers(0, H - s + 1))
            x0 = int(rng.integers(0, W - s + 1))
            if any(abs(y0 - py) < s and abs(x0 - px) < s for py, px in placed):
                continue
            placed.append((y0, x0))
            # a random thick poly-line ("stroke") inside the box
            pts = rng.uniform(0.15 * s, 0.85 * s, size=(4, 2)).astype(np.float32)
            d = np.full((s, s), 1e9, dtype=np.float32)
            for k in range(3):
                p, q = pts[k], pts[k + 1]
                v = q - p
                tt = np.clip(((xx - p[0]) * v[0] + (yy - p[1]) * v[1]) / max(float(v @ v), 1e-6), 0, 1)
                d = np.minimum(d, np.hypot(xx - (p[0] + tt * v[0]), yy - (p[1] + tt * v[1])))
            blob = np.clip(1.6 - d / (0.06 * s), 0, 1)
            imgs[b, y0:y0 + s, x0:x0 + s] = np.maximum(imgs[b, y0:y0 + s, x0:x0 + s],
                                                       (blob * 255).astype(np.uint8))
        nums[:len(placed), b, 0] = 1.0
    return torch.from_numpy(imgs.astype(np.float32) / 255.0), torch.from_numpy(nums)


# --------------------------------------------------------------------------------------
# elementwise primitives, restating the TF 1.1 CPU kernels [upstream]
# --------------------------------------------------------------------------------------
def elu(x):
    """tf.nn.elu [upstream Eigen functor]: x if x > 0 else exp(x) - 1."""
    return torch.where(x > 0, x, torch.exp(torch.clamp(x, max=0.0)) - 1.0)


def softplus(x):
    """tf.nn.softplus [upstream functor]: threshold = log(eps)+2; x if x > -thr; exp(x) if x < thr; else log(exp(x)+1)."""
    thr = math.log(torch.finfo(x.dtype).eps) + 2.0
    ex = torch.exp(torch.clamp(x, max=-thr))
    mid = torch.log(ex + 1.0)
    return torch.where(x > -thr, x, torch.where(x < thr, ex, mid))


def sigmoid(x):
    return torch.sigmoid(x)


def mlp(x, params, prefix, n_hidden, out_linear):
    """neural.py:63-102: Affine(ELU) hidden layers, optional linear output layer (transfer=None)."""
    for i in range(n_hidden):
        x = elu(x @ params[f"{prefix}.{i}.w"] + params[f"{prefix}.{i}.b"])
    if out_linear:
        x = x @ params[f"{prefix}.out.w"] + params[f"{prefix}.out.b"]
    return x


def lstm_step(x, h, c, w, b, forget_bias=1.0):
    """snt.LSTM [upstream, Sonnet v1.1, no peepholes / batch-norm]: gates = [x,h].W + b; i,j,f,o = split."""
    gates = torch.cat([x, h], 1) @ w + b
    i, j, f, o = torch.split(gates, gates.shape[1] // 4, dim=1)
    c_new = sigmoid(f + forget_bias) * c + sigmoid(i) * torch.tanh(j)
    h_new = torch.tanh(c_new) * sigmoid(o)
    return h_new, c_new


# --------------------------------------------------------------------------------------
# spatial transformer: snt.AffineGridWarper (+ .inverse()) + snt.resampler  [upstream]
# --------------------------------------------------------------------------------------
def _linspace(n, dtype):
    # Sonnet builds the grid with np.linspace(-1, 1, n, dtype=float32) (float64 maths, cast at the end)
    return torch.from_numpy(np.linspace(-1.0, 1.0, n, dtype=np.float32)).to(dtype)


def warp_grid_forward(where, src_hw, out_hw):
    """AffineGridWarper(source=img H x W, output=h x w, no_shear_2d) (cell.py:56-58, modules.py:100,108).
    where = (sx, tx, sy, ty) (modules.py:42).  x = sx*(u*S_W) + tx*S_W + S_W, y likewise.  Returns x,y [B,h,w]."""
    Hs, Ws = src_hw
    h, w = out_hw
    dt = where.dtype
    S_W = torch.tensor((Ws - 1.0) * 0.5, dtype=dt)
    S_H = torch.tensor((Hs - 1.0) * 0.5, dtype=dt)
    uS = _linspace(w, dt) * S_W            # feature rows are pre-scaled in numpy float32
    vS = _linspace(h, dt) * S_H
    sx, tx, sy, ty = (where[:, k:k + 1] for k in range(4))
    x = (sx * uS[None, :] + tx * S_W) + S_W     # [B,w]  (K=2 matmul, then the constant offset)
    y = (sy * vS[None, :] + ty * S_H) + S_H     # [B,h]
    B = where.shape[0]
    return x[:, None, :].expand(B, h, w), y[:, :, None].expand(B, h, w)


def warp_grid_inverse(where, glimpse_hw, canvas_hw):
    """AffineGridWarper.inverse() (modules.py:101-102) [upstream]: a,b,c,d = sx,0,0,sy; det = a*d-b*c;
    a' = d/det, d' = a/det; (tx',ty') = M^-1 (tx,ty); new unconstrained warper source=glimpse, output=canvas with
    params (a', b', -tx', c', d', -ty').  Returns glimpse-space pixel coords x,y [B,H,W]."""
    h, w = glimpse_hw
    H, W = canvas_hw
    dt = where.dtype
    S_w = torch.tensor((w - 1.0) * 0.5, dtype=dt)
    S_h = torch.tensor((h - 1.0) * 0.5, dtype=dt)
    US = _linspace(W, dt) * S_w
    VS = _linspace(H, dt) * S_h
    sx, tx, sy, ty = (where[:, k:k + 1] for k in range(4))
    det = sx * sy                              # a*d - b*c with b = c = 0
    a_inv = sy / det
    d_inv = sx / det
    tx_inv = a_inv * tx                        # + b_inv * ty, b_inv = -0/det
    ty_inv = d_inv * ty
    x = (a_inv * US[None, :] + (-tx_inv) * S_w) + S_w     # [B,W]
    y = (d_inv * VS[None, :] + (-ty_inv) * S_h) + S_h     # [B,H]
    B = where.shape[0]
    return x[:, None, :].expand(B, H, W), y[:, :, None].expand(B, H, W)


def resample(data, x, y):
    """snt.resampler [upstream C++ kernel]: bilinear sample of data[B,Hs,Ws] at pixel coords (x,y) [B,...];
    zero outside (-1,Ws)x(-1,Hs); taps outside [0,Ws-1]x[0,Hs-1] read 0.  floor() is a constant for autograd, which
    reproduces the op's registered gradient (wrt data: scatter of the same weights; wrt warp: data differences)."""
    B, Hs, Ws = data.shape
    shp = x.shape
    x = x.reshape(B, -1)
    y = y.reshape(B, -1)
    inside = (x > -1.0) & (y > -1.0) & (x < float(Ws)) & (y < float(Hs))
    xs = torch.where(inside, x, torch.zeros_like(x))
    ys = torch.where(inside, y, torch.zeros_like(y))
    fx = torch.floor(xs.detach())
    fy = torch.floor(ys.detach())
    cx = fx + 1.0
    cy = fy + 1.0
    dx = cx - xs
    dy = cy - ys
    flat = data.reshape(B, Hs * Ws)

    def tap(ix, iy):
        ok = (ix >= 0) & (ix <= Ws - 1) & (iy >= 0) & (iy <= Hs - 1)
        idx = (iy.clamp(0, Hs - 1) * Ws + ix.clamp(0, Ws - 1)).long()
        v = torch.gather(flat, 1, idx)
        return torch.where(ok, v, torch.zeros_like(v))

    one = 1.0
    out = (dx * dy * tap(fx, fy) + (one - dx) * (one - dy) * tap(cx, cy)
           + dx * (one - dy) * tap(fx, cy) + (one - dx) * dy * tap(cx, fy))
    out = torch.where(inside, out, torch.zeros_like(out))
    return out.reshape(shp)


def stn_read(img, where, glimpse_hw):
    """SpatialTransformer forward (cell.py:135): crop[B,h,w] from img[B,H,W]."""
    x, y = warp_grid_forward(where, img.shape[1:], glimpse_hw)
    return resample(img, x, y)


def stn_paint(glimpse, where, canvas_hw):
    """inverse SpatialTransformer (cell.py:159): gather of the glimpse at every canvas pixel -> [B,H,W]."""
    x, y = warp_grid_inverse(where, glimpse.shape[1:], canvas_hw)
    return resample(glimpse, x, y)


# --------------------------------------------------------------------------------------
# AIRCell
# --------------------------------------------------------------------------------------
def initial_state(cfg: AirConfig, params, img):
    """cell.py:101-114: [flat_img, flat_canvas(0), what(0), where(0), (h0,c0) tiled, presence(1)]."""
    B = img.shape[0]
    dt = img.dtype
    return dict(
        img=img.reshape(B, cfg.P),
        canvas=torch.zeros(B, cfg.P, dtype=dt),
        what=torch.zeros(B, cfg.na, dtype=dt),
        where=torch.zeros(B, 4, dtype=dt),
        h=params["lstm.h0"][None, :].expand(B, cfg.nh).to(dt),
        c=params["lstm.c0"][None, :].expand(B, cfg.nh).to(dt),
        presence=torch.ones(B, 1, dtype=dt),
    )


def where_head(cfg, params, hidden):
    """modules.py:58-63, 35-46: MLP -> 8; loc = (sig, tanh, sig, tanh) * (max_crop, 1, max_crop, 1); raw + scale_bias."""
    m = mlp(hidden, params, "transform_estimator", len(cfg.where_hidden), True)
    sx = cfg.max_crop_size * sigmoid(m[:, 0:1])
    tx = torch.tanh(m[:, 1:2])
    sy = cfg.max_crop_size * sigmoid(m[:, 2:3])
    ty = torch.tanh(m[:, 3:4])
    loc = torch.cat([sx, tx, sy, ty], -1)
    raw = m[:, 4:8] + cfg.scale_bias
    return loc, raw


def steps_head(cfg, params, hidden):
    """modules.py:119-122 + cell.py:140-141."""
    logit = mlp(hidden, params, "steps_predictor", len(cfg.steps_hidden), True) + cfg.step_bias
    p = sigmoid(logit)
    if cfg.explore_eps is not None:
        p = cfg.explore_eps / 2 + (1 - cfg.explore_eps) * p
    return p


def cell_step(cfg: AirConfig, params, state, eps_where, eps_what, u_pres):
    """cell.py:116-171, one step for all B samples.  Returns (outputs dict in output_names order, new state)."""
    B = state["img"].shape[0]
    img = state["img"].reshape(B, cfg.H, cfg.W)
    enc = mlp(state["img"], params, "input_encoder", len(cfg.enc_hidden), False)            # cell.py:125
    h, c = lstm_step(enc, state["h"], state["c"], params["lstm.w"], params["lstm.b"], cfg.forget_bias)  # :126-127
    where_loc, raw = where_head(cfg, params, h)                                                # :129
    where_scale = softplus(raw)                                                                # :130-132
    where = eps_where * where_scale + where_loc                                                # :133
    crop = stn_read(img, where, (cfg.h, cfg.w))                                                # :135
    presence_prob = steps_head(cfg, params, h)                                                 # :137-141
    if cfg.discrete_steps:
        z = (u_pres < presence_prob.detach()).to(presence_prob.dtype)                          # :143-148
        presence = state["presence"] * z
    else:
        presence = presence_prob                                                               # :150-151
    q = mlp(crop.reshape(B, cfg.G), params, "glimpse_encoder", len(cfg.glenc_hidden), False)  # :153
    r = q @ params["what.w"] + params["what.b"]                                                # modules.py:20-21
    what_loc = r[:, :cfg.na]
    what_scale = softplus(r[:, cfg.na:] + cfg.what_scale_offset)                               # modules.py:23
    what = eps_what * what_scale + what_loc                                                    # :156
    dec = mlp(what, params, "glimpse_decoder", len(cfg.dec_hidden), True)                     # :158
    inv = stn_paint(dec.reshape(B, cfg.h, cfg.w), where, (cfg.H, cfg.W))                      # :159
    canvas = state["canvas"] + presence * inv.reshape(B, cfg.P)                                # :161-164
    outputs = dict(canvas=canvas, glimpse=dec, what=what, what_loc=what_loc, what_scale=what_scale,
                   where=where, where_loc=where_loc, where_scale=where_scale,
                   presence_prob=presence_prob, presence=presence)
    new_state = dict(img=state["img"], canvas=canvas, what=what, where=where, h=h, c=c, presence=presence)
    return outputs, new_state


OUTPUT_NAMES = "canvas glimpse what what_loc what_scale where where_loc where_scale presence_prob presence".split()


def unroll(cfg: AirConfig, params, img, eps_where, eps_what, u_pres):
    """model.py:81-87: T applications of the cell; every output stacked time-major [T,B,.]."""
    state = initial_state(cfg, params, img)
    outs = {k: [] for k in OUTPUT_NAMES}
    for t in range(cfg.T):
        o, state = cell_step(cfg, params, state, eps_where[t], eps_what[t], u_pres[t])
        for k in OUTPUT_NAMES:
            outs[k].append(o[k])
    outs = {k: torch.stack(v, 0) for k, v in outs.items()}
    return outs, state


# --------------------------------------------------------------------------------------
# prior.py / ops.py
# --------------------------------------------------------------------------------------
def clip_preserve(expr, lo, hi):
    """ops.py:67-76: clip in the forward pass, identity in the backward pass."""
    clipped = torch.maximum(torch.minimum(expr, torch.as_tensor(hi, dtype=expr.dtype)),
                            torch.as_tensor(lo, dtype=expr.dtype))
    return (clipped - expr).detach() + expr


def geometric_prior(success_prob, n_steps, dtype=None):
    """prior.py:26-32 [upstream Geometric(probs=1-s).prob(k) = exp(k*log1p(-(1-s)) + log(1-s))].  dtype follows the
    input: python float -> float32 (test/prior_test.py:15-24), float64 when it comes out of anneal_weight."""
    if not torch.is_tensor(success_prob):
        success_prob = torch.tensor(success_prob, dtype=dtype or F32)
    s = torch.clamp(success_prob, 1e-7, 1.0 - 1e-15)
    probs = 1.0 - s
    k = torch.arange(n_steps + 1, dtype=s.dtype)
    return torch.exp(k * torch.log1p(-probs) + torch.log(probs))


def bernoulli_to_modified_geometric(presence_prob):
    """prior.py:62-68 (float64 island) with the scan-cumprod of prior.py:35-59."""
    p = presence_prob.to(F64)
    inv = 1.0 - p
    prob = torch.cumprod(p, dim=-1)
    mod = torch.cat([inv[..., :1], inv[..., 1:] * prob[..., :-1], prob[..., -1:]], -1)
    mod = mod / mod.sum(-1, keepdim=True)
    return mod.to(F32)


def tabular_kl(p, q, zero_prob_value=0.0):
    """prior.py:71-90 (float64 island; masked_apply prior.py:8-23: exactly 0 where p <= zero_prob_value,
    and no NaN gradient there)."""
    p = p.to(F64)
    q = torch.as_tensor(q).to(F64)
    non_zero = p > zero_prob_value
    logarg = p / q
    safe = torch.where(non_zero, logarg, torch.ones_like(logarg))
    log = torch.where(non_zero, torch.log(safe), torch.zeros_like(safe))
    return (p * log).to(F32)


def num_steps_prob(joint, samples=None):
    """prior.py:141-146, 103-116: joint[b, int(n_b)] by flat gather."""
    if samples is None:
        return joint
    idx = samples.to(torch.int32).long().reshape(-1, 1)
    return torch.gather(joint, 1, idx).reshape(samples.shape)


def num_steps_log_prob(joint, samples):
    """prior.py:148-151."""
    prob = num_steps_prob(joint, samples)
    prob = clip_preserve(prob, 1e-32, prob.detach())
    return torch.log(prob)


class Loss:
    """ops.py:5-43."""

    def __init__(self):
        self._value = None
        self._per_sample = None

    def add(self, loss=None, per_sample=None, weight=1.0):
        if isinstance(loss, Loss):
            per_sample, loss = loss.per_sample, loss.value
        self._value = loss * weight if self._value is None else self._value + loss * weight
        ps = per_sample * weight
        if self._per_sample is not None:
            assert self._per_sample.shape == ps.shape
            ps = self._per_sample + ps
        self._per_sample = ps

    @property
    def value(self):
        return torch.zeros([]) if self._value is None else self._value

    @property
    def per_sample(self):
        return torch.zeros([]) if self._per_sample is None else self._per_sample


# --------------------------------------------------------------------------------------
# AIRModel: post-processing, prior loss, REINFORCE, loss assembly
# --------------------------------------------------------------------------------------
@dataclass
class PriorConfig:
    """scripts/multi_mnist.py:38-51 defaults."""
    what_loc: float = 0.0
    what_scale: float = 1.0
    where_scale_loc: float = 0.0
    where_scale_scale: float = 1.0
    where_shift_loc: Optional[float] = 0.0     # None -> 'loc' not in where_shift_prior -> posterior mean (model.py:202-205)
    where_shift_scale: float = 1.0
    steps_anneal: Optional[str] = "exp"
    steps_init: float = 1.0 - 1e-15
    steps_final: float = 1e-7
    steps_div: float = 1e4
    steps_steps: float = 1e5
    steps_hold_init: float = 1e3
    steps_weight: float = 1.0
    analytic: bool = True
    use_prior: bool = True
    use_reinforce: bool = True


def anneal_weight(init_val, final_val, anneal_type, global_step, anneal_steps, hold_for=0.0, steps_div=1.0):
    """model.py:106-124, float64.  'exp' uses tf.train.exponential_decay(val, step, steps_div, rate) =
    val * rate ** (step / steps_div) [upstream, staircase=False]."""
    # tf.cast(python_float, tf.float64) goes through a float32 constant first (ops.convert_to_tensor) [upstream]:
    # the schedule constants enter the float64 island rounded to float32 (1 - 1e-15 -> 1.0); global_step is an int64
    # variable and is exact.  Confirmed by running the reference's own _anneal_weight (tools/make_golden.py).
    val, final, hold_for, anneal_steps, steps_div = (
        torch.tensor(float(v), dtype=F32).to(F64) for v in (init_val, final_val, hold_for, anneal_steps, steps_div))
    step = torch.tensor(float(global_step), dtype=F64)
    step = torch.clamp(step - hold_for, min=0.0)
    if anneal_type == "exp":
        decay_rate = torch.pow(final / val, steps_div / anneal_steps)
        val = val * torch.pow(decay_rate, step / steps_div)
    elif anneal_type == "linear":
        val = final + (val - final) * (1.0 - step / anneal_steps)
    else:
        raise NotImplementedError
    return torch.maximum(final, val)


def normal_kl(mu_a, s_a, mu_b, s_b):
    """tf.contrib.distributions _kl_normal_normal [upstream]."""
    sa2 = s_a * s_a
    sb2 = torch.as_tensor(s_b, dtype=mu_a.dtype) ** 2
    ratio = sa2 / sb2
    return (mu_a - mu_b) ** 2 / (2.0 * sb2) + 0.5 * (ratio - 1.0 - torch.log(ratio))


def steps_prior_success_prob(pc: PriorConfig, global_step):
    """model.py:133-142."""
    if pc.steps_anneal is not None:
        return anneal_weight(pc.steps_init, pc.steps_final, pc.steps_anneal, global_step, pc.steps_steps,
                             pc.steps_hold_init, pc.steps_div)
    return torch.tensor(pc.steps_init, dtype=F32)


def postprocess(cfg: AirConfig, outs):
    """model.py:89-103."""
    T, B = outs["presence"].shape[:2]
    canvas = outs["canvas"].reshape(T, B, cfg.H, cfg.W) * cfg.output_multiplier
    final_canvas = canvas[-1]
    glimpse_viz = (outs["presence"] * sigmoid(outs["glimpse"])).reshape(T, B, cfg.h, cfg.w)
    step_probs = outs["presence_prob"].reshape(T, B).transpose(0, 1)     # [B,T]
    joint = bernoulli_to_modified_geometric(step_probs)                   # [B,T+1]
    num_step_per_sample = outs["presence"].sum(0).reshape(B)
    return dict(canvas=canvas, final_canvas=final_canvas, glimpse=glimpse_viz, num_steps_posterior=joint,
                num_step_per_sample=num_step_per_sample, num_step=num_step_per_sample.mean())


def prior_loss(cfg: AirConfig, pc: PriorConfig, outs, post, global_step=0):
    """model.py:126-216.  Returns (Loss, dict of the named terms)."""
    T, B = outs["presence"].shape[:2]
    pl = Loss()
    terms = {}
    s = steps_prior_success_prob(pc, global_step)
    terms["steps_prior_success_prob"] = s
    prior = geometric_prior(s, cfg.T)
    posterior = post["num_steps_posterior"]
    steps_kl = tabular_kl(posterior, prior)
    kl_n = steps_kl.sum(1)                                                # [B]
    terms["kl_num_steps_per_sample"] = kl_n
    terms["kl_num_steps"] = kl_n.mean()
    pl.add(terms["kl_num_steps"], kl_n, weight=pc.steps_weight)
    if pc.analytic:
        sw = posterior[..., 1:].transpose(0, 1)                           # [T,B]
        sw = torch.flip(torch.cumsum(torch.flip(sw, [0]), 0), [0])        # reverse cumsum
    else:
        sw = outs["presence"].reshape(T, B)
    terms["prior_step_weight"] = sw
    what_kl = normal_kl(outs["what_loc"], outs["what_scale"], pc.what_loc, pc.what_scale).sum(-1) * sw
    what_ps = what_kl.sum(0)
    terms["kl_what_per_sample"] = what_ps
    terms["kl_what"] = what_ps.mean()
    pl.add(terms["kl_what"], what_ps)
    wl, ws = outs["where_loc"], outs["where_scale"]
    us = torch.stack([wl[..., 0], wl[..., 2]], -1)
    ss = torch.stack([ws[..., 0], ws[..., 2]], -1)
    ut = torch.stack([wl[..., 1], wl[..., 3]], -1)
    st = torch.stack([ws[..., 1], ws[..., 3]], -1)
    scale_kl = normal_kl(us, ss, pc.where_scale_loc, pc.where_scale_scale)
    shift_mean = ut if pc.where_shift_loc is None else pc.where_shift_loc
    shift_kl = normal_kl(ut, st, shift_mean, pc.where_shift_scale)
    where_kl = (scale_kl + shift_kl).sum(-1) * sw
    where_ps = where_kl.sum(0)
    terms["kl_where_per_sample"] = where_ps
    terms["kl_where"] = where_ps.mean()
    pl.add(terms["kl_where"], where_ps)
    return pl, terms


def rec_loss(cfg: AirConfig, obs, final_canvas):
    """model.py:319-322 [upstream Normal.log_prob]: sum_px 0.5*((x-mu)/sigma)^2 + 0.5*log(2 pi) + log(sigma)."""
    z = (obs - final_canvas) / cfg.output_std
    lp = -0.5 * z * z - (0.5 * math.log(2.0 * math.pi) + math.log(cfg.output_std))
    return (-lp).sum((1, 2))


def reinforce(joint, num_step_per_sample, importance_weight, baseline=None, nvil=None):
    """model.py:218-251.  baseline is [B,1] -> importance weight broadcasts to [B,B], reproduced as
    written.  nvil = (imp_weight_moving_mean, imp_weight_moving_var) switches on the decay_rate branch (model.py:232-239):
    iw <- (iw - moving_mean) / max(sqrt(moving_var), 1); the batch moments that feed the moving averages
    (tf.nn.moments over every axis of the squeezed weight, population variance) are returned as well."""
    log_prob = num_steps_log_prob(joint, num_step_per_sample)
    iw = importance_weight
    if baseline is not None:
        iw = iw - baseline
    moments = (iw.detach().mean(), iw.detach().var(unbiased=False))
    if nvil is not None:
        mm, mv = (torch.as_tensor(v, dtype=iw.dtype) for v in nvil)
        iw = (iw - mm) / torch.maximum(torch.sqrt(mv), torch.ones_like(mv))
    rl = (iw.detach() * log_prob).mean()
    return rl, iw, log_prob, moments


def elbo(cfg: AirConfig, pc: PriorConfig, obs, outs, global_step=0, baseline=None, nvil=None):
    """model.py:319-343: loss.value = rec + prior_weight * (kl_n + kl_what + kl_where); ELBO = -loss.value."""
    post = postprocess(cfg, outs)
    res = dict(post)
    loss = Loss()
    rps = rec_loss(cfg, obs, post["final_canvas"])
    res["rec_loss_per_sample"] = rps
    res["rec_loss"] = rps.mean()
    loss.add(res["rec_loss"], rps)
    pl, terms = prior_loss(cfg, pc, outs, post, global_step)
    res.update(terms)
    res["prior_loss"] = pl.value
    res["prior_loss_per_sample"] = pl.per_sample
    loss.add(pl, weight=1.0 if pc.use_prior else 0.0)
    res["loss"] = loss.value
    res["loss_per_sample"] = loss.per_sample
    opt_loss = loss.value
    if pc.use_reinforce:
        iw = rps
        if not pc.analytic:
            iw = iw + pl.per_sample
        rl, iw_full, log_prob, moments = reinforce(post["num_steps_posterior"], post["num_step_per_sample"], iw, baseline,
                                                   nvil)
        res["imp_weight_moments"] = moments
        res["reinforce_loss"] = rl
        res["importance_weight"] = iw_full
        res["num_steps_log_prob"] = log_prob
        opt_loss = opt_loss + rl
    res["opt_loss"] = opt_loss
    res["elbo"] = -loss.value
    return res


def forward(cfg: AirConfig, pc: PriorConfig, params, img, eps_where, eps_what, u_pres, global_step=0, baseline=None,
            nvil=None):
    """Whole hot path: unroll + ELBO.  The encoder is NOT hoisted (as written in the reference)."""
    outs, state = unroll(cfg, params, img, eps_where, eps_what, u_pres)
    res = elbo(cfg, pc, img, outs, global_step, baseline, nvil)
    res["outs"] = outs
    res["final_h"], res["final_c"] = state["h"], state["c"]
    return res


def iwae_bound(cfg: AirConfig, pc: PriorConfig, res, K: int, global_step=0):
    """Importance-weighted bound from a forward() result whose rows are canvases x K particles (row = canvas * K + k).
    NOT in the reference: float64 restatement of log(1/K sum_k w_k),
    log w = log p(x|z) + log p(z,n) - log q(z,n|x), from the same distributions the ELBO uses (Normal priors
    multi_mnist.py:49-51, geometric step prior prior.py:26-32, NumStepsDistribution.log_prob prior.py:148-151);
    latents of steps whose sampled presence is 0 never reach the canvas and are left out.  PARITY UNPINNED."""
    outs = res["outs"]
    T, R = outs["presence"].shape[:2]
    d = lambda t: t.detach().to(F64)
    pres = d(outs["presence"]).reshape(T, R)

    def logpdf(x, mu, s):
        mu, s = torch.as_tensor(mu, dtype=F64), torch.as_tensor(s, dtype=F64)
        return -0.5 * ((x - mu) / s) ** 2 - torch.log(s) - 0.5 * math.log(2.0 * math.pi)

    what, where = d(outs["what"]), d(outs["where"])
    lq = (logpdf(what, d(outs["what_loc"]), d(outs["what_scale"])).sum(-1)
          + logpdf(where, d(outs["where_loc"]), d(outs["where_scale"])).sum(-1))
    lp_where = (logpdf(where[..., 0::2], pc.where_scale_loc, pc.where_scale_scale).sum(-1)
                + logpdf(where[..., 1::2], pc.where_shift_loc, pc.where_shift_scale).sum(-1))
    lp = logpdf(what, pc.what_loc, pc.what_scale).sum(-1) + lp_where
    lq, lp = (lq * pres).sum(0), (lp * pres).sum(0)                      # only steps that were taken
    n = pres.sum(0).long()
    prior_n = geometric_prior(steps_prior_success_prob(pc, global_step), cfg.T).to(F64)
    log_w = -d(res["rec_loss_per_sample"]) + (lp + torch.log(prior_n[n])) - (lq + d(res["num_steps_log_prob"]))
    per_canvas = torch.logsumexp(log_w.reshape(R // K, K), dim=1) - math.log(K)
    return dict(log_w=log_w, bound_per_canvas=per_canvas, bound=per_canvas.mean())


def baseline_mlp(params_b, n_hidden, img, what, where, presence, h, c):
    """modules.py:125-143: concat[img, what (batch-major), where, presence, h, c] -> MLP -> [B,1]."""
    B = img.shape[0]
    parts = [t.transpose(0, 1).reshape(B, -1) for t in (what, where, presence)] + [h, c]
    x = torch.cat([img.reshape(B, -1)] + parts, -1)
    return mlp(x, params_b, "baseline", n_hidden, True)


def moving_average(var, value, decay):
    """ops.py:46-64 [upstream assign_moving_average, zero_debias=False]: var <- var - (1 - decay) * (var - value)."""
    return var - (1.0 - decay) * (var - value)


def centered_rmsprop_step(theta, g, mg, ms, mom, lr, decay=0.9, momentum=0.9, eps=1e-10):
    """tf.train.RMSPropOptimizer(centered=True) [upstream ApplyCenteredRMSProp], model.py:265,355-360.
    ms starts at 1, mg and mom at 0; epsilon is inside the sqrt."""
    mg = mg + (1 - decay) * (g - mg)
    ms = ms + (1 - decay) * (g * g - ms)
    mom = momentum * mom + lr * g / torch.sqrt(ms - mg * mg + eps)
    return theta - mom, mg, ms, mom
