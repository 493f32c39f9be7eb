"""Thin tensor-level wrappers over the stand-alone entry points of libair_b200.so.

Every function takes CUDA float32 tensors, allocates the output with torch and enqueues the C-ABI call on
torch's current stream.  CPU tensors are rejected: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, current_stream_ptr, ptr

ACT_NONE, ACT_ELU = 0, 1


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.AirError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def linear(x, w, b=None, act=ACT_NONE, precision=_lib.AIR_PREC_FP32):
    """snt.Linear + transfer (neural.py:42-60): act(x @ w + b); w is [in, out]."""
    x, w = _cuda_f32(x, "x"), _cuda_f32(w, "w")
    b = None if b is None else _cuda_f32(b, "b")
    M, K = x.shape
    N = w.shape[1]
    assert w.shape[0] == K
    out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.lib().air_linear(ptr(x), ptr(w), ptr(b), ptr(out), M, N, K, act, precision, current_stream_ptr()),
              "air_linear")
    return out


def linear_backward(x, w, dy, dw=None, db=None, need_dx=False, x_is_elu_output=False):
    """Backward of one snt.Linear (+ the ELU of the layer that produced x): dw += x^T dy, db += colsum(dy) in place
    (callers zero them), returns dx = (dy @ w^T) * elu'(x) or None.  dy is the gradient at the layer's pre-activation."""
    x, w, dy = _cuda_f32(x, "x"), _cuda_f32(w, "w"), _cuda_f32(dy, "dy")
    M, K = x.shape
    N = w.shape[1]
    assert w.shape[0] == K and tuple(dy.shape) == (M, N)
    for t in (dw, db):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    dx = torch.empty(M, K, device=x.device, dtype=torch.float32) if need_dx else None
    with torch.cuda.device(x.device):
        check(_lib.lib().air_linear_backward(ptr(x), ptr(w), ptr(dy), ptr(x) if (x_is_elu_output and need_dx) else None,
                                             ptr(dw), ptr(db), ptr(dx), M, N, K, current_stream_ptr()),
              "air_linear_backward")
    return dx


def baseline_grad(target, baseline, target_mean, inv_batch):
    """d baseline_loss / d baseline (model.py:253-259 with the [B]-[B,1] broadcast) -> [B,1]."""
    target, baseline = _cuda_f32(target.reshape(-1), "target"), _cuda_f32(baseline.reshape(-1), "baseline")
    B = baseline.numel()
    out = torch.empty(B, 1, device=baseline.device, dtype=torch.float32)
    with torch.cuda.device(baseline.device):
        check(_lib.lib().air_baseline_grad(ptr(target), ptr(baseline), float(target_mean), float(inv_batch), ptr(out), B,
                                           current_stream_ptr()), "air_baseline_grad")
    return out


def baseline_grad_dev(baseline, target_mean_dev, inv_batch):
    """baseline_grad with the target mean read from a 1-element device tensor when the kernel runs (no host sync)."""
    baseline = _cuda_f32(baseline.reshape(-1), "baseline")
    assert target_mean_dev.is_cuda and target_mean_dev.dtype == torch.float32 and target_mean_dev.numel() >= 1
    B = baseline.numel()
    out = torch.empty(B, 1, device=baseline.device, dtype=torch.float32)
    with torch.cuda.device(baseline.device):
        check(_lib.lib().air_baseline_grad_dev(ptr(baseline), ptr(target_mean_dev), float(inv_batch), ptr(out), B,
                                               current_stream_ptr()), "air_baseline_grad_dev")
    return out


def lstm_step(x, h, c, w, b, forget_bias=1.0):
    """snt.LSTM step; returns new (h, c).  Gate order i, j, f, o; w is [nx + nh, 4 nh]."""
    x, w, b = _cuda_f32(x, "x"), _cuda_f32(w, "w"), _cuda_f32(b, "b")
    h = _cuda_f32(h, "h").clone()
    c = _cuda_f32(c, "c").clone()
    B, nx = x.shape
    nh = h.shape[1]
    with torch.cuda.device(x.device):
        check(_lib.lib().air_lstm_step(ptr(x), ptr(h), ptr(c), ptr(w), ptr(b), B, nx, nh, float(forget_bias),
                                       current_stream_ptr()), "air_lstm_step")
    return h, c


def stn_read(img, where, glimpse_hw):
    """SpatialTransformer forward (modules.py:94-109): crop[B,h,w] of img[B,H,W] at where[B,4]=(sx,tx,sy,ty)."""
    img, where = _cuda_f32(img, "img"), _cuda_f32(where, "where")
    B, H, W = img.shape
    h, w = glimpse_hw
    out = torch.empty(B, h, w, device=img.device, dtype=torch.float32)
    with torch.cuda.device(img.device):
        check(_lib.lib().air_stn_read(ptr(img), ptr(where), ptr(out), B, H, W, h, w, current_stream_ptr()),
              "air_stn_read")
    return out


def stn_paint(glimpse, where, canvas_hw):
    """inverse SpatialTransformer (modules.py:100-102): [B,H,W] gathered from glimpse[B,h,w]."""
    glimpse, where = _cuda_f32(glimpse, "glimpse"), _cuda_f32(where, "where")
    B, h, w = glimpse.shape
    H, W = canvas_hw
    out = torch.empty(B, H, W, device=glimpse.device, dtype=torch.float32)
    with torch.cuda.device(glimpse.device):
        check(_lib.lib().air_stn_paint(ptr(glimpse), ptr(where), ptr(out), B, H, W, h, w, current_stream_ptr()),
              "air_stn_paint")
    return out


def glimpse_viz(glimpse, presence):
    """model.py:90: presence * sigmoid(glimpse) for decoded glimpses [T,B,G] (or [T,B,h,w]) and presence [T,B,1]."""
    glimpse, presence = _cuda_f32(glimpse, "glimpse"), _cuda_f32(presence, "presence")
    rows = presence.numel()
    G = glimpse.numel() // max(rows, 1)
    out = torch.empty_like(glimpse)
    with torch.cuda.device(glimpse.device):
        check(_lib.lib().air_glimpse_viz(ptr(glimpse), ptr(presence), ptr(out), rows, G, current_stream_ptr()),
              "air_glimpse_viz")
    return out


def bernoulli_to_modified_geometric(presence_prob):
    """prior.py:62-68: [..., T] Bernoulli success probabilities -> [..., T+1] pmf over the number of steps."""
    p = _cuda_f32(presence_prob, "presence_prob")
    T = p.shape[-1]
    lead = p.shape[:-1]
    n = int(p.numel() // T) if T > 0 else 0
    out = torch.empty(*lead, T + 1, device=p.device, dtype=torch.float32)
    with torch.cuda.device(p.device):
        check(_lib.lib().air_bernoulli_to_modified_geometric(ptr(p), ptr(out), n, T, current_stream_ptr()),
              "air_bernoulli_to_modified_geometric")
    return out


def geometric_prior(success_prob, n_steps, device=None, float64=None):
    """prior.py:26-32.  A python float gives float32 maths like the reference graph; a float64 tensor / float64=True
    gives the float64 island used when the success probability comes out of _anneal_weight."""
    if isinstance(success_prob, torch.Tensor):
        if float64 is None:
            float64 = success_prob.dtype == torch.float64
        device = device or (success_prob.device if success_prob.is_cuda else None)
        success_prob = float(success_prob)
    float64 = bool(float64)
    device = torch.device(device or "cuda")
    out = torch.empty(n_steps + 1, device=device, dtype=torch.float64 if float64 else torch.float32)
    with torch.cuda.device(device):
        check(_lib.lib().air_geometric_prior(float(success_prob), int(n_steps), int(float64), ptr(out),
                                             current_stream_ptr()), "air_geometric_prior")
    return out


def tabular_kl(p, q, zero_prob_value=0.0):
    """prior.py:71-90: per-entry KL(p||q) of pmfs in tabular form, float64 inside, float32 out.  q broadcasts over rows."""
    p = _cuda_f32(p, "p")
    if not isinstance(q, torch.Tensor):
        q = torch.as_tensor(q, dtype=torch.float64)
    q = q.to(device=p.device, dtype=torch.float64)
    m = p.shape[-1]
    q_full = torch.broadcast_to(q, p.shape) if q.dim() > 1 else None
    out = torch.empty_like(p)
    with torch.cuda.device(p.device):
        if q_full is None:
            assert q.numel() == m
            check(_lib.lib().air_tabular_kl(ptr(p), ptr(q.contiguous()), ptr(out), int(p.numel() // m), m,
                                            float(zero_prob_value), current_stream_ptr()), "air_tabular_kl")
        else:
            # row-specific q: treat the whole table as one row of length numel
            qf = q_full.contiguous().reshape(-1)
            check(_lib.lib().air_tabular_kl(ptr(p), ptr(qf), ptr(out), 1, int(p.numel()), float(zero_prob_value),
                                            current_stream_ptr()), "air_tabular_kl")
    return out


def sample_from_tensor(pmf, samples):
    """prior.py:103-116: pmf[b, int(samples[b])] (flat gather, int32 index)."""
    pmf, samples = _cuda_f32(pmf, "pmf"), _cuda_f32(samples, "samples")
    n, m = pmf.shape
    out = torch.empty(n, device=pmf.device, dtype=torch.float32)
    with torch.cuda.device(pmf.device):
        check(_lib.lib().air_sample_from_tensor(ptr(pmf), ptr(samples.reshape(-1)), ptr(out), n, m,
                                                current_stream_ptr()), "air_sample_from_tensor")
    return out.reshape(samples.shape)


def num_steps_log_prob(pmf, samples):
    """prior.py:141-151: log(max(pmf[b, int(samples[b])], 1e-32))."""
    pmf, samples = _cuda_f32(pmf, "pmf"), _cuda_f32(samples, "samples")
    n, m = pmf.shape
    out = torch.empty(n, device=pmf.device, dtype=torch.float32)
    with torch.cuda.device(pmf.device):
        check(_lib.lib().air_num_steps_log_prob(ptr(pmf), ptr(samples.reshape(-1)), ptr(out), n, m,
                                                current_stream_ptr()), "air_num_steps_log_prob")
    return out.reshape(samples.shape)


def anneal_weight(init_val, final_val, anneal_type, global_step, anneal_steps, hold_for=0.0, steps_div=1.0):
    """model.py:106-124 (float64 scalar schedule)."""
    kinds = {"exp": 0, "linear": 1}
    if anneal_type not in kinds:
        raise NotImplementedError(anneal_type)
    return float(_lib.lib().air_anneal_weight(float(init_val), float(final_val), kinds[anneal_type],
                                              float(global_step), float(anneal_steps), float(hold_for),
                                              float(steps_div)))
