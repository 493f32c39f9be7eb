"""Engine: one air_handle (libair_b200.so) + the caller-owned output buffers of the fused forward pass.

This is the only place where the Python layer talks to the hot-path entry points of the C ABI
(air_create / air_forward / air_forward_host / air_cell_step / air_destroy).  torch provides device memory and the
stream; nothing is computed in torch.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import AIR_MAX_HIDDEN, AIR_MAX_STEPS, AIR_N_SCALARS, SCALAR_INDEX, air_config, air_outputs, air_prior
from ._lib import check, current_stream_ptr, ptr


@dataclass
class CellConfig:
    """Hyper-parameters of the path; defaults = scripts/multi_mnist.py:24-94 + mnist_model.py:13-44."""
    H: int = 50
    W: int = 50
    h: int = 20
    w: int = 20
    na: int = 50
    nh: int = 256
    enc_hidden: Sequence[int] = (256, 256)
    glenc_hidden: Sequence[int] = (256, 256)
    dec_hidden: Sequence[int] = (256, 256)
    where_hidden: Sequence[int] = (256, 256)
    steps_hidden: Sequence[int] = (128, 64)
    output_std: float = 0.3
    output_multiplier: float = 0.5
    explore_eps: Optional[float] = 1e-3
    scale_bias: float = 0.5
    step_bias: float = 0.75
    what_scale_offset: float = 0.5
    forget_bias: float = 1.0
    max_crop_size: float = 1.0
    discrete_steps: bool = True
    precision: int = _lib.AIR_PREC_FP32

    @property
    def P(self):
        return self.H * self.W

    @property
    def G(self):
        return self.h * self.w


def param_spec(cfg: CellConfig) -> List[Tuple[str, Tuple[int, int]]]:
    """Canonical flat parameter layout (the Sonnet variables of cell.py:61-69 in creation order); weights [in, out]
    row-major like snt.Linear.  Must equal the table the C library reports (checked in Engine.__init__)."""
    spec: List[Tuple[str, Tuple[int, int]]] = []

    def mlp(prefix, n_in, hidden, n_out=None):
        d = n_in
        for i, n in enumerate(hidden):
            spec.append((f"{prefix}.{i}.w", (d, n)))
            spec.append((f"{prefix}.{i}.b", (1, n)))
            d = n
        if n_out is not None:
            spec.append((f"{prefix}.out.w", (d, n_out)))
            spec.append((f"{prefix}.out.b", (1, n_out)))
        return d

    n_enc = mlp("input_encoder", cfg.P, cfg.enc_hidden)
    spec.append(("lstm.w", (n_enc + cfg.nh, 4 * cfg.nh)))
    spec.append(("lstm.b", (1, 4 * cfg.nh)))
    spec.append(("lstm.h0", (1, cfg.nh)))
    spec.append(("lstm.c0", (1, cfg.nh)))
    mlp("transform_estimator", cfg.nh, cfg.where_hidden, 8)
    mlp("steps_predictor", cfg.nh, cfg.steps_hidden, 1)
    n_gl = mlp("glimpse_encoder", cfg.G, cfg.glenc_hidden)
    spec.append(("what.w", (n_gl, 2 * cfg.na)))
    spec.append(("what.b", (1, 2 * cfg.na)))
    mlp("glimpse_decoder", cfg.na, cfg.dec_hidden, cfg.G)
    return spec


def param_count(cfg: CellConfig) -> int:
    return sum(r * c for _, (r, c) in param_spec(cfg))


def make_views(spec, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    views, off = {}, 0
    for name, (r, c) in spec:
        n = r * c
        v = flat[off:off + n]
        views[name] = v.view(c) if (r == 1 and not name.endswith(".w")) else v.view(r, c)
        off += n
    assert off == flat.numel()
    return views


def _fill_hidden(dst, src):
    src = list(src)
    if not (1 <= len(src) <= AIR_MAX_HIDDEN):
        raise _lib.AirError(f"an MLP needs 1..{AIR_MAX_HIDDEN} hidden layers, got {len(src)}")
    for i, v in enumerate(src):
        dst[i] = int(v)
    return len(src)


def make_prior(what_prior=None, where_scale_prior=None, where_shift_prior=None, steps_success_prob=0.5,
               steps_prob_is_f64=True, steps_weight=1.0, analytic=True, use_prior=True, use_reinforce=True,
               nvil_shift=0.0, nvil_scale=0.0) -> air_prior:
    """Pack the AttrDict-style priors of AIRModel.train_step (model.py:261-265) into the C struct."""
    def get(d, k, default):
        if d is None:
            return default
        if isinstance(d, dict):
            return d.get(k, default)
        return getattr(d, k, default)

    p = air_prior()
    p.what_loc = float(get(what_prior, "loc", 0.0))
    p.what_scale = float(get(what_prior, "scale", 1.0))
    p.where_scale_loc = float(get(where_scale_prior, "loc", 0.0))
    p.where_scale_scale = float(get(where_scale_prior, "scale", 1.0))
    shift_loc = get(where_shift_prior, "loc", None)
    p.where_shift_has_loc = int(shift_loc is not None)
    p.where_shift_loc = float(shift_loc if shift_loc is not None else 0.0)
    p.where_shift_scale = float(get(where_shift_prior, "scale", 1.0))
    p.steps_success_prob = float(steps_success_prob)
    p.steps_prob_is_f64 = int(bool(steps_prob_is_f64))
    p.steps_weight = float(steps_weight)
    p.analytic = int(bool(analytic))
    p.use_prior = int(bool(use_prior))
    p.use_reinforce = int(bool(use_reinforce))
    p.nvil_shift, p.nvil_scale = float(nvil_shift), float(nvil_scale)      # scale 0 = normalisation off
    return p


def c_config(cfg: CellConfig, B: int, T: int) -> air_config:
    """The C struct of a configuration (include/air_b200.h: air_config)."""
    c = air_config()
    c.B, c.H, c.W, c.h, c.w, c.T, c.na, c.nh = int(B), cfg.H, cfg.W, cfg.h, cfg.w, int(T), cfg.na, cfg.nh
    c.n_enc_hidden = _fill_hidden(c.enc_hidden, cfg.enc_hidden)
    c.n_glenc_hidden = _fill_hidden(c.glenc_hidden, cfg.glenc_hidden)
    c.n_dec_hidden = _fill_hidden(c.dec_hidden, cfg.dec_hidden)
    c.n_where_hidden = _fill_hidden(c.where_hidden, cfg.where_hidden)
    c.n_steps_hidden = _fill_hidden(c.steps_hidden, cfg.steps_hidden)
    c.output_std = float(cfg.output_std)
    c.output_multiplier = float(cfg.output_multiplier)
    c.explore_eps = -1.0 if cfg.explore_eps is None else float(cfg.explore_eps)
    c.scale_bias = float(cfg.scale_bias)
    c.step_bias = float(cfg.step_bias)
    c.what_scale_offset = float(cfg.what_scale_offset)
    c.forget_bias = float(cfg.forget_bias)
    c.max_crop_size = float(cfg.max_crop_size)
    c.discrete_steps = int(bool(cfg.discrete_steps))
    c.precision = int(cfg.precision)
    return c


def row_schedule_check(cfg: CellConfig, B: int = 64, T: int = 3) -> str:
    """Host-only replay of the fused row kernel's schedule for `cfg` (air_row_schedule_check): "" when the row kernel
    covers the configuration, else the reason it does not.  Needs no GPU."""
    c = c_config(cfg, B, T)
    msg = C.create_string_buffer(256)
    _lib.lib().air_row_schedule_check(C.byref(c), msg, 256)
    return msg.value.decode()


class Engine:
    """Owns an air_handle for (cfg, B, T) on one device and the output tensors of the fused call."""

    CELL_OUTPUTS = "canvas glimpse what what_loc what_scale where where_loc where_scale presence_prob presence".split()

    def __init__(self, cfg: CellConfig, B: int, T: int, device=None, materialise_canvas=True, materialise_viz=True):
        if not torch.cuda.is_available():
            raise _lib.AirError("no CUDA device: the AIR hot path has no CPU fallback")
        self.cfg, self.B, self.T = cfg, int(B), int(T)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = _lib.lib()
        c = c_config(cfg, self.B, self.T)
        self._c_cfg = c
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.air_create(C.byref(c), C.byref(self._handle)), "air_create")
        # the C library's parameter table must be the layout the Python side assumes
        spec = param_spec(cfg)
        n = self.lib.air_param_entries(self._handle)
        assert n == len(spec), (n, len(spec))
        off = 0
        for i, (name, (r, cc)) in enumerate(spec):
            nm, o, rr, cl = C.c_char_p(), C.c_int64(), C.c_int32(), C.c_int32()
            check(self.lib.air_param_entry(self._handle, i, C.byref(nm), C.byref(o), C.byref(rr), C.byref(cl)))
            assert (nm.value.decode(), o.value, rr.value, cl.value) == (name, off, r, cc), (nm.value, name)
            off += r * cc
        self.n_params = int(self.lib.air_param_count(self._handle))
        assert self.n_params == off
        self._alloc_outputs(materialise_canvas, materialise_viz)

    # ------------------------------------------------------------------------------------------
    def _alloc_outputs(self, canvas: bool, viz: bool):
        T, B, cfg, dev = self.T, self.B, self.cfg, self.device
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        o = {
            "canvas": f(T, B, cfg.P) if canvas else None,
            "glimpse": f(T, B, cfg.G),
            "glimpse_viz": f(T, B, cfg.G) if viz else None,
            "what": f(T, B, cfg.na), "what_loc": f(T, B, cfg.na), "what_scale": f(T, B, cfg.na),
            "where": f(T, B, 4), "where_loc": f(T, B, 4), "where_scale": f(T, B, 4),
            "presence_prob": f(T, B, 1), "presence": f(T, B, 1),
            "final_h": f(B, cfg.nh), "final_c": f(B, cfg.nh),
            "num_steps_posterior": f(B, T + 1), "num_step_per_sample": f(B), "prior_step_weight": f(T, B),
            "rec_loss_per_sample": f(B), "kl_num_steps_per_sample": f(B), "kl_what_per_sample": f(B),
            "kl_where_per_sample": f(B), "loss_per_sample": f(B), "num_steps_log_prob": f(B),
            "scalars": torch.zeros(AIR_N_SCALARS, device=dev, dtype=torch.float32),
        }
        self.out = o
        s = air_outputs()
        for name in _lib.OUTPUT_FIELDS:
            t = o[name]
            setattr(s, name, None if t is None else t.data_ptr())
        self._c_out = s

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.air_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.air_workspace_bytes(self._handle))

    # ------------------------------------------------------------------------------------------
    def _chk(self, t, shape, name):
        if t is None:
            return None
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.AirError(f"{name} must be a contiguous CUDA float32 tensor")
        if t.numel() != math.prod(shape):
            raise _lib.AirError(f"{name} has {tuple(t.shape)}, expected {tuple(shape)}")
        return t

    def forward(self, params, img, eps_where, eps_what, u_pres, prior: Optional[air_prior] = None, baseline=None):
        """T unrolled steps + ELBO terms: one air_forward enqueue on torch's current stream.  Returns self.out."""
        T, B, cfg = self.T, self.B, self.cfg
        self._chk(params, (self.n_params,), "params")
        self._chk(img, (B, cfg.H, cfg.W), "img")
        self._chk(eps_where, (T, B, 4), "eps_where")
        self._chk(eps_what, (T, B, cfg.na), "eps_what")
        self._chk(u_pres, (T, B, 1), "u_pres")
        self._chk(baseline, (B,), "baseline")
        with torch.cuda.device(self.device):
            check(self.lib.air_forward(self._handle, ptr(params), ptr(img), ptr(eps_where), ptr(eps_what), ptr(u_pres),
                                       ptr(baseline), C.byref(prior) if prior is not None else None,
                                       C.byref(self._c_out), current_stream_ptr()), "air_forward")
        return self.out

    def forward_host(self, params, img_host, eps_where_host, eps_what_host, u_pres_host, prior: air_prior,
                     scalars_host, loss_per_sample_host):
        """End-to-end call with HOST inputs/outputs (H2D + forward + D2H + stream sync inside the C call)."""
        for t in (img_host, eps_where_host, eps_what_host, u_pres_host, scalars_host, loss_per_sample_host):
            assert t is not None and not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        with torch.cuda.device(self.device):
            check(self.lib.air_forward_host(self._handle, ptr(params), ptr(img_host), ptr(eps_where_host),
                                            ptr(eps_what_host), ptr(u_pres_host), C.byref(prior),
                                            C.byref(self._c_out), ptr(scalars_host), ptr(loss_per_sample_host),
                                            current_stream_ptr()), "air_forward_host")
        return scalars_host, loss_per_sample_host

    def elbo_scalars(self, baseline, prior: air_prior):
        """Re-form the batch means with a baseline [B] (model.py:224-230,247-248)."""
        self._chk(baseline, (self.B,), "baseline")
        with torch.cuda.device(self.device):
            check(self.lib.air_elbo_scalars(self._handle, ptr(baseline.contiguous()), C.byref(prior),
                                            C.byref(self._c_out), current_stream_ptr()), "air_elbo_scalars")
        return self.out["scalars"]

    def forward_host_u8(self, params, img_u8_host, eps_where_host, eps_what_host, u_pres_host, prior: air_prior,
                        scalars_host, loss_per_sample_host):
        """forward_host with the images in the reference's dataset format (uint8 [B,H,W]); /255 runs on the device."""
        assert img_u8_host.dtype == torch.uint8 and not img_u8_host.is_cuda and img_u8_host.is_contiguous()
        for t in (eps_where_host, eps_what_host, u_pres_host, scalars_host, loss_per_sample_host):
            assert t is not None and not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        with torch.cuda.device(self.device):
            check(self.lib.air_forward_host_u8(self._handle, ptr(params), ptr(img_u8_host), ptr(eps_where_host),
                                               ptr(eps_what_host), ptr(u_pres_host), C.byref(prior),
                                               C.byref(self._c_out), ptr(scalars_host), ptr(loss_per_sample_host),
                                               current_stream_ptr()), "air_forward_host_u8")
        return scalars_host, loss_per_sample_host

    def cache_weights(self, on: bool = True):
        """Inference loops: reuse the prepared tensor-core weight arena while `params` is unchanged (call
        params_updated() after writing to the buffer)."""
        check(self.lib.air_cache_weights(self._handle, int(on)), "air_cache_weights")

    def set_launch_overlap(self, on: bool = True):
        """Programmatic dependent launch between the kernels of a pass: on for a handle that has the device to itself, off
        when several handles' passes are in flight (EnginePool does this)."""
        check(self.lib.air_set_launch_overlap(self._handle, int(on)), "air_set_launch_overlap")

    def params_updated(self):
        check(self.lib.air_params_updated(self._handle), "air_params_updated")

    def prior_table_device(self, prior: Optional[air_prior]):
        """Replayed CUDA graphs: write geometric_prior(prior.steps_success_prob, T) into the handle's device table (a one-warp
        kernel on the current stream) and make forward() / backward() read it from there, so that a captured step follows the
        annealed prior (model.py:133-142).  None switches back to the table in the kernel arguments."""
        with torch.cuda.device(self.device):
            check(self.lib.air_prior_table_device(self._handle, C.byref(prior) if prior is not None else None,
                                                  current_stream_ptr()), "air_prior_table_device")

    # -- training step -----------------------------------------------------------------
    def train_enable(self, on: bool = True):
        """Keep the activations of every following forward() for backward() (either engine: on an AIR_PREC_TC_SPLIT
        handle the training forward and the gradient GEMMs run on the tensor cores)."""
        with torch.cuda.device(self.device):
            check(self.lib.air_train_enable(self._handle, int(on)), "air_train_enable")

    @property
    def train_workspace_bytes(self) -> int:
        return int(self.lib.air_train_workspace_bytes(self._handle))

    def backward(self, params, img, eps_where, eps_what, prior: air_prior, grad: Optional[torch.Tensor] = None,
                 baseline_mean: float = 0.0, inv_batch: float = 0.0, l2_weight: float = 0.0) -> torch.Tensor:
        """d opt_loss / d params for the batch of the LAST forward() (model.py:355-356).  Returns the flat gradient."""
        if grad is None:
            grad = torch.empty(self.n_params, device=self.device, dtype=torch.float32)
        self._chk(grad, (self.n_params,), "grad")
        with torch.cuda.device(self.device):
            check(self.lib.air_backward(self._handle, ptr(params), ptr(img), ptr(eps_where), ptr(eps_what),
                                        C.byref(prior), C.byref(self._c_out), float(baseline_mean), float(inv_batch),
                                        float(l2_weight), ptr(grad), current_stream_ptr()), "air_backward")
        return grad

    # ---- BaselineMLP on the engine (modules.py:125-143; air_baseline_*) ------------------------------------------
    def baseline_attach(self, hidden: Sequence[int]):
        """Register a BaselineMLP(hidden) with this handle (before train_enable).  Returns (n_in, n_params)."""
        arr = (C.c_int32 * len(hidden))(*[int(v) for v in hidden])
        with torch.cuda.device(self.device):
            check(self.lib.air_baseline_attach(self._handle, len(hidden), arr), "air_baseline_attach")
        return int(self.lib.air_baseline_input_width(self._handle)), int(self.lib.air_baseline_param_count(self._handle))

    def baseline_forward(self, bparams, img, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """baseline [B,1] for the batch of the LAST forward(): input rows gathered from `img` and this engine's cell outputs,
        first layer on the engine's GEMM path."""
        if out is None:
            out = torch.empty(self.B, 1, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.air_baseline_forward(self._handle, ptr(bparams), ptr(img), C.byref(self._c_out), ptr(out),
                                                current_stream_ptr()), "air_baseline_forward")
        return out

    def baseline_backward(self, bparams, d_baseline, bgrad, defer_join: bool = False) -> torch.Tensor:
        """d baseline_loss / d baseline parameters for the last baseline_forward(), into bgrad (flat).  ``defer_join``: the
        weight-gradient GEMMs stay on the side streams and bgrad is complete only after the next backward() on this engine
        (air_baseline_backward_async)."""
        fn = self.lib.air_baseline_backward_async if defer_join else self.lib.air_baseline_backward
        with torch.cuda.device(self.device):
            check(fn(self._handle, ptr(bparams), ptr(d_baseline), ptr(bgrad), current_stream_ptr()), "air_baseline_backward")
        return bgrad

    def rmsprop_step(self, params, grad, mg, ms, mom, learning_rate, decay=0.9, momentum=0.9, epsilon=1e-10,
                     grad_scale=1.0):
        """Centered RMSProp with momentum, TF semantics (model.py:265,355-360), in place on the flat buffers."""
        for t in (params, grad, mg, ms, mom):
            self._chk(t, (params.numel(),), "rmsprop buffer")
        with torch.cuda.device(self.device):
            check(self.lib.air_rmsprop_step(ptr(params), ptr(grad), ptr(mg), ptr(ms), ptr(mom), params.numel(),
                                            float(learning_rate), float(decay), float(momentum), float(epsilon),
                                            float(grad_scale), current_stream_ptr()), "air_rmsprop_step")

    def draw_noise(self, seed: int):
        """(eps_where[T,B,4], eps_what[T,B,na], u_pres[T,B,1]) drawn inside the library (Philox4x32-10, counter-based:
        the same seed gives the same tensors on any device / shard layout)."""
        T, B, cfg, dev = self.T, self.B, self.cfg, self.device
        ew = torch.empty(T, B, 4, device=dev)
        ea = torch.empty(T, B, cfg.na, device=dev)
        u = torch.empty(T, B, 1, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.air_draw_noise(self._handle, int(seed), ptr(ew), ptr(ea), ptr(u), current_stream_ptr()),
                  "air_draw_noise")
        return ew, ea, u

    def forward_host_u8_rng(self, params, img_u8_host, seed: int, prior: air_prior, scalars_host, loss_per_sample_host):
        """forward_host_u8 with the noise drawn on the device from ``seed``: only the uint8 images cross the bus."""
        assert img_u8_host.dtype == torch.uint8 and not img_u8_host.is_cuda and img_u8_host.is_contiguous()
        with torch.cuda.device(self.device):
            check(self.lib.air_forward_host_u8_rng(self._handle, ptr(params), ptr(img_u8_host), int(seed), C.byref(prior),
                                                   C.byref(self._c_out), ptr(scalars_host), ptr(loss_per_sample_host),
                                                   current_stream_ptr()), "air_forward_host_u8_rng")
        return scalars_host, loss_per_sample_host

    # ---- double-buffered host feed: the copy of batch i+1 overlaps the pass over batch i (data.py:121-158's input queue) ----
    def feed_host_u8(self, slot: int, img_u8_host):
        """Enqueue the host->device copy of one pinned uint8 batch [B,H,W] into staging slot 0/1; returns at once."""
        assert img_u8_host.dtype == torch.uint8 and not img_u8_host.is_cuda and img_u8_host.is_contiguous()
        assert img_u8_host.numel() == self.B * self.cfg.H * self.cfg.W, "feed_host_u8: wrong batch shape"
        with torch.cuda.device(self.device):
            check(self.lib.air_feed_host_u8(self._handle, int(slot), ptr(img_u8_host)), "air_feed_host_u8")

    def forward_fed_u8_rng(self, params, slot: int, seed: int, prior: air_prior, scalars_host, loss_per_sample_host):
        """forward_host_u8_rng on the batch fed into ``slot``; does NOT synchronise: call feed_wait(slot) before reading the
        host buffers (which must stay alive, and be distinct per slot, until then)."""
        self._chk(params, (self.n_params,), "params")
        for t, n in ((scalars_host, _lib.AIR_N_SCALARS), (loss_per_sample_host, self.B)):
            assert t is None or (not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n)
        with torch.cuda.device(self.device):
            check(self.lib.air_forward_fed_u8_rng(self._handle, ptr(params), int(slot), int(seed), C.byref(prior),
                                                  C.byref(self._c_out), ptr(scalars_host), ptr(loss_per_sample_host),
                                                  current_stream_ptr()), "air_forward_fed_u8_rng")

    def feed_wait(self, slot: int):
        """Block until the results of the last forward_fed_u8_rng on ``slot`` are in host memory."""
        check(self.lib.air_feed_wait(self._handle, int(slot)), "air_feed_wait")

    def stream_host_u8(self, params, batches, prior: air_prior, seed0: int = 0):
        """Run the forward + ELBO pass over an iterable of pinned uint8 host batches with the double-buffered feed and yield
        (scalars [AIR_N_SCALARS], loss_per_sample [B]) host tensors per batch, in order.  The two tensors are the slot's
        pinned buffers: read (or clone) them before the next ``next()``, which enqueues the pass that overwrites them.
        Batch i uses noise seed ``seed0 + i``."""
        scal = [torch.empty(_lib.AIR_N_SCALARS).pin_memory() for _ in range(2)]
        lps = [torch.empty(self.B).pin_memory() for _ in range(2)]
        it = iter(batches)
        alive = [next(it, None), None]                   # a host batch must outlive its copy: one reference per slot
        if alive[0] is None:
            return
        self.feed_host_u8(0, alive[0])
        i = 0
        while alive[i % 2] is not None:
            alive[(i + 1) % 2] = next(it, None)
            if alive[(i + 1) % 2] is not None:
                self.feed_host_u8((i + 1) % 2, alive[(i + 1) % 2])
            self.forward_fed_u8_rng(params, i % 2, seed0 + i, prior, scal[i % 2], lps[i % 2])
            if i >= 1:
                self.feed_wait((i - 1) % 2)
                yield scal[(i - 1) % 2], lps[(i - 1) % 2]
            i += 1
        self.feed_wait((i - 1) % 2)
        yield scal[(i - 1) % 2], lps[(i - 1) % 2]

    def forward_dataset_u8(self, params, dataset_u8, idx, eps_where, eps_what, u_pres, prior: Optional[air_prior] = None,
                           baseline=None, img_out=None):
        """forward() on the minibatch ``dataset_u8[idx]`` of a device-resident uint8 dataset [N,H,W]:
        gather, /255 and operand preparation run on the device.  ``img_out`` [B,H,W] receives the float32 minibatch."""
        T, B, cfg = self.T, self.B, self.cfg
        assert dataset_u8.is_cuda and dataset_u8.dtype == torch.uint8 and dataset_u8.is_contiguous()
        assert tuple(dataset_u8.shape[1:]) == (cfg.H, cfg.W)
        assert idx.is_cuda and idx.dtype == torch.int32 and idx.numel() == B and idx.is_contiguous()
        self._chk(params, (self.n_params,), "params")
        self._chk(eps_where, (T, B, 4), "eps_where")
        self._chk(eps_what, (T, B, cfg.na), "eps_what")
        self._chk(u_pres, (T, B, 1), "u_pres")
        self._chk(img_out, (B, cfg.H, cfg.W), "img_out")
        with torch.cuda.device(self.device):
            check(self.lib.air_forward_dataset_u8(self._handle, ptr(params), ptr(dataset_u8), dataset_u8.shape[0], ptr(idx),
                                                  ptr(eps_where), ptr(eps_what), ptr(u_pres), ptr(baseline),
                                                  C.byref(prior) if prior is not None else None, C.byref(self._c_out),
                                                  ptr(img_out), current_stream_ptr()), "air_forward_dataset_u8")
        return self.out

    def iwae_bound(self, K: int, prior: air_prior):
        """Importance-weighted bound of the LAST forward(), whose B rows are B/K canvases x K particles (row = canvas * K +
        particle; feed ``img.repeat_interleave(K, 0)`` and independent noise).  Returns (mean bound [scalar tensor],
        bound per canvas [B/K], log_w [B])."""
        if self.B % K:
            raise _lib.AirError(f"batch {self.B} is not a multiple of K = {K}")
        n, o, dev = self.B // K, self.out, self.device
        log_w = torch.empty(self.B, device=dev)
        bound = torch.empty(n, device=dev)
        mean = torch.empty(1, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.air_iwae_bound(n, K, self.T, self.cfg.na, ptr(o["what"]), ptr(o["what_loc"]),
                                          ptr(o["what_scale"]), ptr(o["where"]), ptr(o["where_loc"]),
                                          ptr(o["where_scale"]), ptr(o["presence"]), ptr(o["rec_loss_per_sample"]),
                                          ptr(o["num_steps_log_prob"]), C.byref(prior), ptr(log_w), ptr(bound), ptr(mean),
                                          current_stream_ptr()), "air_iwae_bound")
        return mean[0], bound, log_w

    def cell_step(self, params, img, canvas, h, c, presence, eps_where, eps_what, u_pres):
        """One AIRCell step (cell.py:116-171); canvas / h / c / presence are updated IN PLACE.  Returns the per-step
        outputs glimpse, what, what_loc, what_scale, where, where_loc, where_scale, presence_prob."""
        B, cfg, dev = self.B, self.cfg, self.device
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        o = dict(glimpse=f(B, cfg.G), what=f(B, cfg.na), what_loc=f(B, cfg.na), what_scale=f(B, cfg.na),
                 where=f(B, 4), where_loc=f(B, 4), where_scale=f(B, 4), presence_prob=f(B, 1))
        with torch.cuda.device(dev):
            check(self.lib.air_cell_step(self._handle, ptr(params), ptr(img), ptr(canvas), ptr(h), ptr(c),
                                         ptr(presence), ptr(eps_where), ptr(eps_what), ptr(u_pres),
                                         ptr(o["glimpse"]), ptr(o["what"]), ptr(o["what_loc"]), ptr(o["what_scale"]),
                                         ptr(o["where"]), ptr(o["where_loc"]), ptr(o["where_scale"]),
                                         ptr(o["presence_prob"]), current_stream_ptr()), "air_cell_step")
        return o

    def check_range(self):
        """Raise if the tensor-core split engine saw an operand outside the fp16 range (synchronises)."""
        with torch.cuda.device(self.device):
            check(self.lib.air_check_range(self._handle, current_stream_ptr()), "air_check_range")

    # -- instrumentation -------------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self.lib.air_launch_count(self._handle))

    def profile(self, on: bool):
        check(self.lib.air_profile_enable(self._handle, int(on)), "air_profile_enable")

    def stage_times_ms(self) -> Dict[str, float]:
        """Device time of each stage of the last forward (CUDA events on the launching stream)."""
        buf = (C.c_float * _lib.AIR_N_STAGES)()
        check(self.lib.air_profile_read(self._handle, buf, _lib.AIR_N_STAGES), "air_profile_read")
        return {self.lib.air_stage_name(i).decode(): float(buf[i]) for i in range(_lib.AIR_N_STAGES)}

    def scalar(self, name: str) -> torch.Tensor:
        return self.out["scalars"][SCALAR_INDEX[name]]


class EnginePool:
    """``n_streams`` engines of the same (cfg, B, T), each with its own CUDA stream, workspace and output tensors.

    Consecutive batches are enqueued round-robin, so kernels of DIFFERENT batches may run side by side: the tensor-core
    kernels of one pass hold 96-128 of the 148 SMs with one CTA each and leave the SIMT pipes idle, the per-canvas
    kernels (where_read, paint) do the opposite, and every kernel has a tail in which SMs drain.  With several batches in
    flight the B200 spends that slack on the neighbouring batches (measured at B = 4096: 0.268 ms per batch alone, 0.246
    with two, 0.217 with three, 0.206 with four, 0.205 with six).  The results of a batch are in the ``out`` dict of the engine that ran it;
    they are valid on that engine's stream -- call ``join()`` (or ``torch.cuda.current_stream().wait_stream(stream)``)
    before consuming them elsewhere.  Input tensors must stay alive until the batch has run (they are used on a stream
    other than the one they were allocated on).
    """

    def __init__(self, cfg: CellConfig, B: int, T: int, n_streams: int = 4, device=None, **engine_kwargs):
        assert n_streams >= 1
        self.engines = [Engine(cfg, B, T, device=device, **engine_kwargs) for _ in range(n_streams)]
        self.device = self.engines[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self._i = 0
        for e in self.engines:      # with neighbours on the device, an early-launched CTA that only waits wastes its SM
            e.set_launch_overlap(n_streams == 1)

    def __len__(self):
        return len(self.engines)

    def cache_weights(self, on: bool = True):
        for e in self.engines:
            e.cache_weights(on)

    def params_updated(self):
        for e in self.engines:
            e.params_updated()

    def next(self):
        """Context manager: the next engine of the round robin, with its stream made current.  The stream first waits
        for the work already enqueued on the caller's stream (the producers of this batch's inputs)."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            i = self._i
            self._i = (i + 1) % len(self.engines)
            st = self.streams[i]
            st.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(st):
                yield self.engines[i]
        return ctx()

    def forward(self, params, img, eps_where, eps_what, u_pres, prior: Optional[air_prior] = None, baseline=None):
        """Engine.forward on the next engine / stream of the pool; returns (engine, engine.out)."""
        with self.next() as e:
            return e, e.forward(params, img, eps_where, eps_what, u_pres, prior, baseline)

    def join(self):
        """Make torch's current stream wait for everything enqueued on the pool's streams."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)

    def stream_host_u8(self, params, batches, prior: air_prior, seed0: int = 0):
        """Engine.stream_host_u8 over the whole pool: pinned uint8 host batches in, (scalars, loss_per_sample) host tensors
        out, in order; batch i runs on engine i % n with noise seed ``seed0 + i``.  Every engine double-buffers its own
        feed, so up to 2 n batches are in flight between the host and the device."""
        n = len(self.engines)
        scal = [[torch.empty(_lib.AIR_N_SCALARS).pin_memory() for _ in range(2)] for _ in range(n)]
        lps = [[torch.empty(e.B).pin_memory() for _ in range(2)] for e in self.engines]
        it = iter(batches)
        alive = {}                                        # batch index -> host tensor (a batch must outlive its copy)

        def feed(i):
            b = next(it, None)
            if b is not None:
                alive[i] = b
                self.engines[i % n].feed_host_u8((i // n) % 2, b)
            return b is not None

        fed = 0
        while fed < n and feed(fed):
            fed += 1
        i = 0
        while i < fed:
            e, k = self.engines[i % n], (i // n) % 2
            if fed == i + n and feed(fed):                # the batch this engine runs next time goes into its other slot
                fed += 1
            with torch.cuda.stream(self.streams[i % n]):
                e.forward_fed_u8_rng(params, k, seed0 + i, prior, scal[i % n][k], lps[i % n][k])
            if i >= n:                                    # the engine's previous batch: its results are (long) done
                e.feed_wait(1 - k)
                alive.pop(i - n, None)
                yield scal[i % n][1 - k], lps[i % n][1 - k]
            i += 1
        for j in range(max(0, fed - n), fed):
            self.engines[j % n].feed_wait((j // n) % 2)
            alive.pop(j, None)
            yield scal[j % n][(j // n) % 2], lps[j % n][(j // n) % 2]

    def close(self):
        for e in self.engines:
            e.close()
