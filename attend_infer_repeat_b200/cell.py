"""AIRCell -- mirror of the reference's cell.py (cell.py:9-171) over the CUDA library.

Same constructor, ``state_size`` / ``output_size`` / ``output_names``, ``initial_state(img)`` and
``cell(inpt, state) -> (outputs[10], state[6])`` contract as the Sonnet RNNCore of the reference.  A single call is one
``air_cell_step``; the unrolled T-step path used by AIRModel goes through ``Engine.forward`` (one fused enqueue).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .engine import CellConfig, Engine, make_views, param_count, param_spec
from .modules import (LSTM, Decoder, Encoder, ParametrisedGaussian, SpatialTransformer, StepsPredictor,
                      StochasticTransformParam)


def _init_flat(spec, device, seed=0):
    """Effective reference initialiser: truncated normal sigma = 1/sqrt(fan_in) cut at 2 sigma for
    weights, zeros for biases and the trainable LSTM initial state.  Drawn on the host for reproducibility."""
    g = torch.Generator().manual_seed(seed)
    chunks = []
    for name, (r, c) in spec:
        if name.endswith(".w"):
            std = 1.0 / math.sqrt(r)
            t = torch.empty(r, c, dtype=torch.float32)
            torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=g)
        else:
            t = torch.zeros(r, c, dtype=torch.float32)
        chunks.append(t.reshape(-1))
    flat = torch.cat(chunks).to(device)
    return flat, make_views(spec, flat)


def _scalar(v) -> float:
    return float(v.item()) if isinstance(v, torch.Tensor) else float(v)


class AIRCell:
    """RNN cell implementing one Attend-Infer-Repeat step (https://arxiv.org/abs/1603.08575)."""
    _n_transform_param = 4

    def __init__(self, img_size, crop_size, n_appearance,
                 transition, input_encoder, glimpse_encoder, glimpse_decoder, transform_estimator, steps_predictor,
                 discrete_steps=True, canvas_init=None, explore_eps=None, debug=False,
                 output_std=1.0, output_multiplier=1.0, precision=_lib.AIR_PREC_FP32, device=None, seed=0):
        """Arguments up to ``debug`` are those of the reference (cell.py:15-17); the rest are build options of this
        implementation (loss constants forwarded by AIRModel, arithmetic mode, device, init seed)."""
        self._img_size = tuple(int(i) for i in img_size)
        self._n_pix = int(np.prod(self._img_size))
        self._crop_size = tuple(int(i) for i in crop_size)
        self._n_appearance = int(n_appearance)
        self._transition = transition
        if not isinstance(transition, LSTM):
            raise NotImplementedError("the fused CUDA cell special-cases LSTM transitions (snt.LSTM, mnist_model.py:35)")
        self._n_hidden = self._transition.output_size[0]
        self._sample_presence = bool(discrete_steps)
        self._explore_eps = explore_eps
        self._debug = debug
        if canvas_init is not None:
            raise NotImplementedError("canvas_init: AIRModel always passes None (model.py:75)")

        self._spatial_transformer = SpatialTransformer(self._img_size, self._crop_size)
        self._inverse_transformer = SpatialTransformer(self._img_size, self._crop_size, inverse=True)
        self._transform_estimator = transform_estimator(self._n_transform_param)
        self._input_encoder = input_encoder()
        self._glimpse_encoder = glimpse_encoder()
        self._glimpse_decoder = glimpse_decoder(self._crop_size)
        self._what_distrib = ParametrisedGaussian(self._n_appearance, scale_offset=0.5)
        self._steps_predictor = steps_predictor()

        self.device = torch.device(device if device is not None else "cuda")
        self._output_std, self._output_multiplier, self._precision = output_std, output_multiplier, precision
        cfg = self.config
        self._spec = param_spec(cfg)
        self.params, self.views = _init_flat(self._spec, self.device, seed)
        self._transition.bind(self.views, "lstm")
        self._input_encoder.bind(self.views, "input_encoder")
        self._glimpse_encoder.bind(self.views, "glimpse_encoder")
        self._glimpse_decoder.bind(self.views, "glimpse_decoder")
        self._engines: Dict[Tuple, Engine] = {}

    # -- lowering of the module descriptors into the flat configuration of the fused path -----------------------
    @property
    def config(self) -> CellConfig:
        te, sp = self._transform_estimator, self._steps_predictor
        return CellConfig(
            H=self._img_size[0], W=self._img_size[1], h=self._crop_size[0], w=self._crop_size[1],
            na=self._n_appearance, nh=self._n_hidden,
            enc_hidden=tuple(self._input_encoder._n_hidden), glenc_hidden=tuple(self._glimpse_encoder._n_hidden),
            dec_hidden=tuple(self._glimpse_decoder._n_hidden), where_hidden=tuple(te._n_hidden),
            steps_hidden=tuple(sp._n_hidden),
            output_std=_scalar(self._output_std), output_multiplier=_scalar(self._output_multiplier),
            explore_eps=None if self._explore_eps is None else _scalar(self._explore_eps),
            scale_bias=_scalar(te._scale_bias), step_bias=_scalar(sp._steps_bias),
            what_scale_offset=self._what_distrib._scale_offset, forget_bias=self._transition.forget_bias,
            max_crop_size=te._max_crop_size, discrete_steps=self._sample_presence, precision=self._precision)

    def engine(self, B: int, T: int, precision=None, **kw) -> Engine:
        """The fused-path handle for (B, T); ``precision`` overrides the cell's arithmetic mode (the training step runs on
        the AIR_PREC_FP32 engine, which keeps the activations the backward pass needs)."""
        cfg = self.config
        if precision is not None:
            cfg.precision = precision
        key = (B, T, repr(cfg), tuple(sorted(kw.items())))
        if key not in self._engines:
            self._engines[key] = Engine(cfg, B, T, device=self.device, **kw)
        return self._engines[key]

    # -- RNNCore contract (cell.py:71-114) ----------------------------------------------------------------------
    @property
    def state_size(self):
        return [self._n_pix, self._n_pix, self._n_appearance, self._n_transform_param,
                self._transition.state_size, 1]

    @property
    def output_size(self):
        return [self._n_pix, int(np.prod(self._crop_size)), self._n_appearance, self._n_appearance,
                self._n_appearance, self._n_transform_param, self._n_transform_param, self._n_transform_param, 1, 1]

    @property
    def output_names(self):
        return 'canvas glimpse what what_loc what_scale where where_loc where_scale presence_prob presence'.split()

    def initial_state(self, img):
        B = img.shape[0]
        dev = img.device
        hidden_state = self._transition.initial_state(B, torch.float32, trainable=True)
        where_code = torch.zeros(B, self._n_transform_param, device=dev)
        what_code = torch.zeros(B, self._n_appearance, device=dev)
        flat_canvas = torch.zeros(B, self._n_pix, device=dev)
        flat_img = img.reshape(B, self._n_pix).to(torch.float32).contiguous()
        init_presence = torch.ones(B, 1, device=dev)
        return [flat_img, flat_canvas, what_code, where_code, hidden_state, init_presence]

    def draw_noise(self, B, T=None, generator=None):
        """eps_where ~ N(0,1) [.,B,4], eps_what ~ N(0,1) [.,B,na], u_pres ~ U[0,1) [.,B,1] (cell.py:133,147,156)."""
        lead = (B,) if T is None else (T, B)
        dev = self.device
        return (torch.randn(*lead, 4, device=dev, generator=generator),
                torch.randn(*lead, self._n_appearance, device=dev, generator=generator),
                torch.rand(*lead, 1, device=dev, generator=generator))

    def __call__(self, inpt, state, noise=None):
        """Input is unused (it only forces a number of steps, cell.py:117).  ``noise`` = (eps_where, eps_what, u_pres)
        for this step; drawn on the device when omitted."""
        img_flat, canvas_flat, what_code, where_code, hidden_state, presence = state
        B = img_flat.shape[0]
        eng = self.engine(B, 1, materialise_canvas=False, materialise_viz=False)
        if noise is None:
            noise = self.draw_noise(B)
        eps_where, eps_what, u_pres = (n.contiguous() for n in noise)
        canvas = canvas_flat.clone().contiguous()
        h, c = (s.clone().contiguous() for s in hidden_state)
        pres = presence.clone().contiguous()
        o = eng.cell_step(self.params, img_flat.contiguous(), canvas, h, c, pres, eps_where, eps_what, u_pres)
        output = [canvas, o["glimpse"], o["what"], o["what_loc"], o["what_scale"], o["where"], o["where_loc"],
                  o["where_scale"], o["presence_prob"], pres]
        new_state = [img_flat, canvas, o["what"], o["where"], (h, c), pres]
        return output, new_state
