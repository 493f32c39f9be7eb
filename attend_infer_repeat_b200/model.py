"""AIRModel -- mirror of the reference's model.py (model.py:15-376) over the fused CUDA path.

The reference builds a TF graph once and evaluates it with sess.run; here the "graph" is one fused C-ABI enqueue
(Engine.forward) whose caller-owned output buffers are exposed under the same attribute names
(canvas, glimpse, what, ..., rec_loss, kl_what, loss, ...).  ``forward()`` plays the role of sess.run: it re-evaluates
every attribute in place for a new batch.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from . import functional as F
from .cell import AIRCell
from .engine import make_prior
from ._lib import SCALAR_INDEX
from .ops import Loss
from .prior import NumStepsDistribution, geometric_prior


class _Normal:
    """Stand-in for tf.contrib.distributions.Normal(loc, scale) exposed as ``output_distrib`` (model.py:97)."""

    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale


def _get(d, k, default=None):
    if d is None:
        return default
    if isinstance(d, dict):
        return d.get(k, default)
    return getattr(d, k, default)


class AIRModel:
    """Generic AIR model"""

    def __init__(self, obs, nums, max_steps, glimpse_size,
                 n_appearance, transition, input_encoder, glimpse_encoder, glimpse_decoder, transform_estimator,
                 steps_predictor,
                 output_std=1., discrete_steps=True, output_multiplier=1.,
                 explore_eps=None, debug=False, **kwargs):
        """Arguments as in the reference (model.py:18-22).  ``obs`` [B,H,W] and ``nums`` [n_max+1,B,1] are CUDA tensors;
        **kwargs go to AIRCell (here: precision=, device=, seed=, materialise_canvas=)."""
        if not obs.is_cuda:
            raise _lib.AirError("obs must live on a CUDA device (there is no CPU path)")
        self.obs = obs.to(torch.float32).contiguous()
        self.nums = nums
        self.max_steps = int(max_steps)
        self.glimpse_size = tuple(glimpse_size)
        self.n_appearance = n_appearance
        self.output_std = output_std
        self.discrete_steps = discrete_steps
        self.explore_eps = explore_eps
        self.debug = debug
        self.output_multiplier = output_multiplier
        shape = list(self.obs.shape)
        self.batch_size = shape[0]
        self.img_size = shape[1:]
        self._materialise_canvas = bool(kwargs.pop("materialise_canvas", True))
        self._prior_struct = None
        self._train_cfg = None
        self.global_step = 0
        self._build(transition, input_encoder, glimpse_encoder, glimpse_decoder, transform_estimator,
                    steps_predictor, kwargs)

    def _build(self, transition, input_encoder, glimpse_encoder, glimpse_decoder, transform_estimator,
               steps_predictor, kwargs):
        kwargs.setdefault("device", self.obs.device)
        self.cell = AIRCell(self.img_size, self.glimpse_size, self.n_appearance, transition,
                            input_encoder, glimpse_encoder, glimpse_decoder, transform_estimator, steps_predictor,
                            canvas_init=None,
                            discrete_steps=self.discrete_steps,
                            explore_eps=self.explore_eps,
                            debug=self.debug,
                            output_std=self.output_std, output_multiplier=self.output_multiplier,
                            **kwargs)
        # `glimpse` (model.py:90, presence * sigmoid(decoded glimpse)) is a visualisation attribute: the reference's graph
        # evaluates it only when it is fetched, so it is computed on first access here (air_glimpse_viz), not by every pass
        self.engine = self.cell.engine(self.batch_size, self.max_steps, materialise_canvas=self._materialise_canvas,
                                       materialise_viz=False)
        self.forward()

    @property
    def params(self):
        return self.cell.params

    # ------------------------------------------------------------------------------------------------------
    def _current_prior(self):
        tc = self._train_cfg
        if tc is None:
            return make_prior(steps_success_prob=0.5, steps_prob_is_f64=False)
        nsp = tc["num_steps_prior"]
        if _get(nsp, "anneal") is not None:
            s = F.anneal_weight(_get(nsp, "init"), _get(nsp, "final"), _get(nsp, "anneal"), self.global_step,
                                _get(nsp, "steps"), _get(nsp, "hold_init", 0.), _get(nsp, "steps_div", 1.))
            is64 = True
        else:
            s, is64 = float(_get(nsp, "init")), False
        self.steps_prior_success_prob = s
        nvil_shift = nvil_scale = 0.0
        if getattr(self, "decay_rate", None) is not None:
            nvil_shift = float(self.imp_weight_moving_mean)
            nvil_scale = 1.0 / max(float(self.imp_weight_moving_var) ** 0.5, 1.0)
        return make_prior(tc["what_prior"], tc["where_scale_prior"], tc["where_shift_prior"], s, is64,
                          _get(nsp, "weight", 1.), _get(nsp, "analytic", True), bool(self.use_prior),
                          tc["use_reinforce"], nvil_shift, nvil_scale)

    def forward(self, obs=None, nums=None, noise=None):
        """Re-evaluate the model on a batch (the sess.run of the reference): T fused cell steps + ELBO terms.
        ``noise`` = (eps_where[T,B,4], eps_what[T,B,na], u_pres[T,B,1]); drawn on the device when omitted."""
        if obs is not None:
            assert tuple(obs.shape) == tuple(self.obs.shape)
            self.obs = obs.to(torch.float32).contiguous()
        if nums is not None:
            self.nums = nums
        T, B = self.max_steps, self.batch_size
        if noise is None:
            noise = self.cell.draw_noise(B, T)
        eps_where, eps_what, u_pres = (n.contiguous() for n in noise)
        self._last_noise = (eps_where, eps_what, u_pres)
        self._prior_struct = self._current_prior()
        if self.__dict__.get("_prior_on_device") and not torch.cuda.is_current_stream_capturing():
            self.engine.prior_table_device(self._prior_struct)   # graph mode: kernels read the step prior from device memory
        o = self.engine.forward(self.cell.params, self.obs, eps_where, eps_what, u_pres, self._prior_struct)

        # attributes named by AIRCell.output_names (model.py:86-87) and the post-processing of model.py:89-104
        for name in self.cell.output_names:
            if name != "glimpse":
                setattr(self, name, o[name])
        self.decoded_glimpse = o["glimpse"]
        self.__dict__.pop("glimpse", None)       # lazy: see __getattr__
        self.final_state = (o["final_h"], o["final_c"])
        if o["glimpse_viz"] is not None:
            self.glimpse = o["glimpse_viz"].view(T, B, *self.glimpse_size)
        if o["canvas"] is not None:
            self.canvas = o["canvas"].view(T, B, *self.img_size)
            self.final_canvas = self.canvas[-1]
            self.output_distrib = _Normal(self.final_canvas, self.output_std)
        self.num_steps_distrib = NumStepsDistribution(o["presence_prob"].view(T, B).t(), joint=o["num_steps_posterior"])
        self.num_step_per_sample = o["num_step_per_sample"]
        self.num_step = self.engine.scalar("num_step")
        if self.nums is not None:
            self.gt_num_steps = self.nums.sum(0).reshape(-1)
        if self._train_cfg is not None:
            self._expose_losses(o)
        return o

    # ------------------------------------------------------------------------------------------------------
    # attributes created by train_step / _prior_loss / _reinforce in the reference (model.py:143-372).  The reference's are graph
    # tensors, evaluated only when somebody fetches them; here they are materialised on first access after each forward
    # (the training step itself needs none of them: at B = 4096 building them eagerly cost 0.3 ms of host time per step,
    # and the step is host-bound -- tools/train_op_host_probe.py)
    _LAZY_LOSS_ATTRS = ("rec_loss_per_sample", "rec_loss", "kl_num_steps_per_sample", "kl_num_steps", "kl_what", "kl_where",
                        "prior_step_weight", "prior_loss", "prior_weight", "loss", "baseline_loss", "reinforce_loss",
                        "opt_loss", "num_step_accuracy")

    def __getattr__(self, name):
        # (only reached when normal lookup fails)
        if name == "glimpse" and self.__dict__.get("decoded_glimpse") is not None:
            g = F.glimpse_viz(self.decoded_glimpse, self.presence).view(self.max_steps, self.batch_size, *self.glimpse_size)
            self.__dict__["glimpse"] = g
            return g
        if name in AIRModel._LAZY_LOSS_ATTRS and self.__dict__.get("_loss_outputs") is not None:
            self._materialise_losses()
            if name in self.__dict__:
                return self.__dict__[name]
        raise AttributeError(name)

    def _expose_losses(self, o):
        """The part of the loss assembly the training step needs right away: the REINFORCE importance weight, the baseline
        (BaselineMLP forward on the engine) and the scalar block re-formed with the baseline mean.  Everything else is lazy."""
        eng = self.engine
        for k in AIRModel._LAZY_LOSS_ATTRS:
            self.__dict__.pop(k, None)
        self._loss_outputs = o
        tc = self._train_cfg
        self.reinforce_imp_weight = o["rec_loss_per_sample"]
        if tc["use_reinforce"]:
            if not _get(tc["num_steps_prior"], "analytic", True):
                sw = float(_get(tc["num_steps_prior"], "weight", 1.))
                self.reinforce_imp_weight = (o["rec_loss_per_sample"] + o["kl_num_steps_per_sample"] * sw +
                                             o["kl_what_per_sample"] + o["kl_where_per_sample"])
            if self.baseline_module is not None:
                b = self.baseline_module(self.obs, self.what, self.where, self.presence, self.final_state)
                self.baseline = b                                                 # [B,1]
                eng.elbo_scalars(b.reshape(-1), self._prior_struct)               # REINFORCE with the baseline mean

    def _materialise_losses(self):
        o, eng = self._loss_outputs, self.engine
        self.rec_loss_per_sample = o["rec_loss_per_sample"]
        self.rec_loss = eng.scalar("rec_loss")
        self.kl_num_steps_per_sample = o["kl_num_steps_per_sample"]
        self.kl_num_steps = eng.scalar("kl_num_steps")
        self.kl_what = eng.scalar("kl_what")
        self.kl_where = eng.scalar("kl_where")
        self.prior_step_weight = o["prior_step_weight"]
        tc = self._train_cfg
        sw = float(_get(tc["num_steps_prior"], "weight", 1.))
        self.prior_loss = Loss()
        self.prior_loss.add(self.kl_num_steps, self.kl_num_steps_per_sample, weight=sw)
        self.prior_loss.add(self.kl_what, o["kl_what_per_sample"])
        self.prior_loss.add(self.kl_where, o["kl_where_per_sample"])
        self.prior_weight = float(bool(self.use_prior))
        loss = Loss()
        loss._value, loss._per_sample = eng.scalar("loss"), o["loss_per_sample"]
        self.loss = loss
        if tc["use_reinforce"]:
            if self.baseline_module is not None:
                # [B] - [B,1] broadcasts to [B,B] in the reference (67 MB at B = 4096, 4.3 GB at 32768).
                # Nothing on the path needs the matrix -- the `importance_weight` property forms it on request.
                # .5 * mean over that broadcast of (iw_j - b_i)^2 = .5 (E iw^2 - 2 E iw E b + E b^2): four batch means the
                # scalars kernel has written (device arithmetic on 4 numbers, float64, no host round trip)
                sc = eng.out["scalars"].double()
                m_iw, m_iw2 = sc[SCALAR_INDEX["mean_iw"]], sc[SCALAR_INDEX["mean_iw2"]]
                m_b, m_b2 = sc[SCALAR_INDEX["mean_baseline"]], sc[SCALAR_INDEX["mean_baseline2"]]
                self.baseline_loss = (.5 * (m_iw2 - 2. * m_iw * m_b + m_b2)).float()
            self.reinforce_loss = eng.scalar("reinforce_loss")
        self.opt_loss = eng.scalar("opt_loss")
        if self.nums is not None:
            self.num_step_accuracy = (self.gt_num_steps == self.num_step_per_sample).to(torch.float32).mean()

    @property
    def importance_weight(self):
        """model.py:231: stop_gradient(iw) - baseline.  With a BaselineMLP that is [B] - [B,1] = a [B,B] matrix in the
        reference; it is formed here only when somebody reads the attribute."""
        iw = getattr(self, "reinforce_imp_weight", None)
        if iw is None:
            raise AttributeError("importance_weight exists after train_step() (model.py:224-231)")
        b = self.baseline if (self._train_cfg and self._train_cfg["use_reinforce"] and self.baseline_module is not None) \
            else None
        return iw if b is None else iw - b

    def train_step(self, learning_rate, l2_weight=0., what_prior=None, where_scale_prior=None,
                   where_shift_prior=None,
                   num_steps_prior=None, use_prior=True,
                   use_reinforce=True, baseline=None, decay_rate=None,
                   optimizer=None, opt_kwargs=dict(momentum=.9, centered=True), cuda_graph=None):
        """Creates the train step and the global_step (model.py:261-376).

        The returned ``train_op(obs=None, nums=None, noise=None)`` is the sess.run(train_step) of the reference: forward +
        ELBO on the batch, ``opt.compute_gradients(opt_loss)`` (air_backward), one all-reduce of the flat gradient buffer
        when torch.distributed is initialised (batch shards), and the centered-RMSProp update of the flat
        parameter buffer (air_rmsprop_step, TF semantics).  The engine is switched to training mode, in which every
        activation the backward pass needs is kept.  A baseline module (BaselineMLP) is trained by its own RMSProp at 10x the
        learning rate on .5 * mean((stop_gradient(iw) - baseline)^2) (model.py:253-259,362-367); decay_rate switches on
        the NVIL normalisation of the importance weight by its moving moments (model.py:232-239).

        ``cuda_graph`` (default: on for one process, ``AIR_TRAIN_GRAPH=0`` turns it off; opt-in under torch.distributed, where
        ``release_graphs()`` has to precede ``dist.destroy_process_group()``): like the reference, which builds its graph once and
        re-runs it, the step is enqueued eagerly twice and then captured as ONE CUDA graph (forward, BaselineMLP, both
        backward passes, the gradient all-reduces, both optimiser updates: ~110 launches on three streams) that every
        later ``train_op`` call replays.  Inputs are copied into the graph's static buffers; the annealed step prior is
        re-computed on the host every call and handed over through device memory (``Engine.prior_table_device``).  The
        NVIL variant (``decay_rate``) reads batch moments on the host every step and stays eager."""
        if num_steps_prior is None:
            raise ValueError("num_steps_prior is required (model.py:292 dereferences it)")
        if optimizer is not None:
            raise NotImplementedError("only the reference's default optimiser (tf.train.RMSPropOptimizer) is built")
        opt = dict(decay=.9, momentum=0., epsilon=1e-10, centered=False)     # tf.train.RMSPropOptimizer defaults
        opt.update(opt_kwargs or {})
        if not opt.pop("centered"):
            raise NotImplementedError("only centered RMSProp (the reference's opt_kwargs) is built")
        self._opt = opt
        self.l2_weight = l2_weight
        self.what_prior, self.where_scale_prior = what_prior, where_scale_prior
        self.where_shift_prior, self.num_steps_prior = where_shift_prior, num_steps_prior
        if not hasattr(self, 'baseline'):
            self.baseline = baseline
        if getattr(self, "baseline_module", None) is None:
            self.baseline_module = self.baseline if callable(self.baseline) else None
        self.use_prior = use_prior
        self.use_reinforce = use_reinforce
        self.learning_rate = learning_rate
        # NVIL normalisation of the importance weight (model.py:232-239; make_moving_average, ops.py:46-64)
        self.decay_rate = decay_rate
        self.imp_weight_moving_mean, self.imp_weight_moving_var = 0.0, 1.0
        self._train_cfg = dict(what_prior=what_prior, where_scale_prior=where_scale_prior,
                               where_shift_prior=where_shift_prior, num_steps_prior=num_steps_prior,
                               use_reinforce=use_reinforce)
        # the training engine (same arithmetic mode as the model, activations kept) replaces the inference engine
        self.engine = self.cell.engine(self.batch_size, self.max_steps, materialise_canvas=True, materialise_viz=False)
        if self.baseline_module is not None and hasattr(self.baseline_module, "attach") and use_reinforce:
            self.baseline_module.attach(self.engine)     # BaselineMLP on the engine (before the training workspace is sized)
        self.engine.train_enable(True)
        n = self.engine.n_params
        dev = self.obs.device
        if cuda_graph is None:
            import os
            import torch.distributed as dist
            sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            # under sharding the captured graph holds NCCL kernels, and NCCL cannot tear a communicator down while such a
            # graph is alive: release_graphs() must then precede dist.destroy_process_group() -- a contract the caller has
            # to know about, so it is opt-in there (cuda_graph=True or AIR_TRAIN_GRAPH=1)
            cuda_graph = os.environ.get("AIR_TRAIN_GRAPH", "0" if sharded else "1") != "0"
        self._graph_ok = bool(cuda_graph) and decay_rate is None and self.obs.is_cuda
        self._graphs, self._eager_steps, self._g_obs, self._g_noise = {}, 0, None, None
        self.graph_launches_per_step, self.graph_replays = 0, 0
        self._prior_on_device = False
        self.engine.prior_table_device(None)
        self._grad = torch.zeros(n, device=dev)
        self._slots = dict(mg=torch.zeros(n, device=dev), ms=torch.ones(n, device=dev), mom=torch.zeros(n, device=dev))
        self.forward()

        def train_op(obs=None, nums=None, noise=None):
            return self._run_train_step(obs, nums, noise)

        self._train_step = train_op
        return self._train_step, lambda: self.global_step

    def _run_train_step(self, obs=None, nums=None, noise=None):
        if self._graph_ok and self._eager_steps >= 2:
            return self._replay_train_step(obs, nums, noise)
        self._eager_steps += 1
        out = self._enqueue_train_step(obs, nums, noise)
        # UPDATE_OPS (model.py:357-360): the moving moments of the importance weight absorb this batch AFTER the gradient
        # was taken with their previous values (TF leaves the order of the read and the assign unspecified)
        if self.decay_rate is not None and self._train_cfg["use_reinforce"]:
            sc = out["scalars"].double().cpu()
            from ._lib import SCALAR_INDEX as SI
            m_iw, m_iw2 = float(sc[SI["mean_iw"]]), float(sc[SI["mean_iw2"]])
            m_b, m_b2 = float(sc[SI["mean_baseline"]]), float(sc[SI["mean_baseline2"]])
            # tf.nn.moments over the [B,B] broadcast of iw_j - baseline_i: mean = E iw - E b, var = Var iw + Var b
            mean, var = m_iw - m_b, max(m_iw2 - m_iw * m_iw, 0.0) + max(m_b2 - m_b * m_b, 0.0)
            d = float(self.decay_rate)
            self.imp_weight_moving_mean -= (1.0 - d) * (self.imp_weight_moving_mean - mean)
            self.imp_weight_moving_var -= (1.0 - d) * (self.imp_weight_moving_var - var)
        self.global_step += 1
        return out

    def release_graphs(self):
        """Destroy the captured training-step graphs (the next train_op captures again).  With torch.distributed initialised
        this MUST be called before dist.destroy_process_group(): the graphs hold NCCL kernels of the default group."""
        if self.__dict__.get("_graphs"):
            torch.cuda.synchronize()
            self._graphs.clear()

    def __del__(self):
        try:
            self.release_graphs()
        except Exception:
            pass

    def _replay_train_step(self, obs=None, nums=None, noise=None):
        """The captured step: copy the inputs into the graph's buffers, hand over this iteration's step prior, replay."""
        if self._g_obs is None:
            self._g_obs = self.obs.clone()
        if obs is not None:
            assert tuple(obs.shape) == tuple(self._g_obs.shape)
            self._g_obs.copy_(obs)
        self.obs = self._g_obs
        if nums is not None:
            self.nums = nums
        if noise is not None:
            if self._g_noise is None:
                self._g_noise = tuple(torch.empty(n.shape, device=self._g_obs.device, dtype=torch.float32) for n in noise)
            for d, n in zip(self._g_noise, noise):
                d.copy_(n)
        pr = self._current_prior()                 # host arithmetic: the annealed success probability of THIS iteration
        self._prior_on_device = True
        self.engine.prior_table_device(pr)
        key = (noise is None, bool(self.use_prior), float(self.learning_rate), float(self.l2_weight))
        ent = self._graphs.get(key)
        if ent is None:
            g = torch.cuda.CUDAGraph()
            keep_nums, self.nums = self.nums, None     # (the count of ground-truth objects is not part of the step)
            launches0 = self.engine.launch_count
            try:
                # thread_local: NCCL's watchdog thread polls its events while this thread captures
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    out = self._enqueue_train_step(None, None, self._g_noise if noise is not None else None)
            finally:
                self.nums = keep_nums
            ent = self._graphs[key] = (g, out, self._last_noise, self.__dict__.get("baseline"))
            self.graph_launches_per_step = self.engine.launch_count - launches0   # kernel nodes of one replay (library's own)
        g, out, self._last_noise, baseline = ent
        g.replay()
        self.graph_replays += 1
        # what forward() / _expose_losses leave behind on the Python side
        self._prior_struct = pr
        self.__dict__.pop("glimpse", None)
        for k in AIRModel._LAZY_LOSS_ATTRS:
            self.__dict__.pop(k, None)
        self._loss_outputs = out
        if baseline is not None:
            self.baseline = baseline
        if self.nums is not None:
            self.gt_num_steps = self.nums.sum(0).reshape(-1)
        self.global_step += 1
        return out

    def _enqueue_train_step(self, obs=None, nums=None, noise=None):
        import torch.distributed as dist
        from . import sharding
        out = self.forward(obs, nums, noise)
        eng, pr = self.engine, self._prior_struct
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        B = self.batch_size
        has_baseline = self._train_cfg["use_reinforce"] and self.baseline_module is not None
        if world > 1:
            # 16 floats: after this the scalar block holds the means of the WHOLE batch on every rank -- among them the
            # baseline mean and the importance-weight mean the two backward passes read from device memory below
            sharding.combine_scalars(out["scalars"], B, pr.steps_weight, bool(pr.use_prior), bool(pr.use_reinforce),
                                     nvil_shift=pr.nvil_shift, nvil_scale=pr.nvil_scale)
        eps_where, eps_what, _ = self._last_noise
        o = self._opt
        # The baseline's own gradient (model.py:253-259) needs only this step's forward results: it goes FIRST and its
        # weight-gradient GEMMs (the [B,3177]^T [B,256] product above all) are left on the engine's side streams, where they
        # overlap the cell's backward pass; air_backward's final join covers them.
        g_base = None
        if has_baseline:
            bm = self.baseline_module
            # the (global) importance-weight mean sits in the scalar block: read on the device by the gradient kernel
            tmean = out["scalars"][SCALAR_INDEX["mean_iw"]:SCALAR_INDEX["mean_iw"] + 1]
            g_base = bm.backward(self.reinforce_imp_weight, self.baseline, tmean, 1.0 / (world * B),
                                 defer_join=getattr(bm, "_engine", None) is eng)
        # baseline_mean = NaN: air_backward reads scalars[mean_baseline] on the device (no host synchronisation in the step)
        eng.backward(self.cell.params, self.obs, eps_where, eps_what, pr, self._grad,
                     baseline_mean=float("nan") if has_baseline else 0.0,
                     inv_batch=1.0 / (world * B), l2_weight=float(self.l2_weight) / world)
        # under sharding: both collectives are issued on NCCL's stream (async_op) and joined where their results are consumed
        base_work = grad_work = None
        if world > 1:
            if g_base is not None:
                base_work = dist.all_reduce(g_base, async_op=True)
            grad_work = dist.all_reduce(self._grad, async_op=True)   # the ONE data-path collective of the cell's parameters
        # the baseline's train step at 10x the learning rate (model.py:362-367, _make_baseline_train_step :253-259)
        if has_baseline:
            if base_work is not None:
                base_work.wait()
            eng.rmsprop_step(bm.params, g_base, bm.slots["mg"], bm.slots["ms"], bm.slots["mom"],
                             10.0 * float(self.learning_rate), o["decay"], o["momentum"], o["epsilon"])
        if grad_work is not None:
            grad_work.wait()
        eng.rmsprop_step(self.cell.params, self._grad, self._slots["mg"], self._slots["ms"], self._slots["mom"],
                         float(self.learning_rate), o["decay"], o["momentum"], o["epsilon"])
        return out

    def toggle_prior(self):
        """use_prior.assign(not use_prior) (model.py:306-308)."""
        self.use_prior = not self.use_prior
