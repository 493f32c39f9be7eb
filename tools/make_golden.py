#!/usr/bin/env python
"""Generate tests/golden/reference_*.npz by executing the REFERENCE'S OWN source files
(/root/reference/attend_infer_repeat/{prior,ops,model,cell,modules,neural,mnist_model}.py) in this container.

TensorFlow 1.1 / Sonnet are not installable here, so the reference's TF primitives are served by tools/tf_stub.py
(an eager, torch-backed restatement of the documented semantics of ~60 TF ops) and the ten Sonnet classes it uses by
tools/snt_stub.py; everything ABOVE them is the reference's code, byte for byte (one exception, spelt out in
load_reference_modules_py: the Python-2 integer division of modules.py:60):

  reference_prior.npz, reference_loss.npz   geometric_prior, _cumprod, bernoulli_to_modified_geometric, masked_apply,
                                            tabular_kl, sample_from_tensor, NumStepsDistribution, Loss, clip_preserve,
                                            AIRModel._anneal_weight, AIRModel._prior_loss, AIRModel._reinforce
  reference_cell_<case>.npz                 AIRonMNIST.__init__ / AIRModel.__init__ + _build (AIRCell.initial_state,
                                            AIRCell._build x T through dynamic_rnn, post-processing, model.py:83-104)
                                            with the Encoder / Decoder / StochasticTransformParam / StepsPredictor /
                                            ParametrisedGaussian / SpatialTransformer / MLP / Affine modules

The vectors pin the oracle (tests/test_oracle_golden.py) and, through it and directly, the CUDA library
(tests/test_gpu_golden.py, tests/test_gpu_zz_reference_vectors.py).  /root/reference is only needed to RE-generate; the committed .npz files travel.

    python tools/make_golden.py            # writes tests/golden/reference_{prior,loss,cell_script,cell_odd,cell_soft}.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/attend_infer_repeat"
sys.path.insert(0, HERE)

import tf_stub  # noqa: E402


class TensorShape(list):
    def as_list(self):
        return list(self)


class AttrDict(dict):
    """attrdict.AttrDict as the script uses it (multi_mnist.py:38-51): attribute access + `in`."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__


def np_(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def load_reference_modules_py():
    """modules.py:60 is `n_params = self._n_param / 2` followed by a slice: Python-2 integer division.  The file is
    executed from where it lies with that one operator spelt `//`; nothing else is touched."""
    import types
    path = os.path.join(REF, "modules.py")
    src = open(path).read()
    assert src.count("n_params = self._n_param / 2") == 1
    src = src.replace("n_params = self._n_param / 2", "n_params = self._n_param // 2")
    mod = types.ModuleType("modules")
    mod.__file__ = path
    sys.modules["modules"] = mod
    exec(compile(src, path, "exec"), mod.__dict__)


CELL_CASES = {
    # the script configuration (scripts/multi_mnist.py:55-59,82-94) through AIRonMNIST
    "script": dict(B=6, H=50, W=50, h=20, w=20, T=3, na=50, nh=256, enc=[256, 256], glenc=[256, 256], dec=[256, 256],
                   where=[256, 256], steps=[128, 64], output_std=.3, output_multiplier=.5, explore_eps=1e-3,
                   step_bias=.75, transform_var_bias=.5, discrete_steps=True, baseline=[256, 128], seed=11),
    # odd widths, a NON-square canvas and glimpse (x / y order), through AIRModel directly
    "odd": dict(B=5, H=9, W=14, h=4, w=6, T=4, na=7, nh=12, enc=[10], glenc=[9, 8], dec=[11], where=[13], steps=[6],
                output_std=.5, output_multiplier=1., explore_eps=None, step_bias=0., transform_var_bias=-1.,
                discrete_steps=True, seed=12),
    # all steps taken and weighted by the step probability (cell.py:150-151)
    "soft": dict(B=4, H=12, W=12, h=5, w=5, T=3, na=6, nh=10, enc=[8], glenc=[8], dec=[8], where=[8], steps=[5],
                 output_std=.3, output_multiplier=.5, explore_eps=1e-2, step_bias=.5, transform_var_bias=.5,
                 discrete_steps=False, seed=13),
}


def canonical_name(path, cfg):
    """Sonnet variable path of the stub (-> snt_stub.py) -> the flat-buffer name of include/air_b200.h."""
    parts = path.strip("/").split("/")
    if parts[0] == "lstm":
        return {"w_gates": "lstm.w", "b_gates": "lstm.b", "initial_state_0": "lstm.h0", "initial_state_1": "lstm.c0"}[parts[1]]
    if parts[0] == "BaselineMLP":                         # modules.py:125-143, baseline_hidden of mnist_model.py:18
        assert parts[1] == "MLP" and parts[2].startswith("linear"), path
        i = 0 if parts[2] == "linear" else int(parts[2].split("_")[1])
        return "baseline.%s.%s" % ("out" if i == len(cfg["baseline"]) else i, parts[3])
    assert parts[0] == "AIRCell", path
    if parts[1] == "ParametrisedGaussian":
        assert parts[2] == "linear"
        return "what." + parts[3]
    prefix, hidden, has_out = {"Encoder": ("input_encoder", cfg["enc"], False),
                               "Encoder_1": ("glimpse_encoder", cfg["glenc"], False),
                               "StochasticTransformParam": ("transform_estimator", cfg["where"], True),
                               "StepsPredictor": ("steps_predictor", cfg["steps"], True),
                               "Decoder": ("glimpse_decoder", cfg["dec"], True)}[parts[1]]
    assert parts[2] == "MLP" and parts[3].startswith("linear"), path
    i = 0 if parts[3] == "linear" else int(parts[3].split("_")[1])
    assert i < len(hidden) + int(has_out), path
    return "%s.%s.%s" % (prefix, "out" if i == len(hidden) else i, parts[4])


class GoldenOptimizer:
    """Stands where tf.train.RMSPropOptimizer stands in AIRModel.train_step (model.py:265,355-360): compute_gradients is
    tf.gradients(opt_loss, model_vars) -- autograd through the reference's own loss assembly -- and nothing is applied."""
    captured = None
    captured_baseline = None

    def __init__(self, learning_rate, **kwargs):
        pass

    def compute_gradients(self, loss, var_list=None):
        var_list = list(var_list)
        grads = torch.autograd.grad(loss, var_list, allow_unused=True, retain_graph=True)
        GoldenOptimizer.captured = list(zip(grads, var_list))
        return GoldenOptimizer.captured

    def apply_gradients(self, gvs, global_step=None):
        return None

    def minimize(self, loss, var_list=None):
        """model.py:253-259: the baseline's own train step; the gradient of baseline_loss w.r.t. the baseline variables."""
        var_list = list(var_list)
        GoldenOptimizer.captured_baseline = list(zip(torch.autograd.grad(loss, var_list, retain_graph=True), var_list))
        return None


TRAIN_CASES = {   # global_step, l2_weight, prior variants.  "script" goes through AIRonMNIST, whose BaselineMLP
    # (mnist_model.py:26) _reinforce builds and subtracts ([B] - [B,1] -> [B,B], model.py:224-231); AIRModel alone has none
    "script": dict(global_step=20000, l2_weight=0., analytic=True, shift_has_loc=True),
    "odd": dict(global_step=3000, l2_weight=1e-3, analytic=False, shift_has_loc=False),
    "soft": dict(global_step=60000, l2_weight=0., analytic=True, shift_has_loc=True),
}


def train_vectors(m, cfg, case, G, requested, requested_baseline):
    """AIRModel.train_step on the model the forward vectors came from (no new draws): the loss terms of
    model.py:319-343 and d opt_loss / d (every model variable) as compute_gradients sees it."""
    import snt_stub
    from tests.golden_recipe import golden_subset
    tc = TRAIN_CASES[case]
    tf_stub.train.get_or_create_global_step = staticmethod(lambda: torch.tensor(tc["global_step"], dtype=torch.int64))
    nsp = AttrDict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3,
                   analytic=tc["analytic"])
    shift = AttrDict(loc=0., scale=1.) if tc["shift_has_loc"] else AttrDict(scale=.8)
    GoldenOptimizer.captured = GoldenOptimizer.captured_baseline = None
    m.train_step(1e-5, l2_weight=tc["l2_weight"], what_prior=AttrDict(loc=0., scale=1.),
                 where_scale_prior=AttrDict(loc=0., scale=1.), where_shift_prior=shift, num_steps_prior=nsp,
                 use_prior=True, use_reinforce=True, optimizer=GoldenOptimizer, opt_kwargs=dict(momentum=.9, centered=True))
    name_of = {id(v): canonical_name(p, cfg) for p, v in snt_stub.VARIABLES.items()}
    assert len(GoldenOptimizer.captured) == len(requested), "compute_gradients did not see every model variable"
    def store(prefix, name, g):
        """A gradient tensor: its entries at golden_subset(name) (all of them when small), L2 norm, maximum magnitude."""
        flat = np_(g).reshape(-1).astype(np.float32)
        G[prefix + name] = flat[golden_subset(name, flat.size)]
        G[prefix + "stats:" + name] = np.array([np.sqrt((flat.astype(np.float64) ** 2).sum()), np.abs(flat).max()])

    for g, v in GoldenOptimizer.captured:
        store("grad:", name_of[id(v)], torch.zeros_like(v) if g is None else g)
    if requested_baseline:      # AIRonMNIST: the BaselineMLP of mnist_model.py:26 (built by _reinforce, model.py:224-229)
        assert len(GoldenOptimizer.captured_baseline) == len(requested_baseline)
        G["baseline_param_names"] = np.array(sorted(requested_baseline))
        G["baseline_param_shapes"] = np.array([requested_baseline[k] for k in sorted(requested_baseline)], dtype=np.int64)
        G["baseline_out"] = np_(m.baseline)
        G["train:baseline_loss"] = np_(m.baseline_loss)
        for g, v in GoldenOptimizer.captured_baseline:
            store("bgrad:", name_of[id(v)], g)
    G["train_cfg_json"] = np.array(__import__("json").dumps(tc))
    for name, val in (("loss", m.loss.value), ("loss_per_sample", m.loss.per_sample), ("opt_loss", m.opt_loss),
                      ("rec_loss", m.rec_loss), ("prior_loss", m.prior_loss.value),
                      ("prior_loss_per_sample", m.prior_loss.per_sample), ("kl_num_steps", m.kl_num_steps),
                      ("kl_what", m.kl_what), ("kl_where", m.kl_where), ("reinforce_loss", m.reinforce_loss),
                      ("importance_weight", m.importance_weight),
                      ("steps_prior_success_prob", m.steps_prior_success_prob)):
        G["train:" + name] = np_(val)


def cell_vectors(out_dir):
    """tests/golden/reference_cell_<case>.npz: the reference's AIRonMNIST / AIRModel / AIRCell source run on seeded
    weights (tests/golden_recipe.py), images and noise; every tensor model.py:86-104 exposes, the reconstruction loss of
    model.py:319-321, and the losses / gradients of AIRModel.train_step (train_vectors)."""
    import functools
    import json
    import snt_stub
    sys.path.insert(0, ROOT)
    from tests.golden_recipe import golden_tensor
    import mnist_model as ref_mnist     # noqa: E402  reference source
    import model as ref_model           # noqa: E402
    import modules as ref_modules       # noqa: E402  (the patched load above)
    import sonnet as snt                # noqa: E402  (tools/snt_stub.py)

    def build(case, cfg):
        """Run the reference's model constructor (-> AIRCell x T) on the inputs seeded by cfg['seed']."""
        snt_stub.reset()
        requested, requested_baseline = {}, {}

        def source(path, shape):
            name = canonical_name(path, cfg)
            shape2 = (1, shape[0]) if len(shape) == 1 else shape
            (requested_baseline if name.startswith("baseline.") else requested)[name] = shape2
            return torch.from_numpy(golden_tensor(name, shape2, cfg["seed"])).reshape(shape).requires_grad_(True)

        snt_stub.VARIABLE_SOURCE = source
        B, H, W, T = cfg["B"], cfg["H"], cfg["W"], cfg["T"]
        rs = np.random.RandomState(cfg["seed"])
        inp = {"img": (rs.rand(B, H, W) * (rs.rand(B, H, W) > 0.7)).astype(np.float32),
               "eps_where": rs.standard_normal((T, B, 4)).astype(np.float32),
               "eps_what": rs.standard_normal((T, B, cfg["na"])).astype(np.float32),
               "u_pres": rs.rand(T, B, 1).astype(np.float32)}
        # draws in the order cell.py makes them within a step: where (:133), presence (:147, discrete only), what (:156)
        tf_stub.NOISE_NORMAL[:] = [torch.from_numpy(inp[k][t]) for t in range(T) for k in ("eps_where", "eps_what")]
        tf_stub.NOISE_UNIFORM[:] = [torch.from_numpy(inp["u_pres"][t]) for t in range(T)] if cfg["discrete_steps"] else []
        obs, nums_t = torch.from_numpy(inp["img"]), torch.zeros(3, B, 1)
        if case == "script":
            m = ref_mnist.AIRonMNIST(obs, nums_t, glimpse_size=(cfg["h"], cfg["w"]), max_steps=T,
                                     inpt_encoder_hidden=cfg["enc"], glimpse_encoder_hidden=cfg["glenc"],
                                     glimpse_decoder_hidden=cfg["dec"], transform_estimator_hidden=cfg["where"],
                                     steps_pred_hidden=cfg["steps"], baseline_hidden=cfg["baseline"],
                                     transform_var_bias=cfg["transform_var_bias"], step_bias=cfg["step_bias"],
                                     output_multiplier=cfg["output_multiplier"], discrete_steps=cfg["discrete_steps"],
                                     explore_eps=cfg["explore_eps"])
            assert abs(m.output_std - cfg["output_std"]) < 1e-12 and m.n_appearance == cfg["na"]
        else:
            P = functools.partial
            m = ref_model.AIRModel(obs, nums_t, T, (cfg["h"], cfg["w"]), cfg["na"], snt.LSTM(cfg["nh"]),
                                   P(ref_modules.Encoder, cfg["enc"]), P(ref_modules.Encoder, cfg["glenc"]),
                                   P(ref_modules.Decoder, cfg["dec"]),
                                   P(ref_modules.StochasticTransformParam, cfg["where"],
                                     scale_bias=cfg["transform_var_bias"]),
                                   P(ref_modules.StepsPredictor, cfg["steps"], cfg["step_bias"]),
                                   output_std=cfg["output_std"], discrete_steps=cfg["discrete_steps"],
                                   output_multiplier=cfg["output_multiplier"], explore_eps=cfg["explore_eps"])
        assert not tf_stub.NOISE_NORMAL and not tf_stub.NOISE_UNIFORM, "the cell did not consume every queued draw"
        return m, inp, requested, requested_baseline

    def well_conditioned(m, cfg, u):
        """Seeds are chosen so that the vectors are not dominated by fp32 noise: every glimpse that is painted has
        |sx|, |sy| >= 0.2 (1 / s amplifies rounding in the inverse transformer), no step draw is within 1e-3 of a tie, and
        the discrete cases hold both taken and skipped steps."""
        wh, pres = np_(m.where), np_(m.presence)[..., 0]
        painted = pres > 0
        if cfg["discrete_steps"]:
            if np.abs(u[..., 0] - np_(m.presence_prob)[..., 0]).min() < 1e-3 or painted.sum() < 3 or painted.all():
                return False
        s = np.minimum(np.abs(wh[..., 0]), np.abs(wh[..., 2]))[painted]
        return s.size > 0 and s.min() >= 0.2

    for case, base_cfg in CELL_CASES.items():
        for seed in range(base_cfg["seed"], base_cfg["seed"] + 5000, 10):
            cfg = dict(base_cfg, seed=seed)
            m, inp, requested, requested_baseline = build(case, cfg)
            if well_conditioned(m, cfg, inp["u_pres"]):
                break
        else:
            raise RuntimeError("no well-conditioned seed found for case " + case)
        # model.py:319-321
        rec_ps = tf_stub.reduce_sum(-m.output_distrib.log_prob(m.obs), axis=(1, 2))
        G = dict(inp)
        G.update({"cfg_json": np.array(json.dumps(cfg)), "param_names": np.array(sorted(requested)),
                  "param_shapes": np.array([requested[k] for k in sorted(requested)], dtype=np.int64)})
        for name in ("what", "what_loc", "what_scale", "where", "where_loc", "where_scale", "presence_prob", "presence",
                     "canvas", "glimpse", "final_canvas", "num_step_per_sample"):
            G[name] = np_(getattr(m, name))
        G["final_h"], G["final_c"] = np_(m.final_state[0]), np_(m.final_state[1])
        G["num_steps_posterior"] = np_(m.num_steps_distrib.prob())
        G["rec_loss_per_sample"] = np_(rec_ps)
        train_vectors(m, cfg, case, G, requested, requested_baseline)
        np.savez_compressed(os.path.join(out_dir, "reference_cell_%s.npz" % case), **G)
        wh = G["where"]
        print("cell case", case, "seed", cfg["seed"], {k: tuple(G[k].shape) for k in ("canvas", "glimpse", "what")},
              "params", len(requested), "num_step", float(m.num_step.detach()),
              "min |s| painted", float(np.minimum(np.abs(wh[..., 0]), np.abs(wh[..., 2]))[G["presence"][..., 0] > 0].min()))


def main():
    tf_stub.install()
    # TF rebinding semantics for augmented assignment (`expr *= weight`, `importance_weight -= baseline`)
    saved = {}
    for name, fn in (("__iadd__", lambda a, b: a + b), ("__isub__", lambda a, b: a - b),
                     ("__imul__", lambda a, b: a * b), ("__itruediv__", lambda a, b: a / b)):
        saved[name] = getattr(torch.Tensor, name)
        setattr(torch.Tensor, name, fn)
    torch.Tensor.get_shape = lambda self: TensorShape(self.shape)
    torch.Tensor.assign = lambda self, value: value          # tf.Variable.assign builds an op; nothing runs it here
    sys.path.insert(0, REF)
    load_reference_modules_py()
    import model as ref_model      # noqa: E402  reference source
    import ops as ref_ops          # noqa: E402
    import prior as ref_prior      # noqa: E402

    rng = np.random.RandomState(20171017)
    t32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32))
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # ------------------------------------------------------------------ prior.py ---------------------------------
    g = {}
    g["geom_075_10"] = np_(ref_prior.geometric_prior(.75, 10))
    g["geom_0005_3"] = np_(ref_prior.geometric_prior(.005, 3))
    probs = rng.rand(257, 3).astype(np.float32)
    probs[0], probs[1], probs[2], probs[3] = 0., 1., [1., 1., 0.], [.5, 0., 0.]
    g["b2mg_in"] = probs
    g["b2mg_out"] = np_(ref_prior.bernoulli_to_modified_geometric(t32(probs)))
    probs5 = rng.rand(4, 6, 5).astype(np.float32)
    g["b2mg5_in"] = probs5
    g["b2mg5_out"] = np_(ref_prior.bernoulli_to_modified_geometric(t32(probs5)))
    p = rng.rand(64, 4).astype(np.float32)
    p /= p.sum(1, keepdims=True)
    p[0] = [0., .25, .25, .5]
    p[1] = [0., 1., 0., 0.]
    q = rng.rand(4).astype(np.float32)
    q /= q.sum()
    g["tkl_p"], g["tkl_q"] = p, q
    g["tkl_out"] = np_(ref_prior.tabular_kl(t32(p), t32(q), 0.))
    dist = ref_prior.NumStepsDistribution(t32(probs))
    n = rng.randint(0, 4, size=257).astype(np.float32)
    g["nsd_samples"] = n
    g["nsd_prob"] = np_(dist.prob(t32(n)))
    g["nsd_log_prob"] = np_(dist.log_prob(t32(n)))
    g["nsd_joint"] = np_(dist.prob())
    x = t32([1e-40, 0.5, 2.0]).requires_grad_(True)
    y = ref_ops.clip_preserve(x, 1e-32, 1.0)
    y.sum().backward()
    g["clip_out"], g["clip_grad"] = np_(y), np_(x.grad)
    np.savez(os.path.join(out_dir, "reference_prior.npz"), **g)

    # ------------------------------------------------------------------ model.py: schedules + losses ------------
    L = {}
    steps = np.array([0, 500, 1000, 1500, 5000, 20000, 60000, 101000, 500000], dtype=np.int64)
    L["anneal_steps"] = steps
    L["anneal_exp"] = np.array([float(ref_model.AIRModel._anneal_weight(1. - 1e-15, 1e-7, 'exp', int(s), 1e5, 1e3, 1e4))
                                for s in steps])
    L["anneal_linear"] = np.array([float(ref_model.AIRModel._anneal_weight(.9, .1, 'linear', int(s), 1e5, 1e3, 1.))
                                   for s in steps])

    T, B, na = 3, 32, 50
    case_id = 0
    for analytic in (True, False):
        for shift_has_loc in (True, False):
            for global_step, anneal in ((0, 'exp'), (20000, 'exp'), (200000, 'exp'), (0, None)):
                m = ref_model.AIRModel.__new__(ref_model.AIRModel)
                m.max_steps = T
                pp = rng.rand(T, B, 1).astype(np.float32) * 0.98 + 0.01
                pres = (rng.rand(T, B, 1) < pp).astype(np.float32).cumprod(0).astype(np.float32)
                m.presence_prob, m.presence = t32(pp), t32(pres)
                m.what_loc, m.what_scale = t32(rng.randn(T, B, na)), t32(rng.rand(T, B, na) * 2 + 0.05)
                m.where_loc, m.where_scale = t32(rng.randn(T, B, 4)), t32(rng.rand(T, B, 4) * 2 + 0.05)
                m.num_steps_distrib = ref_prior.NumStepsDistribution(tf_stub.transpose(tf_stub.squeeze(m.presence_prob)))
                m.num_step_per_sample = tf_stub.to_float(tf_stub.squeeze(tf_stub.reduce_sum(m.presence, 0)))
                nsp = AttrDict(anneal=anneal, init=(1. - 1e-15) if anneal else 0.3, final=1e-7, steps_div=1e4, steps=1e5,
                               hold_init=1e3, analytic=analytic)
                if case_id % 3 == 1:
                    nsp['weight'] = 0.5
                what_prior = AttrDict(loc=0., scale=1.) if case_id % 2 == 0 else AttrDict(loc=.2, scale=1.5)
                scale_prior = AttrDict(loc=0., scale=1.) if case_id % 2 == 0 else AttrDict(loc=.5, scale=.7)
                shift_prior = AttrDict(loc=0., scale=1.) if shift_has_loc else AttrDict(scale=.8)
                pl = m._prior_loss(what_prior, scale_prior, shift_prior, nsp, global_step)
                rec = t32(rng.rand(B) * 800 + 50)
                base = t32(rng.randn(B, 1) * 100 + 400)
                m.baseline = None
                r_nob = m._reinforce(rec + (0 if analytic else pl.per_sample), None)
                iw_nob = np_(m.importance_weight)
                m.baseline = base
                r_b = m._reinforce(rec + (0 if analytic else pl.per_sample), None)
                k = f"c{case_id}_"
                L[k + "cfg"] = np.array([int(analytic), int(shift_has_loc), global_step, 1 if anneal else 0,
                                         nsp.get('weight', 1.0), what_prior.loc, what_prior.scale, scale_prior.loc,
                                         scale_prior.scale, shift_prior.get('loc', 0.0), shift_prior.scale,
                                         nsp.init], dtype=np.float64)
                for name, val in (("presence_prob", pp), ("presence", pres), ("what_loc", m.what_loc),
                                  ("what_scale", m.what_scale), ("where_loc", m.where_loc),
                                  ("where_scale", m.where_scale), ("rec", rec), ("baseline", base),
                                  ("success_prob", m.steps_prior_success_prob), ("posterior", m.num_steps_distrib.prob()),
                                  ("step_weight", m.prior_step_weight), ("kl_num_steps_ps", m.kl_num_steps_per_sample),
                                  ("kl_num_steps", m.kl_num_steps), ("kl_what", m.kl_what), ("kl_where", m.kl_where),
                                  ("prior_value", pl.value), ("prior_per_sample", pl.per_sample),
                                  ("reinforce_nobaseline", r_nob), ("imp_weight_nobaseline", iw_nob),
                                  ("reinforce_baseline", r_b), ("imp_weight_baseline", m.importance_weight),
                                  ("log_prob", m.num_steps_distrib.log_prob(m.num_step_per_sample))):
                    L[k + name] = np_(val)
                case_id += 1
    L["n_cases"] = np.array(case_id)
    np.savez_compressed(os.path.join(out_dir, "reference_loss.npz"), **L)
    cell_vectors(out_dir)
    for name, fn in saved.items():
        setattr(torch.Tensor, name, fn)
    print("wrote", sorted(os.listdir(out_dir)), "cases:", case_id)


if __name__ == "__main__":
    main()
