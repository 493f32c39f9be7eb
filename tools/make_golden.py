#!/usr/bin/env python
"""Generate tests/golden/reference_*.npz by executing the REFERENCE'S OWN source files
(/root/reference/attend_infer_repeat/{prior,ops,model}.py) in this container.

TensorFlow 1.1 / Sonnet are not installable here, so the reference's TF primitives are served by tools/tf_stub.py
(an eager, torch-backed restatement of the documented semantics of ~50 TF ops); everything ABOVE the primitives --
geometric_prior, _cumprod, bernoulli_to_modified_geometric, masked_apply, tabular_kl, sample_from_tensor,
NumStepsDistribution, Loss, clip_preserve, AIRModel._anneal_weight, AIRModel._prior_loss, AIRModel._reinforce -- is
the reference's code, byte for byte.  The vectors pin the oracle (tests/test_oracle_golden.py) and, through it and
directly, the CUDA library (tests/test_gpu_golden.py).  /root/reference is only needed to RE-generate; the committed
.npz files travel.

    python tools/make_golden.py            # writes tests/golden/reference_prior.npz, reference_loss.npz
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


def main():
    tf_stub.install()
    # TF rebinding semantics for augmented assignment (`expr *= weight`, `importance_weight -= baseline`)
    saved = {}
    for name, fn in (("__iadd__", lambda a, b: a + b), ("__isub__", lambda a, b: a - b),
                     ("__imul__", lambda a, b: a * b), ("__itruediv__", lambda a, b: a / b)):
        saved[name] = getattr(torch.Tensor, name)
        setattr(torch.Tensor, name, fn)
    torch.Tensor.get_shape = lambda self: TensorShape(self.shape)
    sys.path.insert(0, REF)
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
    for name, fn in saved.items():
        setattr(torch.Tensor, name, fn)
    print("wrote", sorted(os.listdir(out_dir)), "cases:", case_id)


if __name__ == "__main__":
    main()
