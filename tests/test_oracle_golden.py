"""The oracle against golden vectors produced by the REFERENCE'S OWN prior.py / ops.py / model.py source, executed in
this container over a torch-backed stand-in for the TF primitives (tools/make_golden.py, tools/tf_stub.py).
This pins the step-count algebra, the anneal schedule, _prior_loss (all KL terms, analytic and sampled weights,
shift prior with / without loc, prior weights) and _reinforce (with the [B]-[B,1] -> [B,B] broadcast)."""
import os

import numpy as np
import pytest
import torch

from oracle import air_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P = np.load(os.path.join(GOLD, "reference_prior.npz"))
L = np.load(os.path.join(GOLD, "reference_loss.npz"))
T32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)


def test_geometric_prior():
    assert np.allclose(O.geometric_prior(.75, 10).numpy(), P["geom_075_10"], rtol=1e-6, atol=0)
    assert np.allclose(O.geometric_prior(.005, 3).numpy(), P["geom_0005_3"], rtol=1e-6, atol=0)
    assert np.allclose(P["geom_075_10"], .25 * .75 ** np.arange(11), atol=1e-6)      # test/prior_test.py:15-24


def test_bernoulli_to_modified_geometric_bitwise():
    assert np.array_equal(O.bernoulli_to_modified_geometric(T32(P["b2mg_in"])).numpy(), P["b2mg_out"])
    assert np.array_equal(O.bernoulli_to_modified_geometric(T32(P["b2mg5_in"])).numpy(), P["b2mg5_out"])
    assert np.array_equal(P["b2mg_out"][0], [1, 0, 0, 0]) and np.array_equal(P["b2mg_out"][1], [0, 0, 0, 1])


def test_tabular_kl():
    out = O.tabular_kl(T32(P["tkl_p"]), T32(P["tkl_q"])).numpy()
    assert np.allclose(out, P["tkl_out"], rtol=1e-6, atol=1e-9)
    assert out[0, 0] == 0.0 and P["tkl_out"][0, 0] == 0.0


def test_num_steps_distribution():
    joint = O.bernoulli_to_modified_geometric(T32(P["b2mg_in"]))
    assert np.array_equal(joint.numpy(), P["nsd_joint"])
    n = T32(P["nsd_samples"])
    assert np.array_equal(O.num_steps_prob(joint, n).numpy(), P["nsd_prob"])
    assert np.allclose(O.num_steps_log_prob(joint, n).numpy(), P["nsd_log_prob"], rtol=1e-6, atol=1e-7)


def test_clip_preserve():
    x = T32([1e-40, 0.5, 2.0]).requires_grad_(True)
    y = O.clip_preserve(x, 1e-32, 1.0)
    y.sum().backward()
    assert np.array_equal(y.detach().numpy(), P["clip_out"]) and np.array_equal(x.grad.numpy(), P["clip_grad"])


def test_anneal_weight():
    for i, s in enumerate(L["anneal_steps"]):
        a = float(O.anneal_weight(1. - 1e-15, 1e-7, "exp", int(s), 1e5, 1e3, 1e4))
        assert a == pytest.approx(float(L["anneal_exp"][i]), rel=1e-13), (s, a)
        b = float(O.anneal_weight(.9, .1, "linear", int(s), 1e5, 1e3, 1.))
        assert b == pytest.approx(float(L["anneal_linear"][i]), rel=1e-13)


def _case(i):
    k = f"c{i}_"
    cfg = L[k + "cfg"]
    analytic, has_loc, gstep, anneal, weight = bool(cfg[0]), bool(cfg[1]), int(cfg[2]), bool(cfg[3]), float(cfg[4])
    pc = O.PriorConfig(what_loc=float(cfg[5]), what_scale=float(cfg[6]), where_scale_loc=float(cfg[7]),
                       where_scale_scale=float(cfg[8]), where_shift_loc=float(cfg[9]) if has_loc else None,
                       where_shift_scale=float(cfg[10]), steps_anneal="exp" if anneal else None,
                       steps_init=float(cfg[11]), steps_weight=weight, analytic=analytic)
    g = {n[len(k):]: L[n] for n in L.files if n.startswith(k)}
    return pc, gstep, g


@pytest.mark.parametrize("i", range(int(L["n_cases"])))
def test_prior_loss_and_reinforce(i):
    pc, gstep, g = _case(i)
    T, B = g["presence"].shape[:2]
    cfg = O.AirConfig(T=T)
    outs = {k: T32(g[k]) for k in ("presence_prob", "presence", "what_loc", "what_scale", "where_loc", "where_scale")}
    post = dict(num_steps_posterior=O.bernoulli_to_modified_geometric(outs["presence_prob"].reshape(T, B).t()),
                num_step_per_sample=outs["presence"].sum(0).reshape(B))
    assert np.array_equal(post["num_steps_posterior"].numpy(), g["posterior"])
    pl, terms = O.prior_loss(cfg, pc, outs, post, gstep)
    # annealed: float64 scalar; fixed: the python float itself (it only becomes float32 inside geometric_prior)
    assert float(terms["steps_prior_success_prob"]) == pytest.approx(
        float(g["success_prob"]), rel=1e-12 if pc.steps_anneal else 1e-7)
    close = lambda a, b, rt=2e-6: np.allclose(np.asarray(a), b, rtol=rt, atol=1e-6)
    assert close(terms["prior_step_weight"].numpy(), g["step_weight"])
    assert close(terms["kl_num_steps_per_sample"].numpy(), g["kl_num_steps_ps"])
    assert close(float(terms["kl_num_steps"]), g["kl_num_steps"])
    assert close(float(terms["kl_what"]), g["kl_what"]) and close(float(terms["kl_where"]), g["kl_where"])
    assert close(float(pl.value), g["prior_value"]) and close(pl.per_sample.numpy(), g["prior_per_sample"])
    rec = T32(g["rec"])
    iw = rec if pc.analytic else rec + pl.per_sample
    r0, iw0, lp, _ = O.reinforce(post["num_steps_posterior"], post["num_step_per_sample"], iw, None)
    assert close(float(r0), g["reinforce_nobaseline"], 1e-5) and close(iw0.numpy(), g["imp_weight_nobaseline"])
    assert close(lp.numpy(), g["log_prob"])
    r1, iw1, _, _ = O.reinforce(post["num_steps_posterior"], post["num_step_per_sample"], iw, T32(g["baseline"]))
    assert tuple(iw1.shape) == (B, B) == g["imp_weight_baseline"].shape
    assert close(iw1.numpy(), g["imp_weight_baseline"]) and close(float(r1), g["reinforce_baseline"], 1e-5)


# ---- the cell and the unroll: golden vectors from the reference's own cell.py / modules.py / neural.py / model.py /
# mnist_model.py, executed over tools/snt_stub.py + tools/tf_stub.py (tools/make_golden.py: cell_vectors) -------------------
CELL_KEYS = ("what", "what_loc", "what_scale", "where", "where_loc", "where_scale", "presence_prob")


@pytest.mark.parametrize("case", ["script", "odd", "soft"])
def test_cell_and_unroll_match_the_reference_source(case):
    """AIRCell._build x T through dynamic_rnn + the post-processing of model.py:83-104 + the reconstruction loss of
    model.py:319-321, as the reference's source computes them on seeded weights / images / draws: the script
    configuration through AIRonMNIST, a non-square canvas + glimpse with odd widths, and the non-discrete mode.  Pins the
    oracle's glue -- (sx, tx, sy, ty) order, biases, explore-eps mix, presence product, LSTM wiring, canvas accumulation."""
    from tests import util as U
    ocfg, params, img, noise, ref = U.load_cell_golden(case)
    with torch.no_grad():
        res = O.forward(ocfg, O.PriorConfig(), params, img, *noise, global_step=0)
    for k in CELL_KEYS:
        U.assert_close(res["outs"][k].reshape(ref[k].shape), ref[k], atol=2e-5, rtol=1e-5, name=k)
    if ocfg.discrete_steps:
        assert torch.equal(res["outs"]["presence"].reshape(ref["presence"].shape), ref["presence"])
        assert torch.equal(res["num_step_per_sample"].reshape(-1), ref["num_step_per_sample"].reshape(-1))
    else:
        U.assert_close(res["outs"]["presence"].reshape(ref["presence"].shape), ref["presence"], atol=1e-6, name="presence")
        U.assert_close(res["num_step_per_sample"], ref["num_step_per_sample"], atol=1e-5, name="num_step_per_sample")
    for k in ("canvas", "glimpse", "final_canvas", "final_h", "final_c"):
        U.assert_close(res[k].reshape(ref[k].shape), ref[k], atol=5e-5, rtol=1e-5, name=k)
    U.assert_close(res["num_steps_posterior"], ref["num_steps_posterior"], atol=1e-6, rtol=1e-5, name="q(n)")
    U.assert_close(res["rec_loss_per_sample"], ref["rec_loss_per_sample"], atol=0, rtol=1e-5, name="rec_loss_per_sample")
    # the vectors are not degenerate: something was painted, and (discrete cases) both outcomes of the step draw occur
    assert float(ref["canvas"].abs().max()) > 0.1
    if ocfg.discrete_steps:
        assert 0.0 < float(ref["presence"].mean()) < 1.0


@pytest.mark.parametrize("case", ["script", "odd", "soft"])
def test_train_step_losses_and_gradients_match_the_reference_source(case):
    """AIRModel.train_step (model.py:261-376) executed from the reference's source on the model of the forward vectors, with
    compute_gradients = autograd through its own loss assembly: every loss term, and d opt_loss / d (every model variable)
    against autograd on the oracle -- which pins the gradient oracle of the backward kernels (stop_gradient placement,
    REINFORCE, the straight-through clip, L2 on the 2-D variables only).  The script case goes through AIRonMNIST, so the
    BaselineMLP's input order (modules.py:125-143), its [B]-[B,1] broadcast and its own gradient are covered too."""
    from tests import util as U
    ocfg, params, img, noise, ref = U.load_cell_golden(case)
    pc, gstep, l2, g = U.load_train_golden(case)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    baseline, pb = None, None
    if "baseline_out" in g.files:
        seed = __import__("json").loads(str(g["cfg_json"]))["seed"]
        pb = {k: v.clone().requires_grad_(True) for k, v in U.golden_baseline_params(g, seed).items()}
        T, B = ocfg.T, img.shape[0]
        b_in = {k: ref[k].reshape(T, B, -1) for k in ("what", "where", "presence")}
        baseline = O.baseline_mlp(pb, 2, img, b_in["what"], b_in["where"], b_in["presence"], ref["final_h"], ref["final_c"])
        U.assert_close(baseline.detach(), T32(g["baseline_out"]), atol=1e-5, rtol=1e-5, name="baseline")
    res = O.forward(ocfg, pc, p, img, *noise, global_step=gstep,
                    baseline=None if baseline is None else baseline.detach())
    for k in ("loss", "rec_loss", "prior_loss", "kl_num_steps", "kl_what", "kl_where", "reinforce_loss"):
        U.assert_close(res[k].detach().float(), T32(g["train:" + k]), atol=1e-5, rtol=2e-5, name=k)
    U.assert_close(res["loss_per_sample"].detach(), T32(g["train:loss_per_sample"]), atol=1e-4, rtol=1e-5, name="loss_ps")
    U.assert_close(res["prior_loss_per_sample"].detach(), T32(g["train:prior_loss_per_sample"]), atol=1e-5, rtol=1e-5,
                   name="prior_ps")
    U.assert_close(res["importance_weight"].detach().reshape(-1), T32(g["train:importance_weight"]).reshape(-1),
                   atol=1e-3, rtol=1e-5, name="importance_weight")
    assert float(res["steps_prior_success_prob"]) == pytest.approx(float(g["train:steps_prior_success_prob"]), rel=1e-12)
    loss = res["opt_loss"]
    if l2 > 0:      # model.py:345-350: the 2-D variables (weights and the [1, nh] trainable initial state)
        loss = loss + l2 * sum(0.5 * (v ** 2).sum() for k, v in p.items() if k.endswith(".w") or k in ("lstm.h0", "lstm.c0"))
    U.assert_close(loss.detach().float(), T32(g["train:opt_loss"]), atol=1e-5, rtol=2e-5, name="opt_loss")
    loss.backward()
    for name, _ in O.param_spec(ocfg):
        grad = p[name].grad if p[name].grad is not None else torch.zeros_like(p[name])
        U.compare_with_golden_gradient("grad:", name, grad, g)
    if baseline is not None:
        # model.py:253-259 with the [B] - [B,1] broadcast
        target = res["rec_loss_per_sample"].detach()
        b_loss = 0.5 * ((target[None, :] - baseline) ** 2).mean()
        U.assert_close(b_loss.detach(), T32(g["train:baseline_loss"]), atol=0, rtol=2e-5, name="baseline_loss")
        b_loss.backward()
        for name in pb:
            U.compare_with_golden_gradient("bgrad:", name, pb[name].grad, g)
