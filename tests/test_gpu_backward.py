"""Gradient parity of the CUDA backward pass (air_backward, through the C ABI) against autograd on the CPU oracle
(tf.gradients(opt_loss, model_vars) on the reference graph, model.py:335-360), and of the centered RMSProp kernel against
the oracle's restatement of ApplyCenteredRMSProp.

Tolerance (written here): per parameter tensor, |g_cuda - g_oracle| <= 2e-4 * max|g_oracle| of that tensor + 1e-7 --
both sides are fp32 sums over up to B*T*P terms in different orders."""
import pytest
import torch

import attend_infer_repeat_b200 as air
from oracle import air_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


def oracle_grads(ocfg, pc, params, img, noise, global_step, baseline=None, l2_weight=0.0, dtype=torch.float32):
    p = {k: v.clone().to(dtype).requires_grad_(True) for k, v in params.items()}
    img, noise = img.to(dtype), tuple(n.to(dtype) for n in noise)
    res = O.forward(ocfg, pc, p, img, *noise, global_step=global_step, baseline=baseline)
    loss = res["opt_loss"]
    if l2_weight > 0:   # model.py:345-350: 2-D variables only (weights and the [1,nh] trainable initial state)
        loss = loss + l2_weight * sum(0.5 * (v ** 2).sum() for k, v in p.items()
                                      if k.endswith(".w") or k in ("lstm.h0", "lstm.c0"))
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).float() for k, v in p.items()}
    return res, grads


def cuda_grads(ocfg, pc, params, img, noise, global_step, baseline=None, l2_weight=0.0, inv_batch=0.0,
               precision=air.AIR_PREC_FP32):
    B = img.shape[0]
    dev = "cuda"
    eng = air.Engine(U.cell_cfg(ocfg, precision), B, ocfg.T, device=dev)
    eng.train_enable(True)
    flat = O.flatten_params(ocfg, params).to(dev)
    ew, ea, u = (n.to(dev).contiguous() for n in noise)
    img_d = img.to(dev).contiguous()
    pr = U.prior_struct(pc, global_step)
    bl = None if baseline is None else baseline.reshape(-1).to(dev).contiguous()
    out = eng.forward(flat, img_d, ew, ea, u, pr, bl)
    bmean = 0.0 if baseline is None else float(baseline.mean())
    g = eng.backward(flat, img_d, ew, ea, pr, baseline_mean=bmean, inv_batch=inv_batch, l2_weight=l2_weight)
    torch.cuda.synchronize()
    eng.check_range()
    res = {k: (None if v is None else v.detach().cpu().clone()) for k, v in out.items()}
    g = g.cpu()
    eng.close()
    return res, g


def compare(ocfg, g_cuda, grads_ref, rel=2e-4):
    off, worst = 0, (0.0, "")
    for name, shape in O.param_spec(ocfg):
        ref = grads_ref[name].reshape(-1).double()
        got = g_cuda[off:off + ref.numel()].double()
        off += ref.numel()
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        tol = rel * scale + 1e-7
        if err / tol > worst[0]:
            worst = (err / tol, f"{name}: err {err:.3e} vs max|g| {scale:.3e}")
        assert err <= tol, f"{name}: max err {err:.3e} > tol {tol:.3e} (max|g| {scale:.3e})"
    return worst


def well_conditioned(ocfg, pc, params, img, noise, min_scale=0.05):
    """The gradient of a canvas whose SAMPLED scale s_x or s_y lands within min_scale of 0 is ill-conditioned (1 / s and
    1 / s^2 factors of the inverse transformer amplify the forward pass's fp32 rounding by 1e3 and more; measured: the
    fp32 oracle itself is then 1e-3 of max|g| away from the float64 oracle).  Such draws are replaced by the posterior
    mean (eps_where = 0) so that the comparison measures the backward kernels, not the conditioning of the draw."""
    with torch.no_grad():
        res = O.forward(ocfg, pc, {k: v.double() for k, v in params.items()}, img.double(),
                        *(n.double() for n in noise), global_step=0)
    where = res["outs"]["where"]
    bad = (where[..., 0].abs() < min_scale) | (where[..., 2].abs() < min_scale)       # [T,B]
    ew = noise[0].clone()
    ew[bad] = 0.0
    return (ew, noise[1], noise[2]), int(bad.sum())


CASES = {
    "default": dict(),
    "sampled_step_weights": dict(analytic=False),
    "no_reinforce": dict(use_reinforce=False),
    "no_prior": dict(use_prior=False),
    "shift_prior_at_posterior_mean": dict(where_shift_loc=None),
    "fixed_step_prior": dict(steps_anneal=None, steps_init=0.3),
    "weighted_step_kl": dict(steps_weight=2.5, what_scale=0.7, where_scale_loc=0.4, where_shift_scale=1.3),
}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("shape", ["tiny", "script"])
def test_backward_matches_oracle_autograd(shape, case):
    kw, B = (U.TINY, 12) if shape == "tiny" else (U.SCRIPT, 16)
    ocfg = U.oracle_cfg(**kw)
    pc = O.PriorConfig(**CASES[case])
    params, img, nums, noise = U.make_problem(ocfg, B, seed=5)
    res_o, g_ref = oracle_grads(ocfg, pc, params, img, noise, 20000)
    res_c, g = cuda_grads(ocfg, pc, params, img, noise, 20000)
    assert torch.equal(res_c["presence"].reshape(-1), res_o["outs"]["presence"].detach().reshape(-1))
    worst = compare(ocfg, g, g_ref)
    print(f"{shape}/{case}: worst {worst[1]} ({worst[0]:.2f} of tolerance)")


@pytest.mark.parametrize("tc_bwd,precision", [(True, air.AIR_PREC_FP32), (False, air.AIR_PREC_FP32),
                                              (True, air.AIR_PREC_TC_SPLIT)])
def test_backward_batch64_tensor_core_weight_gradients(tc_bwd, precision, monkeypatch):
    """At >= 64 batch rows the gradient GEMMs (dW = X^T dY, dX = dY W^T) run on the tcgen05 split engine (bf16 hi/lo
    planes, split-K); AIR_NO_TC_BWD=1 keeps them on the fp32 SIMT GEMMs; with an AIR_PREC_TC_SPLIT handle the training
    forward runs on the tensor cores as well.  All three must meet the same bar."""
    if not tc_bwd:
        monkeypatch.setenv("AIR_NO_TC_BWD", "1")
    ocfg = U.oracle_cfg(**U.SCRIPT)
    pc = O.PriorConfig()
    params, img, nums, noise = U.make_problem(ocfg, 64, seed=13)
    noise, n_fixed = well_conditioned(ocfg, pc, params, img, noise)
    res_o, g_ref = oracle_grads(ocfg, pc, params, img, noise, 20000, dtype=torch.float64)
    res_c, g = cuda_grads(ocfg, pc, params, img, noise, 20000, precision=precision)
    assert torch.equal(res_c["presence"].reshape(-1), res_o["outs"]["presence"].detach().float().reshape(-1))
    worst = compare(ocfg, g, g_ref)
    print(f"B=64 tc_bwd={tc_bwd} precision={precision} ({n_fixed} ill-conditioned draws replaced): worst {worst[1]} ({worst[0]:.2f} of tolerance)")


def test_backward_with_baseline_and_l2():
    ocfg = U.oracle_cfg(**U.SCRIPT)
    pc = O.PriorConfig()
    B = 8
    params, img, nums, noise = U.make_problem(ocfg, B, seed=7)
    baseline = (50.0 * torch.randn(B, 1, generator=torch.Generator().manual_seed(1))) + 300.0
    _, g_ref = oracle_grads(ocfg, pc, params, img, noise, 5000, baseline=baseline, l2_weight=1e-2)
    _, g = cuda_grads(ocfg, pc, params, img, noise, 5000, baseline=baseline, l2_weight=1e-2)
    compare(ocfg, g, g_ref)


def test_backward_config_d_shapes():
    """BASELINE.json configs[3] shapes (100x100 canvas, 28x28 glimpse, 5 steps) at a small batch."""
    ocfg = U.oracle_cfg(**U.CONFIG_D)
    pc = O.PriorConfig()
    params, img, nums, noise = U.make_problem(ocfg, 6, seed=2)
    # K = 10,000-term contractions and 5 steps: the fp32 oracle itself is ~3e-4 away from the exact gradient here, so
    # the reference is the oracle evaluated in float64 (the value both fp32 implementations approximate)
    res64, g_ref = oracle_grads(ocfg, pc, params, img, noise, 20000, dtype=torch.float64)
    res_c, g = cuda_grads(ocfg, pc, params, img, noise, 20000)
    assert torch.equal(res_c["presence"].reshape(-1), res64["outs"]["presence"].detach().float().reshape(-1))
    _, g32 = oracle_grads(ocfg, pc, params, img, noise, 20000)
    flat32 = O.flatten_params(ocfg, g32)
    w_cuda, w_o32 = compare(ocfg, g, g_ref), compare(ocfg, flat32, g_ref, rel=1.0)
    print(f"config D vs float64 oracle: cuda worst {w_cuda[1]}; fp32 oracle worst {w_o32[1]}")


def test_sharded_gradients_sum_to_the_whole_batch_gradient():
    """Two 'ranks' on one device: each runs its contiguous shard with inv_batch = 1 / global batch and the global baseline
    mean; the SUM of the two gradient buffers (what the all-reduce forms) is the whole-batch gradient."""
    ocfg = U.oracle_cfg(**U.TINY)
    pc = O.PriorConfig()
    B = 10
    params, img, nums, noise = U.make_problem(ocfg, B, seed=11)
    _, g_ref = oracle_grads(ocfg, pc, params, img, noise, 20000)
    total = None
    for a, b in ((0, 6), (6, 10)):
        _, g = cuda_grads(ocfg, pc, params, img[a:b], tuple(n[:, a:b].contiguous() for n in noise), 20000,
                          inv_batch=1.0 / B)
        total = g if total is None else total + g
    compare(ocfg, total, g_ref)


def test_backward_small_batch_on_the_tensor_core_engine():
    """AIR_PREC_TC_SPLIT handle, batch below the tensor-core gradient threshold: tcgen05 forward with kept activations,
    SIMT gradient GEMMs."""
    ocfg = U.oracle_cfg(**U.SCRIPT)
    pc = O.PriorConfig()
    params, img, nums, noise = U.make_problem(ocfg, 16, seed=5)
    res_o, g_ref = oracle_grads(ocfg, pc, params, img, noise, 20000)
    res_c, g = cuda_grads(ocfg, pc, params, img, noise, 20000, precision=air.AIR_PREC_TC_SPLIT)
    assert torch.equal(res_c["presence"].reshape(-1), res_o["outs"]["presence"].detach().reshape(-1))
    U.assert_close(res_c["loss_per_sample"], res_o["loss_per_sample"].detach(), atol=0, rtol=1e-4, name="loss_per_sample")
    compare(ocfg, g, g_ref)


def test_backward_requires_training_mode():
    ocfg = U.oracle_cfg(**U.TINY)
    eng = air.Engine(U.cell_cfg(ocfg, air.AIR_PREC_FP32), 4, ocfg.T, device="cuda")
    flat = torch.zeros(eng.n_params, device="cuda")
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(air.AirError):
        eng.backward(flat, z(4, 3, 3), z(3, 4, 4), z(3, 4, 10), U.prior_struct(O.PriorConfig(), 0))
    eng.close()


@pytest.mark.parametrize("case", ["default", "sampled_step_weights", "no_reinforce"])
@pytest.mark.parametrize("shape,precision", [("tiny", air.AIR_PREC_FP32), ("script", air.AIR_PREC_FP32),
                                             ("script", air.AIR_PREC_TC_SPLIT)])
def test_backward_non_discrete_steps(shape, precision, case):
    """discrete_steps = False (cell.py:150-151): presence_t = presence_prob_t, so the painted canvas -- and with
    analytic = False the KL weights -- are differentiable functions of the steps predictor.  Autograd on the oracle is the
    reference (float64 at the script sizes, where it is what both fp32 implementations approximate)."""
    kw, B = (U.TINY, 12) if shape == "tiny" else (U.SCRIPT, 64 if precision == air.AIR_PREC_TC_SPLIT else 16)
    ocfg = U.oracle_cfg(**kw, discrete_steps=False)
    pc = O.PriorConfig(**CASES[case])
    params, img, nums, noise = U.make_problem(ocfg, B, seed=15)
    if shape == "script":
        noise, _ = well_conditioned(ocfg, pc, params, img, noise)
    dt = torch.float64 if shape == "script" else torch.float32
    res_o, g_ref = oracle_grads(ocfg, pc, params, img, noise, 20000, dtype=dt)
    res_c, g = cuda_grads(ocfg, pc, params, img, noise, 20000, precision=precision)
    U.assert_close(res_c["presence"], res_o["outs"]["presence"].detach().float(), atol=1e-5, name="presence")
    worst = compare(ocfg, g, g_ref)
    # the steps predictor must receive the canvas's gradient: its output weights' gradient is not what the discrete path gives
    print(f"non-discrete {shape}/{case}: worst {worst[1]} ({worst[0]:.2f} of tolerance)")


def test_centered_rmsprop_matches_oracle():
    n = 10007
    g = torch.Generator().manual_seed(3)
    theta = torch.randn(n, generator=g)
    mg, ms, mom = torch.zeros(n), torch.ones(n), torch.zeros(n)
    d = [t.clone().cuda() for t in (theta, mg, ms, mom)]
    ocfg = U.oracle_cfg(**U.TINY)
    eng = air.Engine(U.cell_cfg(ocfg), 2, ocfg.T, device="cuda")
    for step in range(5):
        grad = torch.randn(n, generator=g) * (10.0 ** (step - 2))
        theta, mg, ms, mom = O.centered_rmsprop_step(theta, grad, mg, ms, mom, lr=1e-3)
        eng.rmsprop_step(d[0], grad.cuda(), d[1], d[2], d[3], 1e-3)
    torch.cuda.synchronize()
    for got, ref, name in zip(d, (theta, mg, ms, mom), ("theta", "mg", "ms", "mom")):
        U.assert_close(got.cpu(), ref, atol=1e-6, rtol=1e-5, name=name)
    eng.close()


def test_centered_rmsprop_stays_finite_on_a_nearly_constant_gradient():
    """ms - mg^2 rounds below zero in fp32 once an element's gradient all but stops changing (a fixed batch: NaN parameters
    after ~140 steps with the formula taken literally); the kernel clamps the variance estimate at 0."""
    n = 20000
    g = torch.Generator().manual_seed(5)
    g0 = torch.randn(n, generator=g) * torch.logspace(-6, 1, n)
    theta, mg, ms, mom = torch.zeros(n).cuda(), torch.zeros(n).cuda(), torch.ones(n).cuda(), torch.zeros(n).cuda()
    mg_r, ms_r, negative = torch.zeros(n), torch.ones(n), 0
    ocfg = U.oracle_cfg(**U.TINY)
    eng = air.Engine(U.cell_cfg(ocfg), 2, ocfg.T, device="cuda")
    for step in range(300):
        grad = g0 * (1.0 + 1e-4 * torch.randn(n, generator=g))
        eng.rmsprop_step(theta, grad.cuda(), mg, ms, mom, 1e-5)
        # the literal formula on the same sequence (otherwise this test checks nothing)
        mg_r = mg_r + (1.0 - 0.9) * (grad - mg_r)
        ms_r = ms_r + (1.0 - 0.9) * (grad * grad - ms_r)
        negative += int(((ms_r - mg_r * mg_r + 1e-10) < 0).sum())
    torch.cuda.synchronize()
    assert negative > 0
    for t in (theta, mg, ms, mom):
        assert bool(torch.isfinite(t).all())
    eng.close()


def test_training_loop_matches_oracle_loop():
    """Three full training steps (forward + ELBO, backward, centered RMSProp) on the engine against the same loop on the
    oracle (autograd + the restated ApplyCenteredRMSProp): losses and the final parameters agree."""
    ocfg = U.oracle_cfg(**U.SCRIPT)
    pc = O.PriorConfig()
    B, lr, steps = 8, 1e-4, 3
    params, img, nums, _ = U.make_problem(ocfg, B, seed=9)
    dev = "cuda"
    eng = air.Engine(U.cell_cfg(ocfg, air.AIR_PREC_FP32), B, ocfg.T, device=dev)
    eng.train_enable(True)
    flat = O.flatten_params(ocfg, params).to(dev)
    n = flat.numel()
    mg_d, ms_d, mom_d = torch.zeros(n, device=dev), torch.ones(n, device=dev), torch.zeros(n, device=dev)
    grad_d = torch.empty(n, device=dev)
    theta = O.flatten_params(ocfg, params).clone()
    mg, ms, mom = torch.zeros(n), torch.ones(n), torch.zeros(n)
    img_d = img.to(dev).contiguous()
    for step in range(steps):
        noise = O.make_noise(ocfg, B, seed=100 + step)
        gs = 2000 * step
        # oracle
        p = {k: v.clone().requires_grad_(True) for k, v in O.unflatten_params(ocfg, theta).items()}
        res = O.forward(ocfg, pc, p, img, *noise, global_step=gs)
        res["opt_loss"].backward()
        g = O.flatten_params(ocfg, {k: v.grad for k, v in p.items()})
        theta, mg, ms, mom = O.centered_rmsprop_step(theta.detach(), g, mg, ms, mom, lr)
        # engine
        ew, ea, u = (t.to(dev).contiguous() for t in noise)
        pr = U.prior_struct(pc, gs)
        out = eng.forward(flat, img_d, ew, ea, u, pr)
        eng.backward(flat, img_d, ew, ea, pr, grad_d)
        eng.rmsprop_step(flat, grad_d, mg_d, ms_d, mom_d, lr)
        torch.cuda.synchronize()
        assert torch.equal(out["presence"].cpu().reshape(-1), res["outs"]["presence"].detach().reshape(-1)), step
        # after the first update the two parameter vectors differ at the fp32-noise level in EVERY low-gradient coordinate
        # (RMSProp normalises each coordinate's step to ~lr whatever the gradient's size), hence the wider band
        U.assert_close(out["scalars"][air._lib.SCALAR_INDEX["opt_loss"]].cpu(), res["opt_loss"].detach(), atol=0,
                       rtol=2e-4 if step == 0 else 2e-3, name=f"opt_loss step {step}")
    # RMSProp normalises every coordinate's step to ~lr: compare the parameter CHANGE, coordinates with a tiny gradient
    # (where fp32 noise decides the direction) are bounded by the step size itself
    d_ref = theta - O.flatten_params(ocfg, params)
    d_got = flat.cpu() - O.flatten_params(ocfg, params)
    err = (d_got - d_ref).abs()
    assert float(err.max()) <= 2.5 * steps * lr * 3.2, float(err.max())
    assert float((err > 0.05 * steps * lr).float().mean()) < 0.02, float((err > 0.05 * steps * lr).float().mean())
    eng.close()


def test_model_train_op_decreases_the_loss():
    """AIRModel.train_step -> train_op on a fixed batch: the optimised loss goes down (model.py:261-376 surface)."""
    B, T = 64, 3
    img, nums = O.synthetic_multi_mnist(B, 50, 50, seed=5)
    x, y = img.cuda(), nums.cuda()
    from functools import partial
    model = air.AIRModel(x, y, T, (20, 20), 50, air.LSTM(256), partial(air.Encoder, [256, 256]),
                         partial(air.Encoder, [256, 256]), partial(air.Decoder, [256, 256]),
                         partial(air.StochasticTransformParam, [256, 256], scale_bias=.5),
                         partial(air.StepsPredictor, [128, 64], .75), output_std=.3, output_multiplier=.5,
                         explore_eps=1e-3)
    pr = dict(loc=0., scale=1.)
    nsp = dict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3)
    train_op, global_step = model.train_step(1e-4, 0., pr, pr, pr, nsp, use_reinforce=True)
    gen = torch.Generator(device="cuda").manual_seed(0)
    losses = []
    for i in range(60):
        train_op(noise=model.cell.draw_noise(B, T, generator=gen))
        losses.append(float(model.loss.value))
    assert global_step() == 60
    first, last = sum(losses[:5]) / 5, sum(losses[-5:]) / 5
    print(f"loss {first:.1f} -> {last:.1f}")
    assert last < 0.8 * first, (first, last)


def test_baseline_mlp_training_step_matches_oracle():
    """BaselineMLP (modules.py:125-143) forward, the gradient of baseline_loss (model.py:253-259, [B]-[B,1] broadcast) and
    one RMSProp step at 10x lr against autograd on the oracle's baseline_mlp."""
    B, T, na, nh, P = 24, 3, 50, 256, 2500
    g = torch.Generator().manual_seed(4)
    img = torch.rand(B, 50, 50, generator=g)
    what, where, pres = torch.randn(T, B, na, generator=g), torch.randn(T, B, 4, generator=g), torch.rand(T, B, 1, generator=g)
    h, c = torch.randn(B, nh, generator=g), torch.randn(B, nh, generator=g)
    target = 300.0 + 40.0 * torch.randn(B, generator=g)
    bm = air.BaselineMLP([256, 128])
    d = lambda t: t.cuda()
    b = bm(d(img), d(what), d(where), d(pres), (d(h), d(c)))
    # oracle with the module's own initial parameters
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in bm.views.items()}
    b_ref = O.baseline_mlp(p, 2, img, what, where, pres, h, c)
    U.assert_close(b.cpu(), b_ref.detach(), atol=2e-4, rtol=1e-4, name="baseline")
    loss = 0.5 * ((target[None, :] - b_ref) ** 2).mean()          # [B] - [B,1] -> [B,B]
    loss.backward()
    theta0 = bm.params.detach().cpu().clone()
    grad = bm.backward(d(target), b).cpu()
    off = 0
    for name, (r, cc) in bm._spec:
        ref = p[name].grad.reshape(-1)
        got = grad[off:off + r * cc]
        off += r * cc
        tol = 2e-4 * float(ref.abs().max()) + 1e-7
        assert float((got - ref).abs().max()) <= tol, (name, float((got - ref).abs().max()), tol)
    eng = air.Engine(U.cell_cfg(U.oracle_cfg(**U.TINY)), 2, 3, device="cuda")
    eng.rmsprop_step(bm.params, bm.grad, bm.slots["mg"], bm.slots["ms"], bm.slots["mom"], 1e-3)
    n = theta0.numel()
    ref_theta, _, _, _ = O.centered_rmsprop_step(theta0, O.flatten_params, None, None, None, 0) if False else \
        O.centered_rmsprop_step(theta0, torch.cat([p[nm].grad.reshape(-1) for nm, _ in bm._spec]), torch.zeros(n),
                                torch.ones(n), torch.zeros(n), 1e-3)
    U.assert_close(bm.params.cpu() - theta0, ref_theta - theta0, atol=2e-6, rtol=2e-3, name="baseline update")
    eng.close()


def test_nvil_normalised_reinforce_matches_oracle():
    """decay_rate branch of _reinforce (model.py:232-239): importance weight shifted by the moving mean and divided by
    max(sqrt(moving var), 1) -- loss scalars, the batch moments that feed the moving averages, and the gradient."""
    ocfg = U.oracle_cfg(**U.SCRIPT)
    pc = O.PriorConfig()
    B = 16
    params, img, nums, noise = U.make_problem(ocfg, B, seed=21)
    baseline = 250.0 + 30.0 * torch.randn(B, 1, generator=torch.Generator().manual_seed(2))
    mm, mv = 40.0, 900.0
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    res = O.forward(ocfg, pc, p, img, *noise, global_step=20000, baseline=baseline, nvil=(mm, mv))
    res["opt_loss"].backward()
    g_ref = {k: v.grad for k, v in p.items()}
    dev = "cuda"
    eng = air.Engine(U.cell_cfg(ocfg, air.AIR_PREC_FP32), B, ocfg.T, device=dev)
    eng.train_enable(True)
    flat = O.flatten_params(ocfg, params).to(dev)
    ew, ea, u = (n.to(dev).contiguous() for n in noise)
    img_d = img.to(dev).contiguous()
    pr = U.prior_struct(pc, 20000)
    pr.nvil_shift, pr.nvil_scale = mm, 1.0 / max(mv ** 0.5, 1.0)
    out = eng.forward(flat, img_d, ew, ea, u, pr, baseline.reshape(-1).to(dev).contiguous())
    g = eng.backward(flat, img_d, ew, ea, pr, baseline_mean=float(baseline.mean())).cpu()
    sc = out["scalars"].cpu()
    SI = air._lib.SCALAR_INDEX
    U.assert_close(sc[SI["reinforce_loss"]], res["reinforce_loss"].detach(), atol=1e-3, rtol=2e-4, name="reinforce_loss")
    U.assert_close(sc[SI["opt_loss"]], res["opt_loss"].detach(), atol=1e-3, rtol=2e-4, name="opt_loss")
    mean_ref, var_ref = (float(v) for v in res["imp_weight_moments"])
    mean = float(sc[SI["mean_iw"]] - sc[SI["mean_baseline"]])
    var = float(sc[SI["mean_iw2"]] - sc[SI["mean_iw"]] ** 2 + sc[SI["mean_baseline2"]] - sc[SI["mean_baseline"]] ** 2)
    assert abs(mean - mean_ref) <= 1e-4 * abs(mean_ref) + 1e-3, (mean, mean_ref)
    assert abs(var - var_ref) <= 2e-3 * abs(var_ref), (var, var_ref)
    compare(ocfg, g, g_ref)
    eng.close()


def test_model_decay_rate_tracks_the_moving_moments():
    """AIRModel.train_step(decay_rate=...): after one train_op the moving mean / variance of the importance weight are
    make_moving_average(init 0 / 1) of the batch moments (ops.py:46-64)."""
    from functools import partial
    B, T = 32, 3
    img, nums = O.synthetic_multi_mnist(B, 50, 50, seed=6)
    model = air.AIRModel(img.cuda(), nums.cuda(), T, (20, 20), 50, air.LSTM(256), partial(air.Encoder, [256, 256]),
                         partial(air.Encoder, [256, 256]), partial(air.Decoder, [256, 256]),
                         partial(air.StochasticTransformParam, [256, 256], scale_bias=.5),
                         partial(air.StepsPredictor, [128, 64], .75), output_std=.3, output_multiplier=.5,
                         explore_eps=1e-3)
    pr = dict(loc=0., scale=1.)
    nsp = dict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3)
    train_op, _ = model.train_step(1e-5, 0., pr, pr, pr, nsp, decay_rate=0.9)
    noise = model.cell.draw_noise(B, T, generator=torch.Generator(device="cuda").manual_seed(1))
    train_op(noise=noise)
    iw = model.rec_loss_per_sample.double().cpu()
    mean, var = float(iw.mean()), float(iw.var(unbiased=False))
    assert abs(model.imp_weight_moving_mean - 0.1 * mean) <= 1e-4 * abs(mean)
    assert abs(model.imp_weight_moving_var - (0.9 + 0.1 * var)) <= 2e-3 * var
    assert model._prior_struct.nvil_scale == 1.0 and model._prior_struct.nvil_shift == 0.0   # values BEFORE the update
    train_op(noise=noise)
    assert abs(model._prior_struct.nvil_shift - 0.1 * mean) <= 1e-4 * abs(mean)
    assert abs(model._prior_struct.nvil_scale - 1.0 / max((0.9 + 0.1 * var) ** 0.5, 1.0)) < 1e-3


@pytest.mark.parametrize("T,B", [(1, 5), (8, 3), (2, 1)])
def test_backward_edge_step_counts_and_batch_sizes(T, B):
    """One step, AIR_MAX_STEPS steps, a single canvas: every templated backward kernel at the ends of its range."""
    kw = dict(U.TINY)
    kw["T"] = T
    ocfg = U.oracle_cfg(**kw)
    pc = O.PriorConfig()
    params, img, nums, noise = U.make_problem(ocfg, B, seed=31 + T)
    noise, _ = well_conditioned(ocfg, pc, params, img, noise, min_scale=0.1)
    res_o, g_ref = oracle_grads(ocfg, pc, params, img, noise, 20000, dtype=torch.float64)
    res_c, g = cuda_grads(ocfg, pc, params, img, noise, 20000)
    assert torch.equal(res_c["presence"].reshape(-1), res_o["outs"]["presence"].detach().float().reshape(-1))
    compare(ocfg, g, g_ref)
