"""GPU parity tests: the CUDA library (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): absolute 1e-4 on O(1) outputs, relative 1e-4 on the ELBO sums, exact on
presence / step counts / gather indices (with the near-tie rule of tests/util.py:presence_mismatches).
"""
import math

import numpy as np
import pytest
import torch

import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200 import functional as AF
from oracle import air_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _where(B, g, lo=0.2, hi=1.1):
    sx = torch.rand(B, 1, generator=g) * (hi - lo) + lo
    sy = torch.rand(B, 1, generator=g) * (hi - lo) + lo
    tx = torch.rand(B, 1, generator=g) * 1.6 - 0.8
    ty = torch.rand(B, 1, generator=g) * 1.6 - 0.8
    return torch.cat([sx, tx, sy, ty], 1)


# ----------------------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W,h,w", [(50, 50, 20, 20), (100, 100, 28, 28), (3, 3, 2, 2), (17, 31, 5, 9)])
def test_stn_read(H, W, h, w):
    g = torch.Generator().manual_seed(0)
    B = 33
    img = torch.rand(B, H, W, generator=g)
    where = _where(B, g)
    where[0] = torch.tensor([1., 0., 1., 0.])            # identity-sized crop
    where[1] = torch.tensor([-0.7, 0.3, 0.5, -0.2])      # negative scale (mirrored), unconstrained after sampling
    where[2] = torch.tensor([2.5, 0.9, 2.5, -0.9])       # mostly outside the image -> zero padding
    where[3] = torch.tensor([1e-4, 0., 1e-4, 0.])        # degenerate: all samples in one cell
    ref = O.stn_read(img, where, (h, w))
    out = AF.stn_read(img.to(DEV), where.to(DEV), (h, w)).cpu()
    U.assert_close(out, ref, atol=2e-6, rtol=1e-6, name="stn_read")


@pytest.mark.parametrize("H,W,h,w", [(50, 50, 20, 20), (100, 100, 28, 28), (3, 3, 2, 2), (17, 31, 5, 9)])
def test_stn_paint(H, W, h, w):
    g = torch.Generator().manual_seed(1)
    B = 33
    gl = torch.randn(B, h, w, generator=g)
    where = _where(B, g)
    where[0] = torch.tensor([1., 0., 1., 0.])
    where[1] = torch.tensor([-0.7, 0.3, 0.5, -0.2])
    where[2] = torch.tensor([0.05, 0.9, 0.05, -0.9])     # tiny object near the border
    where[3] = torch.tensor([0.0, 0.1, 0.5, 0.1])        # sx = 0 -> det = 0 -> inf/nan coords -> zeros (App. C5)
    ref = O.stn_paint(gl, where, (H, W))
    out = AF.stn_paint(gl.to(DEV), where.to(DEV), (H, W)).cpu()
    assert torch.isfinite(out).all()
    U.assert_close(out, ref, atol=1e-5, rtol=1e-5, name="stn_paint")


def test_stn_identity_roundtrip():
    img = torch.rand(5, 20, 20)
    where = torch.tensor([[1., 0., 1., 0.]]).repeat(5, 1)
    assert torch.allclose(AF.stn_read(img.to(DEV), where.to(DEV), (20, 20)).cpu(), img, atol=1e-5)
    assert torch.allclose(AF.stn_paint(img.to(DEV), where.to(DEV), (20, 20)).cpu(), img, atol=1e-5)


@pytest.mark.parametrize("M,K,N,act", [(64, 2500, 256, 1), (192, 256, 8, 0), (130, 50, 256, 1), (7, 17, 1, 0),
                                       (257, 400, 100, 0), (1, 3, 5, 1), (300, 256, 1024, 0)])
def test_linear(M, K, N, act):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    ref = x @ w + b
    if act:
        ref = O.elu(ref)
    out = AF.linear(x.to(DEV), w.to(DEV), b.to(DEV), act).cpu()
    U.assert_close(out, ref, atol=2e-5, rtol=2e-5, name="linear")


def test_lstm_step():
    g = torch.Generator().manual_seed(3)
    B, nx, nh = 37, 256, 256
    x, h, c = torch.randn(B, nx, generator=g), torch.randn(B, nh, generator=g), torch.randn(B, nh, generator=g)
    w = torch.randn(nx + nh, 4 * nh, generator=g) / math.sqrt(nx + nh)
    b = 0.1 * torch.randn(4 * nh, generator=g)
    h_ref, c_ref = O.lstm_step(x, h, c, w, b, 1.0)
    h_out, c_out = AF.lstm_step(x.to(DEV), h.to(DEV), c.to(DEV), w.to(DEV), b.to(DEV), 1.0)
    U.assert_close(h_out.cpu(), h_ref, atol=1e-5, name="lstm h")
    U.assert_close(c_out.cpu(), c_ref, atol=1e-5, name="lstm c")


# ----------------------------------------------------------------------------------------------------------
# the unrolled forward pass + ELBO
# ----------------------------------------------------------------------------------------------------------
CELL_KEYS = ["glimpse", "what", "what_loc", "what_scale", "where", "where_loc", "where_scale", "presence_prob"]


def _check_forward(ocfg, B, pc, seed=0, global_step=0, baseline=None, precision=air.AIR_PREC_FP32, atol=1e-4,
                   weight_gain=1.0, noise_floor=False):
    """noise_floor=True: the fp32 reference itself is only defined up to its own rounding noise, which large weights
    amplify through the 30-layer chain.  The oracle is then also run in float64 and every absolute tolerance is
    widened by 4x the fp32-oracle-vs-fp64-oracle distance of that tensor ('float64 copy'); 8x for the
    tensor-core split engine, whose operands carry 22 significand bits instead of 24 (tools/accuracy_probe.py prints
    every engine's distance to the float64 oracle: at unit gain all of them sit at 1e-6..3e-5)."""
    params, img, nums, noise = U.make_problem(ocfg, B, seed, weight_gain)
    ref = O.forward(ocfg, pc, params, img, *noise, global_step=global_step, baseline=baseline)
    out = U.run_cuda(ocfg, params, img, noise, pc, global_step, baseline, precision)
    T = ocfg.T
    floor = {}
    if noise_floor:
        r64 = O.forward(ocfg, pc, {k: v.double() for k, v in params.items()}, img.double(),
                        *(n.double() for n in noise), global_step=global_step,
                        baseline=None if baseline is None else baseline.double())
        mult = 8.0 if precision == air.AIR_PREC_TC_SPLIT else 4.0
        for k in CELL_KEYS:
            floor[k] = mult * float((ref["outs"][k].double() - r64["outs"][k]).abs().max())
        for k in ("canvas", "glimpse", "final_h", "final_c", "rec_loss_per_sample", "kl_what_per_sample",
                  "kl_where_per_sample", "kl_num_steps_per_sample", "loss_per_sample", "num_steps_log_prob"):
            if k in ref and k in r64:
                floor["ps:" + k] = mult * float((ref[k].double() - r64[k]).abs().max())
    for k in CELL_KEYS:
        U.assert_close(out[k], ref["outs"][k], atol=atol + floor.get(k, 0.0), rtol=1e-4, name=k)
    if ocfg.discrete_steps:
        bad, unsafe = U.presence_mismatches(out["presence"], ref["outs"]["presence_prob"].reshape(T, B),
                                            noise[2].reshape(T, B))
        assert bad == 0, f"{bad} presence mismatches away from ties"
        assert unsafe <= max(1, B * T // 1000)
        same = (out["presence"].reshape(T, B) == ref["outs"]["presence"].reshape(T, B)).all(0)
    else:
        U.assert_close(out["presence"], ref["outs"]["presence"], atol=atol, name="presence")
        same = torch.ones(B, dtype=torch.bool)
    # canvas / losses only compared on samples whose discrete path agrees (all of them unless a near-tie flipped)
    assert same.float().mean() > 0.995
    canvas_ref = ref["canvas"].reshape(T, B, -1)
    U.assert_close(out["canvas"][:, same], canvas_ref[:, same], atol=atol + floor.get("ps:canvas", 0.0), rtol=1e-4,
                   name="canvas")
    U.assert_close(out["glimpse_viz"][:, same], ref["glimpse"].reshape(T, B, -1)[:, same],
                   atol=atol + floor.get("ps:glimpse", 0.0), name="glimpse_viz")
    U.assert_close(out["final_h"], ref["final_h"], atol=atol + floor.get("ps:final_h", 0.0), name="final_h")
    U.assert_close(out["final_c"], ref["final_c"], atol=atol + floor.get("ps:final_c", 0.0), name="final_c")
    U.assert_close(out["num_steps_posterior"], ref["num_steps_posterior"], atol=1e-5, rtol=1e-4, name="q(n)")
    if ocfg.discrete_steps:
        assert torch.equal(out["num_step_per_sample"][same], ref["num_step_per_sample"][same])
    else:
        U.assert_close(out["num_step_per_sample"], ref["num_step_per_sample"], atol=1e-5, name="num_step_per_sample")
    U.assert_close(out["prior_step_weight"], ref["prior_step_weight"], atol=1e-5, rtol=1e-4, name="step weight")
    for k_c, k_o in [("rec_loss_per_sample", "rec_loss_per_sample"), ("kl_num_steps_per_sample", "kl_num_steps_per_sample"),
                     ("kl_what_per_sample", "kl_what_per_sample"), ("kl_where_per_sample", "kl_where_per_sample"),
                     ("loss_per_sample", "loss_per_sample"), ("num_steps_log_prob", "num_steps_log_prob")]:
        if k_o in ref:      # num_steps_log_prob only exists when REINFORCE is on
            # per-sample sums range over several hundred and change sign across the batch: "relative 1e-4" is taken
            # against the batch's mean magnitude of the term (never tighter than 1e-4 absolute)
            scale = max(1.0, float(ref[k_o][same].abs().mean())) if bool(same.any()) else 1.0
            U.assert_close(out[k_c][same], ref[k_o][same], atol=1e-4 * scale + floor.get("ps:" + k_o, 0.0), rtol=1e-4,
                           name=k_c)
    if bool(same.all()) and not noise_floor:
        s = out["scalars"]
        idx = air._lib.SCALAR_INDEX
        for name, key in [("rec_loss", "rec_loss"), ("kl_num_steps", "kl_num_steps"), ("kl_what", "kl_what"),
                          ("kl_where", "kl_where"), ("prior_loss", "prior_loss"), ("loss", "loss"),
                          ("opt_loss", "opt_loss"), ("num_step", "num_step")]:
            U.assert_close(s[idx[name]], ref[key].float(), atol=1e-4, rtol=1e-4, name="scalar " + name)
        if pc.use_reinforce:
            U.assert_close(s[idx["reinforce_loss"]], ref["reinforce_loss"].float(), atol=2e-4, rtol=2e-4,
                           name="reinforce")
        U.assert_close(-s[idx["loss"]], ref["elbo"].float(), atol=0, rtol=1e-4, name="ELBO")
    return out, ref


def test_forward_script_config_b64():
    """BASELINE.json configs[0]: 50x50, T=3, B=64 -- annealed float64 step prior at a mid-anneal global step."""
    _check_forward(U.oracle_cfg(**U.SCRIPT), 64, O.PriorConfig(), seed=0, global_step=20000)


def test_forward_step0_prior_edge():
    """global_step 0: success prob 1 - 1e-15 (not representable in fp32) -> float64 island matters."""
    _check_forward(U.oracle_cfg(**U.SCRIPT), 32, O.PriorConfig(), seed=1, global_step=0)


def test_forward_tiny_odd_shapes():
    """Widths of test/cell_test.py (5/7/11/13/17 hidden units, 3x3 image, 2x2 crop): unaligned everything."""
    _check_forward(U.oracle_cfg(**U.TINY), 10, O.PriorConfig(), seed=2, global_step=5000, weight_gain=2.0)


def test_forward_config_d_small_batch():
    """BASELINE.json configs[3] shapes (100x100 canvas, 28x28 glimpse, T=5) at a small batch."""
    _check_forward(U.oracle_cfg(**U.CONFIG_D), 12, O.PriorConfig(), seed=3, global_step=30000)


def test_forward_variants():
    base = dict(U.SCRIPT)
    # no explore eps, continuous steps, non-analytic weights, fixed float32 prior, shift prior without loc, no prior
    _check_forward(U.oracle_cfg(explore_eps=None, **base), 16,
                   O.PriorConfig(steps_anneal=None, steps_init=0.3, where_shift_loc=None), seed=4)
    _check_forward(U.oracle_cfg(discrete_steps=False, **base), 16, O.PriorConfig(analytic=False, steps_weight=0.5),
                   seed=5, global_step=50000)
    _check_forward(U.oracle_cfg(**base), 16, O.PriorConfig(use_prior=False, use_reinforce=False,
                                                             what_scale=2.0, where_scale_loc=0.5, where_shift_scale=0.7),
                   seed=6, global_step=90000)


def test_forward_with_baseline_bb_broadcast():
    """REINFORCE with a [B,1] baseline: the reference's [B]-[B,1] -> [B,B] broadcast."""
    B = 24
    baseline = 50.0 * torch.randn(B, 1, generator=torch.Generator().manual_seed(9))
    _check_forward(U.oracle_cfg(**U.SCRIPT), B, O.PriorConfig(), seed=7, global_step=15000, baseline=baseline)


def test_forward_ragged_batch_sizes():
    for B in (2, 3, 129):
        _check_forward(U.oracle_cfg(**U.SCRIPT), B, O.PriorConfig(), seed=10 + B, global_step=12000)


def test_forward_trained_like_weights():
    """Larger weights push activations, softplus and the STN through their non-linear ranges."""
    _check_forward(U.oracle_cfg(**U.SCRIPT), 48, O.PriorConfig(), seed=8, global_step=60000, weight_gain=2.5,
                   noise_floor=True)


def test_cell_step_chain_equals_unroll():
    """AIRCell.__call__ T times (the RNNCore contract, cell.py:116-171) == the fused unroll == the oracle."""
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B = 20
    params, img, nums, noise = U.make_problem(ocfg, B, seed=11)
    ref, _ = O.unroll(ocfg, params, img, *noise)
    cell = air.AIRCell((ocfg.H, ocfg.W), (ocfg.h, ocfg.w), ocfg.na, air.LSTM(ocfg.nh),
                       lambda: air.Encoder(ocfg.enc_hidden), lambda: air.Encoder(ocfg.glenc_hidden),
                       lambda s: air.Decoder(ocfg.dec_hidden, s),
                       lambda n: air.StochasticTransformParam(ocfg.where_hidden, n, scale_bias=ocfg.scale_bias),
                       lambda: air.StepsPredictor(ocfg.steps_hidden, ocfg.step_bias),
                       explore_eps=ocfg.explore_eps, device=DEV)
    cell.params.copy_(O.flatten_params(ocfg, params).to(DEV))
    state = cell.initial_state(img.to(DEV))
    assert [tuple(s.shape) if torch.is_tensor(s) else tuple(tuple(x.shape) for x in s) for s in state] == \
        [(B, 2500), (B, 2500), (B, 50), (B, 4), ((B, 256), (B, 256)), (B, 1)]
    for t in range(ocfg.T):
        outs, state = cell(None, state, noise=tuple(n[t].to(DEV) for n in noise))
        assert len(outs) == 10 and len(state) == 6
        for name, o in zip(cell.output_names, outs):
            U.assert_close(o.cpu(), ref[name][t], atol=1e-4, name=f"step {t} {name}")


# ----------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]: B = 4096)
# ----------------------------------------------------------------------------------------------------------
def test_full_size_batch_split_invariance_and_oracle_sample():
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B = 4096
    pc = O.PriorConfig()
    params, img, nums, noise = U.make_problem(ocfg, B, seed=21)
    full = U.run_cuda(ocfg, params, img, noise, pc, 20000)
    # (1) samples are independent: the two halves run separately give bit-identical per-sample results
    half = B // 2
    for lo in (0, half):
        part = U.run_cuda(ocfg, params, img[lo:lo + half], tuple(n[:, lo:lo + half].contiguous() for n in noise), pc, 20000)
        for k in ("loss_per_sample", "rec_loss_per_sample", "num_step_per_sample", "kl_what_per_sample"):
            assert torch.equal(part[k], full[k][lo:lo + half]), k
        assert torch.equal(part["what"], full["what"][:, lo:lo + half])
        assert torch.equal(part["canvas"], full["canvas"][:, lo:lo + half])
    # (2) batch mean == mean of the per-sample vector
    s = full["scalars"]
    idx = air._lib.SCALAR_INDEX
    assert abs(float(s[idx["loss"]]) - float(full["loss_per_sample"].double().mean())) < 1e-4 * abs(float(s[idx["loss"]]))
    # (3) q(n) is a pmf; presence is monotone non-increasing in t and in {0,1}
    assert torch.allclose(full["num_steps_posterior"].sum(1), torch.ones(B), atol=1e-6)
    pres = full["presence"].reshape(ocfg.T, B)
    assert ((pres == 0) | (pres == 1)).all() and (pres[1:] <= pres[:-1]).all()
    assert torch.equal(full["num_step_per_sample"], pres.sum(0))
    # (4) a strided sample of 64 canvases against the oracle
    sel = torch.arange(0, B, 64)
    ref = O.forward(ocfg, pc, params, img[sel], *(n[:, sel] for n in noise), global_step=20000)
    U.assert_close(full["loss_per_sample"][sel], ref["loss_per_sample"], atol=0, rtol=1e-4, name="loss sample")
    U.assert_close(full["what"][:, sel], ref["outs"]["what"], atol=1e-4, name="what sample")
    assert torch.equal(full["num_step_per_sample"][sel], ref["num_step_per_sample"])


# ----------------------------------------------------------------------------------------------------------
# the Python class surface (model.py / mnist_model.py / scripts/multi_mnist.py:82-100)
# ----------------------------------------------------------------------------------------------------------
def test_air_on_mnist_surface():
    B, T = 32, 3
    img, nums = O.synthetic_multi_mnist(B, 50, 50, seed=5)
    x, y = img.to(DEV), nums.to(DEV)
    n_hiddens = [256, 256]
    model = air.AIRonMNIST(x, y, max_steps=T, explore_eps=1e-3, inpt_encoder_hidden=n_hiddens,
                           glimpse_encoder_hidden=n_hiddens, glimpse_decoder_hidden=n_hiddens,
                           transform_estimator_hidden=n_hiddens, steps_pred_hidden=[128, 64],
                           baseline_hidden=[256, 128], transform_var_bias=.5, step_bias=.75, output_multiplier=.5)
    assert model.params.numel() == 1782525
    for name, shape in [("canvas", (T, B, 50, 50)), ("glimpse", (T, B, 20, 20)), ("what", (T, B, 50)),
                        ("what_loc", (T, B, 50)), ("what_scale", (T, B, 50)), ("where", (T, B, 4)),
                        ("where_loc", (T, B, 4)), ("where_scale", (T, B, 4)), ("presence_prob", (T, B, 1)),
                        ("presence", (T, B, 1)), ("final_canvas", (B, 50, 50)), ("num_step_per_sample", (B,)),
                        ("gt_num_steps", (B,))]:
        assert tuple(getattr(model, name).shape) == shape, name
    assert tuple(model.num_steps_distrib.prob().shape) == (B, T + 1)
    num_steps_prior = dict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3)
    pr = dict(loc=0., scale=1.)
    train_op, global_step = model.train_step(1e-4, 0., pr, pr, pr, num_steps_prior)
    noise = model.cell.draw_noise(B, T, generator=torch.Generator(device=DEV).manual_seed(3))
    params_before = model.params.detach().cpu().clone()
    bview = {k: v.detach().cpu().clone() for k, v in model.baseline_module.views.items()}
    train_op(noise=noise)
    assert global_step() == 1
    assert not torch.equal(params_before, model.params.detach().cpu()), "train_op must update the parameters"
    # cross-check every exposed loss attribute against the oracle, with the weights the step was evaluated at
    ocfg = U.oracle_cfg(**U.SCRIPT)
    params = O.unflatten_params(ocfg, params_before)
    bl = model.baseline.detach().cpu()
    ref = O.forward(ocfg, O.PriorConfig(), params, img, *(n.cpu() for n in noise), global_step=0, baseline=bl)
    assert any(not torch.equal(v, model.baseline_module.views[k].cpu()) for k, v in bview.items()), \
        "train_op must also run the baseline's own optimiser (model.py:362-367)"
    bl_ref = O.baseline_mlp(bview, 2, img, ref["outs"]["what"], ref["outs"]["where"], ref["outs"]["presence"],
                            ref["final_h"], ref["final_c"])
    U.assert_close(bl, bl_ref, atol=2e-4, rtol=1e-4, name="baseline")
    for attr, key in [("rec_loss", "rec_loss"), ("kl_num_steps", "kl_num_steps"), ("kl_what", "kl_what"),
                      ("kl_where", "kl_where"), ("reinforce_loss", "reinforce_loss"), ("opt_loss", "opt_loss")]:
        U.assert_close(getattr(model, attr).cpu(), ref[key].float(), atol=2e-4, rtol=2e-4, name=attr)
    U.assert_close(model.loss.value.cpu(), ref["loss"].float(), rtol=1e-4, name="loss.value")
    U.assert_close(model.loss.per_sample.cpu(), ref["loss_per_sample"], rtol=1e-4, name="loss.per_sample")
    U.assert_close(model.prior_loss.value.cpu(), ref["prior_loss"].float(), rtol=1e-4, name="prior_loss.value")
    assert tuple(model.importance_weight.shape) == (B, B)
    U.assert_close(model.importance_weight.cpu(), ref["importance_weight"], atol=1e-2, rtol=1e-4, name="imp weight")
    assert 0.0 <= float(model.num_step_accuracy) <= 1.0


def test_cpu_tensors_are_rejected():
    with pytest.raises(air.AirError):
        AF.stn_read(torch.rand(1, 4, 4), torch.rand(1, 4), (2, 2))
    with pytest.raises(air.AirError):
        air.AIRonMNIST(torch.rand(2, 50, 50), None, max_steps=3)


def test_bad_config_is_an_error():
    with pytest.raises(air.AirError):
        air.Engine(air.CellConfig(), 4, air._lib.AIR_MAX_STEPS + 1)
    with pytest.raises(air.AirError):
        air.Engine(air.CellConfig(output_std=0.0), 4, 3)


# ----------------------------------------------------------------------------------------------------------
# tensor-core engine (AIR_PREC_TC_SPLIT): tcgen05 fp16x2-split GEMMs, same tolerances as the fp32 engine
# ----------------------------------------------------------------------------------------------------------
TC = air.AIR_PREC_TC_SPLIT


@pytest.mark.parametrize("M,K,N,act", [(64, 2500, 256, 1), (192, 256, 8, 0), (130, 50, 256, 1), (7, 17, 1, 0),
                                       (257, 400, 100, 0), (1, 3, 5, 1), (300, 256, 1024, 0), (4096, 256, 256, 1),
                                       (128, 64, 64, 0), (129, 65, 65, 0)])
def test_linear_tc(M, K, N, act):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    ref = (x.double() @ w.double() + b.double())
    ref32 = x @ w + b
    if act:
        ref, ref32 = O.elu(ref), O.elu(ref32)
    out = AF.linear(x.to(DEV), w.to(DEV), b.to(DEV), act, precision=TC).cpu()
    err = float((out.double() - ref).abs().max())
    err32 = float((ref32.double() - ref).abs().max())
    print(f"tc linear M={M} K={K} N={N}: max err vs fp64 {err:.3e} (fp32 matmul itself: {err32:.3e})")
    # the tensor core truncates its fp32 accumulator once per MMA: a bias that grows with K / 16
    tol = 2e-5 if K <= 512 else 1e-4
    U.assert_close(out, ref.float(), atol=tol, rtol=tol, name="linear_tc")


def test_linear_tc_structured_operands():
    """Catches layout mistakes (swizzle, K-advance, hi/lo plane mix-ups) that random data can hide: one-hot rows and
    columns, and values whose lo halves matter."""
    M, K, N = 256, 192, 128
    x = torch.zeros(M, K)
    x[torch.arange(M), torch.arange(M) % K] = 1.0 + torch.arange(M) * 2.0 ** -13      # needs the lo half
    w = (torch.arange(K * N, dtype=torch.float32).reshape(K, N) % 251 - 125) / 1000.0 + 2.0 ** -14
    ref = x.double() @ w.double()
    out = AF.linear(x.to(DEV), w.to(DEV), None, 0, precision=TC).cpu()
    U.assert_close(out, ref.float(), atol=1e-6, rtol=2e-6, name="linear_tc structured")


def test_linear_tc_range_overflow_is_reported():
    x = torch.full((8, 64), 1.0e5)            # > 65504: not representable as an fp16 hi half
    w = torch.ones(64, 64)
    with pytest.raises(air.AirError):
        AF.linear(x.to(DEV), w.to(DEV), None, 0, precision=TC)


def test_forward_tc_script_config_b64():
    _check_forward(U.oracle_cfg(**U.SCRIPT), 64, O.PriorConfig(), seed=0, global_step=20000, precision=TC)


def test_forward_tc_ragged_and_odd():
    _check_forward(U.oracle_cfg(**U.SCRIPT), 3, O.PriorConfig(), seed=31, global_step=12000, precision=TC)
    _check_forward(U.oracle_cfg(**U.SCRIPT), 129, O.PriorConfig(), seed=32, global_step=12000, precision=TC)
    _check_forward(U.oracle_cfg(**U.TINY), 10, O.PriorConfig(), seed=2, global_step=5000, weight_gain=2.0, precision=TC)


def test_forward_tc_config_d_small_batch():
    _check_forward(U.oracle_cfg(**U.CONFIG_D), 12, O.PriorConfig(), seed=3, global_step=30000, precision=TC)


def test_forward_tc_trained_like_weights():
    _check_forward(U.oracle_cfg(**U.SCRIPT), 48, O.PriorConfig(), seed=8, global_step=60000, weight_gain=2.5,
                   noise_floor=True, precision=TC)


def test_forward_tc_full_size_sample_and_engine_agreement():
    """B = 4096: the two engines agree with each other everywhere, and a strided sample agrees with the oracle."""
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B = 4096
    pc = O.PriorConfig()
    params, img, nums, noise = U.make_problem(ocfg, B, seed=21)
    a = U.run_cuda(ocfg, params, img, noise, pc, 20000, precision=air.AIR_PREC_FP32)
    b = U.run_cuda(ocfg, params, img, noise, pc, 20000, precision=TC)
    same = (a["presence"] == b["presence"]).reshape(ocfg.T, B).all(0)
    assert same.float().mean() > 0.999
    for k in ("what", "where", "glimpse", "presence_prob"):
        U.assert_close(b[k], a[k], atol=1e-4, rtol=1e-4, name="tc vs fp32 " + k)
    # per-sample losses change sign across the batch: "relative 1e-4" is taken against the batch's mean magnitude
    scale = float(a["loss_per_sample"].abs().mean())
    U.assert_close(b["loss_per_sample"][same], a["loss_per_sample"][same], atol=1e-4 * scale, rtol=1e-4,
                   name="tc vs fp32 loss")
    err = (b["loss_per_sample"][same] - a["loss_per_sample"][same]).abs()
    print(f"tc vs fp32 per-sample loss: max {float(err.max()):.3e} mean {float(err.mean()):.3e} (mean |loss| {scale:.1f})")
    sel = torch.arange(0, B, 64)
    ref = O.forward(ocfg, pc, params, img[sel], *(n[:, sel] for n in noise), global_step=20000)
    U.assert_close(b["what"][:, sel], ref["outs"]["what"], atol=1e-4, name="what sample")
    ok = same[sel]
    U.assert_close(b["loss_per_sample"][sel][ok], ref["loss_per_sample"][ok], atol=1e-4 * scale, rtol=1e-4,
                   name="loss sample")


# ----------------------------------------------------------------------------------------------------------
# end-to-end entry points with host buffers (what bench.py's e2e arm calls)
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", [air.AIR_PREC_FP32, TC])
def test_forward_host_entry_points_match_device_call(precision):
    from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B = 96
    pc = O.PriorConfig()
    params, _, _, noise = U.make_problem(ocfg, B, seed=41)
    u8 = torch.from_numpy(synthetic_multi_mnist_u8(B, 50, 50, seed=41)[0]).contiguous()
    img = u8.to(torch.float32) / 255.0                       # load_data (data.py:116)
    ref = O.forward(ocfg, pc, params, img, *noise, global_step=20000)
    eng = air.Engine(U.cell_cfg(ocfg, precision), B, ocfg.T, device=DEV)
    flat = O.flatten_params(ocfg, params).to(DEV)
    pr = U.prior_struct(pc, 20000)
    dev_out = eng.forward(flat, img.to(DEV), *(n.to(DEV).contiguous() for n in noise), pr)
    dev_loss = dev_out["loss_per_sample"].cpu().clone()
    dev_scal = dev_out["scalars"].cpu().clone()
    scal, lps = torch.empty(air._lib.AIR_N_SCALARS).pin_memory(), torch.empty(B).pin_memory()
    pinned = [n.contiguous().pin_memory() for n in noise]
    eng.forward_host(flat, img.contiguous().pin_memory(), *pinned, pr, scal, lps)
    assert torch.equal(lps, dev_loss) and torch.equal(scal, dev_scal)
    scal2, lps2 = torch.empty_like(scal).pin_memory(), torch.empty_like(lps).pin_memory()
    eng.forward_host_u8(flat, u8.pin_memory(), *pinned, pr, scal2, lps2)        # /255 on the device
    assert torch.equal(lps2, dev_loss) and torch.equal(scal2, dev_scal)
    eng.check_range()
    scale = float(ref["loss_per_sample"].abs().mean())
    U.assert_close(lps2, ref["loss_per_sample"], atol=1e-4 * scale, rtol=1e-4, name="e2e loss vs oracle")
    eng.close()


# ----------------------------------------------------------------------------------------------------------
# the tensor-core engine has two schedules: fused chains (chain_tc.cuh, default) and one launch per layer
# (linear_tc.cuh; AIR_NO_CHAIN=1, also the fall-back for hidden widths > 256).  Both must hold parity.
# ----------------------------------------------------------------------------------------------------------
def test_forward_tc_per_layer_schedule(monkeypatch):
    monkeypatch.setenv("AIR_NO_CHAIN", "1")
    _check_forward(U.oracle_cfg(**U.SCRIPT), 64, O.PriorConfig(), seed=0, global_step=20000, precision=TC)
    _check_forward(U.oracle_cfg(**U.TINY), 10, O.PriorConfig(), seed=2, global_step=5000, weight_gain=2.0, precision=TC)


def test_forward_tc_without_cluster_lstm(monkeypatch):
    """fused chains on, cluster LSTM kernel off: per-step recurrent GEMM + gate kernel feeding the chains"""
    monkeypatch.setenv("AIR_NO_LSTM_CLUSTER", "1")
    _check_forward(U.oracle_cfg(**U.SCRIPT), 130, O.PriorConfig(), seed=4, global_step=20000, precision=TC)


def test_forward_tc_chain_deep_and_ragged_widths():
    """3-4 hidden layers of odd widths, na not a multiple of 16, nh > 256 (chunked A operand), 130 rows (tile tail)."""
    cfg = dict(H=20, W=24, h=9, w=7, T=4, na=21, nh=272, enc_hidden=(96, 40), glenc_hidden=(200, 72, 256),
               dec_hidden=(33, 256, 100, 64), where_hidden=(256, 17, 48), steps_hidden=(24,))
    _check_forward(U.oracle_cfg(**cfg), 130, O.PriorConfig(), seed=5, global_step=15000, precision=TC)


def test_forward_tc_wide_hidden_falls_back_to_per_layer():
    cfg = dict(U.SCRIPT, where_hidden=(320, 64), T=2)
    _check_forward(U.oracle_cfg(**cfg), 20, O.PriorConfig(), seed=7, global_step=15000, precision=TC, noise_floor=True)


def test_resident_dataset_forward_matches_host_fed_forward():
    """the uint8 dataset stays in HBM; the minibatch gather + /255 inside the library gives exactly the
    forward pass of the same canvases fed as float32 (both engines), and air_gather_u8 is exact."""
    from attend_infer_repeat_b200.data import ResidentDataset, synthetic_multi_mnist_u8
    imgs_u8, nums_u8 = synthetic_multi_mnist_u8(200, 50, 50, seed=2)
    ds = ResidentDataset(imgs_u8, nums_u8, device=DEV, seed=1)
    B, T = 48, 3
    idx = ds.next_indices(B)
    assert idx.dtype == torch.int32 and int(idx.min()) >= 0 and int(idx.max()) < 200
    img, nums = ds.gather(idx)
    ref = torch.from_numpy(imgs_u8.astype("float32") / 255.)[idx.cpu().long()]
    assert torch.equal(img.cpu(), ref)
    assert torch.equal(nums.cpu(), torch.from_numpy(nums_u8.astype("float32"))[:, idx.cpu().long()])
    ocfg = U.oracle_cfg(**U.SCRIPT)
    params = O.flatten_params(ocfg, O.init_params(ocfg, 0)).to(DEV)
    noise = tuple(n.to(DEV).contiguous() for n in O.make_noise(ocfg, B, 4))
    pr = U.prior_struct(O.PriorConfig(), 20000)
    for prec in (air.AIR_PREC_FP32, air.AIR_PREC_TC_SPLIT):
        eng = air.Engine(U.cell_cfg(ocfg, prec), B, T, device=DEV)
        a = {k: v.clone() for k, v in eng.forward(params, img, *noise, pr).items() if v is not None}
        img_out = torch.empty(B, 50, 50, device=DEV)
        b = eng.forward_dataset_u8(params, ds.imgs, idx, *noise, pr, img_out=img_out)
        torch.cuda.synchronize()
        assert torch.equal(img_out, img)
        for k in ("canvas", "what", "where", "presence", "loss_per_sample", "scalars"):
            assert torch.equal(a[k], b[k]), (prec, k)
        eng.close()


def _philox_numpy(n, seed, stream):
    """Philox4x32-10 (Salmon et al. 2011) with counter (i, 0, stream, 0)... restated in numpy: the spec the kernel follows."""
    import numpy as np
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    q = np.arange((n + 3) // 4, dtype=np.uint64)
    c = [q & 0xFFFFFFFF, q >> np.uint64(32), np.full_like(q, stream), np.zeros_like(q)]
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & 0xFFFFFFFF, p1 >> np.uint64(32), p1 & 0xFFFFFFFF
        c = [(hi1 ^ c[1] ^ np.uint64(k0)) & 0xFFFFFFFF, lo1, (hi0 ^ c[3] ^ np.uint64(k1)) & 0xFFFFFFFF, lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c


def test_in_library_noise_is_reproducible_and_well_distributed():
    """air_draw_noise (Philox4x32-10 + Box-Muller): deterministic in the seed, different across seeds and tensors, equal
    to a numpy restatement of the generator (uniforms bit-exact, normals to 1e-5), statistically N(0,1) / U[0,1); a forward
    pass fed the drawn tensors explicitly equals the *_rng end-to-end call."""
    import numpy as np
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B, T = 512, 3
    eng = air.Engine(U.cell_cfg(ocfg, air.AIR_PREC_FP32), B, T, device=DEV)
    a, b, c = eng.draw_noise(7), eng.draw_noise(7), eng.draw_noise(8)
    for x, y, z in zip(a, b, c):
        assert torch.equal(x, y) and not torch.equal(x, z)
    ew, ea, u = (t.double().cpu() for t in a)
    assert not torch.equal(ew.reshape(-1)[:1000], ea.reshape(-1)[:1000])
    # numpy restatement: u_pres is stream 2 (24-bit uniforms, exact); eps_where stream 0 through Box-Muller
    r = _philox_numpy(u.numel(), 7, 2)
    u_ref = np.stack([(x >> np.uint64(8)).astype(np.float64) / 16777216.0 for x in r], 1).reshape(-1)[:u.numel()]
    assert np.array_equal(u.reshape(-1).numpy(), u_ref)
    r = _philox_numpy(ew.numel(), 7, 0)
    uu = [((x >> np.uint64(8)).astype(np.float64) + 1.0) / 16777216.0 for x in r]
    r0, r1 = np.sqrt(-2 * np.log(uu[0])), np.sqrt(-2 * np.log(uu[2]))
    n_ref = np.stack([r0 * np.cos(2 * np.pi * uu[1]), r0 * np.sin(2 * np.pi * uu[1]), r1 * np.cos(2 * np.pi * uu[3]),
                      r1 * np.sin(2 * np.pi * uu[3])], 1).reshape(-1)[:ew.numel()]
    assert float(np.abs(ew.reshape(-1).numpy() - n_ref).max()) < 1e-5
    for x in (ew, ea):   # sample moments within 6 standard errors of N(0,1)
        n = x.numel()
        assert abs(float(x.mean())) < 6.0 / n ** 0.5 and abs(float(x.var()) - 1.0) < 6.0 * (2.0 / n) ** 0.5
        assert abs(float((x ** 3).mean())) < 6.0 * (15.0 / n) ** 0.5
        assert abs(float((x ** 4).mean()) - 3.0) < 6.0 * (96.0 / n) ** 0.5
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0 and abs(float(u.mean()) - 0.5) < 6.0 / (12 * u.numel()) ** 0.5
    # the end-to-end call with the same seed reproduces an explicit-noise forward pass
    from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8
    u8 = torch.from_numpy(synthetic_multi_mnist_u8(B, 50, 50, seed=4)[0])
    params = O.flatten_params(ocfg, O.init_params(ocfg, 0)).to(DEV)
    pr = U.prior_struct(O.PriorConfig(), 20000)
    img = (u8.float() / 255.0).to(DEV)
    ref = {k: v.clone() for k, v in eng.forward(params, img, *a, pr).items() if v is not None}
    sc, lps = torch.empty(16).pin_memory(), torch.empty(B).pin_memory()
    eng.forward_host_u8_rng(params, u8.pin_memory(), 7, pr, sc, lps)
    assert torch.equal(lps, ref["loss_per_sample"].cpu()) and torch.equal(sc, ref["scalars"].cpu())
    eng.close()


@pytest.mark.parametrize("precision", [air.AIR_PREC_FP32, air.AIR_PREC_TC_SPLIT])
def test_double_buffered_host_feed_equals_synchronous_calls(precision):
    """air_feed_host_u8 / air_forward_fed_u8_rng / air_feed_wait (the copy of batch i+1 overlaps the pass over batch i):
    every batch's scalars and per-sample loss are bit-identical to the synchronous air_forward_host_u8_rng call on the same
    batch and seed -- seven distinct batches through two slots, plus the error paths."""
    from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B, T, n = 192, 3, 7
    eng = air.Engine(U.cell_cfg(ocfg, precision), B, T, device=DEV)
    params = O.flatten_params(ocfg, O.init_params(ocfg, 0)).to(DEV)
    pr = U.prior_struct(O.PriorConfig(), 20000)
    batches = [torch.from_numpy(synthetic_multi_mnist_u8(B, 50, 50, seed=10 + i)[0]).pin_memory() for i in range(n)]
    sc, lps = torch.empty(16).pin_memory(), torch.empty(B).pin_memory()
    with pytest.raises(air.AirError):
        eng.feed_wait(0)                                      # the feed was never used
    ref = []
    for i, b in enumerate(batches):
        eng.forward_host_u8_rng(params, b, 100 + i, pr, sc, lps)
        ref.append((sc.clone(), lps.clone()))
    assert not torch.equal(ref[0][1], ref[1][1])
    got = [(s.clone(), l.clone()) for s, l in eng.stream_host_u8(params, batches, pr, seed0=100)]
    assert len(got) == n
    for (s0, l0), (s1, l1) in zip(ref, got):
        assert torch.equal(s0, s1) and torch.equal(l0, l1)
    assert list(eng.stream_host_u8(params, [], pr)) == []
    with pytest.raises(air.AirError):
        eng.forward_fed_u8_rng(params, 1, 0, pr, sc, lps)     # nothing pending in the slot
    with pytest.raises(air.AirError):
        eng.feed_host_u8(2, batches[0])                       # only slots 0 and 1 exist
    eng.close()


@pytest.mark.parametrize("precision", [air.AIR_PREC_FP32, air.AIR_PREC_TC_SPLIT])
def test_iwae_bound_matches_oracle(precision):
    """BASELINE.json configs[4] shapes in small: K = 5 particles per canvas as consecutive rows of one forward pass;
    log w, the per-canvas bound log(1/K sum_k w_k) and its batch mean against the float64 oracle restatement (the
    reference has no IWAE: parity unpinned), and the bound is never below the same pass's mean ELBO estimate."""
    ocfg = U.oracle_cfg(**U.SCRIPT)
    pc = O.PriorConfig()
    n, K = 12, 5
    params, img, nums, _ = U.make_problem(ocfg, n, seed=17)
    img_k = img.repeat_interleave(K, 0)
    noise = O.make_noise(ocfg, n * K, seed=18)
    ref = O.forward(ocfg, pc, params, img_k, *noise, global_step=20000)
    ref_iw = O.iwae_bound(ocfg, pc, ref, K, global_step=20000)
    eng = air.Engine(U.cell_cfg(ocfg, precision), n * K, ocfg.T, device=DEV)
    pr = U.prior_struct(pc, 20000)
    eng.forward(O.flatten_params(ocfg, params).to(DEV), img_k.to(DEV).contiguous(),
                *(t.to(DEV).contiguous() for t in noise), pr)
    mean, bound, log_w = eng.iwae_bound(K, pr)
    torch.cuda.synchronize()
    eng.check_range()
    U.assert_close(log_w.cpu(), ref_iw["log_w"], atol=1e-3, rtol=1e-4, name="log_w")
    U.assert_close(bound.cpu(), ref_iw["bound_per_canvas"], atol=1e-3, rtol=1e-4, name="bound per canvas")
    U.assert_close(mean.cpu(), ref_iw["bound"], atol=1e-3, rtol=1e-4, name="bound")
    # Jensen: log mean_k w >= mean_k log w, per canvas
    assert bool((bound.cpu() >= log_w.cpu().reshape(n, K).mean(1) - 1e-3).all())
    with pytest.raises(air.AirError):
        eng.iwae_bound(7, pr)
    eng.close()
