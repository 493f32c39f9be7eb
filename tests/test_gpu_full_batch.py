"""Parity on the configuration the metric is quoted on (BASELINE.json configs[1]: 50x50, 20x20, T=3, B=4096), tensor-core
engine, through the C ABI, against the CPU oracle on the WHOLE batch: all ten AIRCell outputs, every per-sample ELBO
vector, and the training gradient.

The conditioning rule (explicit, like the presence near-tie rule): the inverse transformer divides by the sampled scales
(modules.py:100-102), so the canvas of a step whose |s_x| or |s_y| is small is an ill-conditioned function of the where
code -- a relative perturbation eps of `where` moves the canvas coordinates by eps / |s|.  The fp32 oracle is itself
only defined up to that amplification of its own rounding.  Measured at B = 4096 (tools/cond_probe.py, this batch): the
`where` code of either engine sits 0.7-1.0e-6 from the fp32 oracle, which itself sits 0.6e-6 from its float64 run; the
canvas error is below 0.8e-4 on every canvas with |s|_min >= 1e-2 and reaches 1.6e-4 in the bin [3e-3, 1e-2), where the
fp32 oracle is 1.1e-4 away from float64.  Therefore:
  * every output that is NOT downstream of 1 / s (glimpse, what*, where*, presence_prob, num_steps_posterior, the KL
    vectors) is held to 1e-4 on every element of every canvas;
  * presence / num_step_per_sample are bit-exact wherever |u - p| > 1e-5 at every step (near-ties counted and bounded);
  * canvas, rec_loss_per_sample and loss_per_sample are held to 1e-4 (absolute + relative; losses relative to the batch's
    mean magnitude) on every canvas whose painted steps all have min(|s_x|, |s_y|) >= TAU; the canvases below TAU are
    counted, their number is bounded, and their error is bounded by the amplification the rule predicts.
"""
import pytest
import torch

import attend_infer_repeat_b200 as air
from oracle import air_oracle as O
from tests import util as U
from tests.test_gpu_backward import compare, cuda_grads, oracle_grads, well_conditioned

pytestmark = pytest.mark.gpu

TC = air.AIR_PREC_TC_SPLIT
TAU = 0.02          # conditioning threshold on the sampled scales of painted steps
B_FULL = 4096


def _ill_conditioned(ref, T, B):
    """[B] bool: some PAINTED step (presence 1) of the canvas has |s_x| or |s_y| < TAU; and the smallest such |s|."""
    where = ref["outs"]["where"].reshape(T, B, 4)
    pres = ref["outs"]["presence"].reshape(T, B)
    s_min = torch.minimum(where[..., 0].abs(), where[..., 2].abs())
    s_min = torch.where(pres > 0, s_min, torch.full_like(s_min, 1e9)).min(0).values      # [B]
    return s_min < TAU, s_min


@pytest.fixture(scope="module")
def full_batch():
    ocfg = U.oracle_cfg(**U.SCRIPT)
    pc = O.PriorConfig()
    params, img, nums, noise = U.make_problem(ocfg, B_FULL, seed=4096)
    with torch.no_grad():
        ref = O.forward(ocfg, pc, params, img, *noise, global_step=20000)
    out = U.run_cuda(ocfg, params, img, noise, pc, 20000, precision=TC)
    return ocfg, pc, params, img, noise, ref, out


def test_full_batch_cell_outputs_match_oracle_everywhere(full_batch):
    ocfg, pc, params, img, noise, ref, out = full_batch
    for k in ("glimpse", "what", "what_loc", "what_scale", "where", "where_loc", "where_scale", "presence_prob"):
        U.assert_close(out[k], ref["outs"][k], atol=1e-4, rtol=1e-4, name=k)
    U.assert_close(out["final_h"], ref["final_h"], atol=1e-4, name="final_h")
    U.assert_close(out["final_c"], ref["final_c"], atol=1e-4, name="final_c")
    U.assert_close(out["num_steps_posterior"], ref["num_steps_posterior"], atol=1e-5, rtol=1e-4, name="q(n)")
    U.assert_close(out["prior_step_weight"], ref["prior_step_weight"], atol=1e-5, rtol=1e-4, name="step weight")
    for k in ("kl_what_per_sample", "kl_where_per_sample", "kl_num_steps_per_sample"):
        scale = max(1.0, float(ref[k].abs().mean()))
        U.assert_close(out[k], ref[k], atol=1e-4 * scale, rtol=1e-4, name=k)


def test_full_batch_presence_is_bit_exact_away_from_ties(full_batch):
    ocfg, pc, params, img, noise, ref, out = full_batch
    T, B = ocfg.T, B_FULL
    bad, unsafe = U.presence_mismatches(out["presence"], ref["outs"]["presence_prob"].reshape(T, B), noise[2].reshape(T, B))
    assert bad == 0, f"{bad} presence mismatches away from ties"
    assert unsafe <= 8, f"{unsafe} draws within 1e-5 of a tie in {T * B}"      # expected 2 * 1e-5 * T * B = 0.25
    same = (out["presence"].reshape(T, B) == ref["outs"]["presence"].reshape(T, B)).all(0)
    assert int((~same).sum()) <= unsafe
    assert torch.equal(out["num_step_per_sample"][same], ref["num_step_per_sample"][same])


def test_full_batch_canvas_and_losses_under_the_conditioning_rule(full_batch):
    ocfg, pc, params, img, noise, ref, out = full_batch
    T, B = ocfg.T, B_FULL
    same = (out["presence"].reshape(T, B) == ref["outs"]["presence"].reshape(T, B)).all(0)
    ill, s_min = _ill_conditioned(ref, T, B)
    n_ill = int(ill.sum())
    # with untrained weights s ~ loc + softplus(.) * N(0, 1) is broad: P(|s| < TAU) ~ 1.4 % per axis and painted step,
    # ~5 % of the canvases (200 of 4096 on this seed); bound the count at 8 %
    assert n_ill <= (8 * B) // 100, f"{n_ill} canvases below TAU = {TAU}"
    good = same & ~ill
    assert int(good.sum()) >= int(0.9 * B)
    canvas, canvas_ref = out["canvas"].reshape(T, B, -1), ref["canvas"].reshape(T, B, -1)
    U.assert_close(canvas[:, good], canvas_ref[:, good], atol=1e-4, rtol=1e-4, name="canvas (|s| >= TAU)")
    for k in ("rec_loss_per_sample", "loss_per_sample", "num_steps_log_prob"):
        if k in ref:
            scale = max(1.0, float(ref[k][good].abs().mean()))
            U.assert_close(out[k][good], ref[k][good], atol=1e-4 * scale, rtol=1e-4, name=k + " (|s| >= TAU)")
    # below TAU: the error the rule predicts is the where error (~1e-6 on either engine, the fp32 oracle's own distance to
    # float64) amplified by 1 / |s| into glimpse coordinates: 1e-4 + 2e-6 / |s|_min
    idx = torch.nonzero(ill & same).reshape(-1)
    worst = 0.0
    for b in idx.tolist():
        err = float((canvas[:, b] - canvas_ref[:, b]).abs().max())
        bound = 1e-4 + 2e-6 / max(float(s_min[b]), 1e-6)
        worst = max(worst, err / bound)
        assert err <= bound, f"canvas {b}: |s|_min {float(s_min[b]):.2e}, error {err:.2e} > predicted bound {bound:.2e}"
    print(f"B={B}: {n_ill} canvases below TAU={TAU} (worst error / predicted bound {worst:.2f}); "
          f"{int((~same).sum())} near-tie canvases; canvas max err on the rest "
          f"{float((canvas[:, good] - canvas_ref[:, good]).abs().max()):.2e}")


def test_full_batch_scalars_match_oracle(full_batch):
    ocfg, pc, params, img, noise, ref, out = full_batch
    T, B = ocfg.T, B_FULL
    same = (out["presence"].reshape(T, B) == ref["outs"]["presence"].reshape(T, B)).all(0)
    if not bool(same.all()):
        pytest.skip("a near-tie presence draw flipped: batch means are compared in the per-sample tests")
    s, idx = out["scalars"], air._lib.SCALAR_INDEX
    for name in ("rec_loss", "kl_num_steps", "kl_what", "kl_where", "prior_loss", "loss", "num_step"):
        U.assert_close(s[idx[name]], ref[name].float(), atol=1e-4, rtol=1e-4, name="scalar " + name)
    U.assert_close(-s[idx["loss"]], ref["elbo"].float(), atol=0, rtol=1e-4, name="ELBO")


def test_full_batch_backward_matches_float64_oracle_autograd():
    """air_backward at B = 4096 on the tensor-core engine (tcgen05 forward with kept activations, bf16 hi/lo gradient
    GEMMs) against autograd on the oracle in float64.  Draws below the backward conditioning threshold (|s| < 0.05:
    1 / s^2 factors) are replaced by the posterior mean exactly as in tests/test_gpu_backward.py."""
    ocfg = U.oracle_cfg(**U.SCRIPT)
    pc = O.PriorConfig()
    params, img, nums, noise = U.make_problem(ocfg, B_FULL, seed=409)
    noise, n_fixed = well_conditioned(ocfg, pc, params, img, noise)
    res_o, g_ref = oracle_grads(ocfg, pc, params, img, noise, 20000, dtype=torch.float64)
    res_c, g = cuda_grads(ocfg, pc, params, img, noise, 20000, precision=TC)
    T = ocfg.T
    same = (res_c["presence"].reshape(T, B_FULL) == res_o["outs"]["presence"].detach().float().reshape(T, B_FULL)).all(0)
    assert int((~same).sum()) <= 2, f"{int((~same).sum())} presence flips"
    worst = compare(ocfg, g, g_ref, rel=4e-4 if not bool(same.all()) else 2e-4)
    print(f"B={B_FULL} backward vs float64 oracle ({n_fixed} ill-conditioned draws replaced): worst {worst[1]} "
          f"({worst[0]:.2f} of tolerance)")
