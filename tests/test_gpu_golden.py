"""The CUDA library against the golden vectors produced by the reference's own prior.py / model.py source
(tools/make_golden.py): stand-alone prior kernels, and the KL / step-count / REINFORCE part of the fused ELBO kernel
through air_prior_terms + air_elbo_scalars."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200 import _lib
from attend_infer_repeat_b200 import functional as AF
from tests.test_oracle_golden import L, P, _case

pytestmark = pytest.mark.gpu
DEV = "cuda"
D32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=DEV).contiguous()


def test_prior_kernels_vs_reference_vectors():
    assert np.allclose(AF.geometric_prior(.75, 10, device=DEV).cpu().numpy(), P["geom_075_10"], rtol=1e-6)
    assert np.allclose(AF.geometric_prior(.005, 3, device=DEV).cpu().numpy(), P["geom_0005_3"], rtol=1e-6)
    assert np.array_equal(AF.bernoulli_to_modified_geometric(D32(P["b2mg_in"])).cpu().numpy(), P["b2mg_out"])
    assert np.array_equal(AF.bernoulli_to_modified_geometric(D32(P["b2mg5_in"])).cpu().numpy(), P["b2mg5_out"])
    kl = AF.tabular_kl(D32(P["tkl_p"]), torch.as_tensor(P["tkl_q"], dtype=torch.float32)).cpu().numpy()
    assert np.allclose(kl, P["tkl_out"], rtol=1e-6, atol=1e-9) and kl[0, 0] == 0.0
    d = air.NumStepsDistribution(D32(P["b2mg_in"]))
    assert np.array_equal(d.prob().cpu().numpy(), P["nsd_joint"])
    assert np.array_equal(d.prob(D32(P["nsd_samples"])).cpu().numpy(), P["nsd_prob"])
    assert np.allclose(d.log_prob(D32(P["nsd_samples"])).cpu().numpy(), P["nsd_log_prob"], rtol=1e-6, atol=1e-6)


def test_anneal_weight_vs_reference_vectors():
    for i, s in enumerate(L["anneal_steps"]):
        assert AF.anneal_weight(1. - 1e-15, 1e-7, "exp", int(s), 1e5, 1e3, 1e4) == \
            pytest.approx(float(L["anneal_exp"][i]), rel=1e-13)
        assert AF.anneal_weight(.9, .1, "linear", int(s), 1e5, 1e3, 1.) == \
            pytest.approx(float(L["anneal_linear"][i]), rel=1e-13)


@pytest.mark.parametrize("i", range(int(L["n_cases"])))
def test_fused_elbo_kernel_vs_reference_prior_loss_and_reinforce(i):
    pc, gstep, g = _case(i)
    T, B, na = g["what_loc"].shape
    f = lambda *s: torch.zeros(*s, device=DEV, dtype=torch.float32)
    bufs = dict(num_steps_posterior=f(B, T + 1), num_step_per_sample=f(B), prior_step_weight=f(T, B),
                rec_loss_per_sample=f(B), kl_num_steps_per_sample=f(B), kl_what_per_sample=f(B),
                kl_where_per_sample=f(B), loss_per_sample=f(B), num_steps_log_prob=f(B), scalars=f(_lib.AIR_N_SCALARS))
    outs = _lib.air_outputs()
    for k, v in bufs.items():
        setattr(outs, k, v.data_ptr())
    success = float(g["success_prob"])
    pr = air.make_prior(dict(loc=pc.what_loc, scale=pc.what_scale),
                        dict(loc=pc.where_scale_loc, scale=pc.where_scale_scale),
                        dict(scale=pc.where_shift_scale) if pc.where_shift_loc is None else
                        dict(loc=pc.where_shift_loc, scale=pc.where_shift_scale),
                        success, pc.steps_anneal is not None, pc.steps_weight, pc.analytic, True, True)
    t = {k: D32(g[k]) for k in ("what_loc", "what_scale", "where_loc", "where_scale", "presence_prob", "presence")}
    lib = _lib.lib()
    st = _lib.current_stream_ptr()
    _lib.check(lib.air_prior_terms(B, T, na, _lib.ptr(t["what_loc"]), _lib.ptr(t["what_scale"]),
                                   _lib.ptr(t["where_loc"]), _lib.ptr(t["where_scale"]), _lib.ptr(t["presence_prob"]),
                                   _lib.ptr(t["presence"]), C.byref(pr), C.byref(outs), st), "air_prior_terms")
    torch.cuda.synchronize()
    close = lambda a, b, rt=3e-6: np.allclose(a.detach().cpu().numpy(), b, rtol=rt, atol=2e-6)
    assert np.array_equal(bufs["num_steps_posterior"].cpu().numpy(), g["posterior"])
    assert close(bufs["prior_step_weight"], g["step_weight"])
    assert close(bufs["kl_num_steps_per_sample"], g["kl_num_steps_ps"])
    assert close(bufs["loss_per_sample"], g["prior_per_sample"])          # rec = 0 here -> loss == weighted prior terms
    assert close(bufs["num_steps_log_prob"], g["log_prob"])
    assert np.array_equal(bufs["num_step_per_sample"].cpu().numpy(), g["presence"].sum(0).reshape(-1))
    # batch means + REINFORCE with the golden reconstruction losses and baseline
    bufs["rec_loss_per_sample"].copy_(D32(g["rec"]))
    eng_like = lambda baseline: _lib.check(
        lib.air_elbo_scalars_raw(B, _lib.ptr(baseline), C.byref(pr), C.byref(outs), st), "air_elbo_scalars_raw")
    eng_like(None)
    s = bufs["scalars"].cpu().numpy()
    idx = _lib.SCALAR_INDEX
    assert np.isclose(s[idx["kl_what"]], g["kl_what"], rtol=3e-6) and np.isclose(s[idx["kl_where"]], g["kl_where"], rtol=3e-6)
    assert np.isclose(s[idx["kl_num_steps"]], g["kl_num_steps"], rtol=3e-6)
    assert np.isclose(s[idx["prior_loss"]], g["prior_value"], rtol=3e-6)
    assert np.isclose(s[idx["reinforce_loss"]], g["reinforce_nobaseline"], rtol=2e-5, atol=1e-4)
    eng_like(D32(g["baseline"]).reshape(-1))
    s = bufs["scalars"].cpu().numpy()
    assert np.isclose(s[idx["reinforce_loss"]], g["reinforce_baseline"], rtol=2e-5, atol=2e-4)   # [B,B] broadcast mean


# the tensor-core engine on the script configuration (what it is built for); the fp32 engine on every case
CELL_RUNS = [("script", air.AIR_PREC_FP32), ("script", air.AIR_PREC_TC_SPLIT), ("odd", air.AIR_PREC_FP32),
             ("soft", air.AIR_PREC_FP32)]


@pytest.mark.parametrize("case,precision", CELL_RUNS)
def test_unrolled_forward_vs_reference_cell_vectors(case, precision):
    """air_forward against the vectors the reference's own AIRCell / AIRModel / AIRonMNIST source produced
    (tools/make_golden.py: cell_vectors): 1e-4 absolute + 1e-4 relative on every tensor model.py:86-104 exposes, exact
    presence / step counts away from ties, 1e-4 of the batch's mean magnitude on the reconstruction loss."""
    from oracle import air_oracle as O
    from tests import util as U
    ocfg, params, img, noise, ref = U.load_cell_golden(case)
    T, B = ocfg.T, img.shape[0]
    out = U.run_cuda(ocfg, params, img, noise, O.PriorConfig(), 0, precision=precision, device=DEV)
    for k in ("what", "what_loc", "what_scale", "where", "where_loc", "where_scale", "presence_prob"):
        U.assert_close(out[k].reshape(ref[k].shape), ref[k], atol=1e-4, rtol=1e-4, name=k)
    if ocfg.discrete_steps:
        bad, unsafe = U.presence_mismatches(out["presence"], ref["presence_prob"].reshape(T, B), noise[2].reshape(T, B))
        assert bad == 0, f"{bad} presence mismatches away from ties"
        same = (out["presence"].reshape(T, B) == ref["presence"].reshape(T, B)).all(0)
        assert bool(same.any())
        assert torch.equal(out["num_step_per_sample"].reshape(-1)[same], ref["num_step_per_sample"].reshape(-1)[same])
    else:
        U.assert_close(out["presence"].reshape(ref["presence"].shape), ref["presence"], atol=1e-4, name="presence")
        same = torch.ones(B, dtype=torch.bool)
    U.assert_close(out["canvas"].reshape(T, B, -1)[:, same], ref["canvas"].reshape(T, B, -1)[:, same], atol=1e-4,
                   rtol=1e-4, name="canvas")
    U.assert_close(out["glimpse_viz"].reshape(T, B, -1)[:, same], ref["glimpse"].reshape(T, B, -1)[:, same], atol=1e-4,
                   name="glimpse")
    U.assert_close(out["final_h"], ref["final_h"], atol=1e-4, name="final_h")
    U.assert_close(out["final_c"], ref["final_c"], atol=1e-4, name="final_c")
    U.assert_close(out["num_steps_posterior"], ref["num_steps_posterior"], atol=1e-5, rtol=1e-4, name="q(n)")
    scale = max(1.0, float(ref["rec_loss_per_sample"].abs().mean()))
    U.assert_close(out["rec_loss_per_sample"][same], ref["rec_loss_per_sample"][same], atol=1e-4 * scale, rtol=1e-4,
                   name="rec_loss_per_sample")


@pytest.mark.parametrize("case,precision", [r for r in CELL_RUNS if r[0] != "soft"])     # the backward pass is discrete-only
def test_backward_vs_reference_train_step_vectors(case, precision):
    """air_backward against d opt_loss / d (model variables) as the reference's own AIRModel.train_step computes it
    (tools/make_golden.py: train_vectors; autograd through the reference's loss assembly): 5e-4 of each tensor's max |g| on
    the stored entries, 2.5e-3 on each tensor's norm.  (The 2e-4 bar against the oracle is tests/test_gpu_backward.py; the
    oracle itself sits 3.4e-5 from these vectors.)  The script case carries the BaselineMLP output of AIRonMNIST."""
    from tests import util as U
    from tests.test_gpu_backward import cuda_grads
    from oracle import air_oracle as O
    ocfg, params, img, noise, ref = U.load_cell_golden(case)
    pc, gstep, l2, g = U.load_train_golden(case)
    baseline = torch.from_numpy(np.asarray(g["baseline_out"])) if "baseline_out" in g.files else None
    out, grad = cuda_grads(ocfg, pc, params, img, noise, gstep, baseline=baseline, l2_weight=l2, precision=precision)
    idx = air._lib.SCALAR_INDEX
    for k in ("loss", "rec_loss", "prior_loss", "kl_num_steps", "kl_what", "kl_where", "reinforce_loss"):
        U.assert_close(out["scalars"][idx[k]], torch.as_tensor(np.asarray(g["train:" + k]), dtype=torch.float32),
                       atol=2e-4, rtol=2e-4, name=k)
    off = 0
    for name, shape in O.param_spec(ocfg):
        n = int(np.prod(shape))
        U.compare_with_golden_gradient("grad:", name, grad[off:off + n], g, rel=5e-4)
        off += n
    assert off == grad.numel()
