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
