"""The reference's own prior tests (test/prior_test.py) run against the CUDA implementations of prior.py -- same
inputs, same known answers, same assertions (assert_array_equal where the reference is exact)."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_array_almost_equal, assert_array_equal

from attend_infer_repeat_b200.prior import (NumStepsDistribution, bernoulli_to_modified_geometric, geometric_prior,
                                            tabular_kl)
from oracle import air_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
_N_STRESS_ITER = 100


def _t(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.float32, device=DEV)


def test_geometric_prior_known_answer():
    """test/prior_test.py:13-24"""
    prob, n_steps = .75, 10
    expected = (1. - prob) * prob ** np.arange(n_steps + 1)
    p = geometric_prior(prob, n_steps, device=DEV).cpu().numpy()
    assert p.dtype == np.float32 and p.shape == (n_steps + 1,)
    assert_array_almost_equal(p, expected)
    # float64 island with the script's extreme values (multi_mnist.py:40-42)
    for s in (1. - 1e-15, 1e-7, 0.3):
        p64 = geometric_prior(s, 3, device=DEV, float64=True).cpu()
        ref = O.geometric_prior(torch.tensor(s, dtype=torch.float64), 3)
        assert torch.allclose(p64, ref, rtol=1e-12, atol=0), (s, p64, ref)


def test_tabular_kl_same_zero_one():
    """test/prior_test.py:40-60"""
    p = np.asarray([.25] * 4).reshape((1, 4))
    kl = tabular_kl(_t(p), torch.as_tensor(p[0])).cpu().numpy()
    assert kl.shape == (1, 4)
    assert kl.sum() == 0.
    p = np.asarray([0., .25, .25, .5]).reshape((1, 4))
    q = np.asarray([.25] * 4)
    kl = tabular_kl(_t(p), torch.as_tensor(q)).cpu().numpy()
    assert kl.sum() > 0. and np.isfinite(kl).all() and kl[0, 0] == 0.
    p = np.asarray([0., 1., 0., 0.]).reshape((1, 4))
    q = np.asarray([1. - 1e-7, 1e-7, 0., 0.])
    kl = tabular_kl(_t(p), torch.as_tensor(q)).cpu().numpy()
    assert kl.sum() > 0. and np.isfinite(kl).all()


def test_tabular_kl_always_positive_on_random():
    """test/prior_test.py:62-74"""
    rng = np.random.RandomState(0)

    def gen():
        a = abs(rng.rand(1, 4))
        return a / a.sum()

    for _ in range(_N_STRESS_ITER):
        p, q = gen(), gen()
        kl = tabular_kl(_t(p), torch.as_tensor(q[0])).cpu()
        assert kl.sum() > 0.
        ref = O.tabular_kl(torch.as_tensor(p, dtype=torch.float32), torch.as_tensor(q[0]))
        assert torch.allclose(kl, ref, rtol=1e-6, atol=1e-9)


def test_modified_geometric_shapes():
    """test/prior_test.py:86-98"""
    assert tuple(bernoulli_to_modified_geometric(torch.rand(3, device=DEV)).shape) == (4,)
    assert tuple(bernoulli_to_modified_geometric(torch.rand(7, 3, device=DEV)).shape) == (7, 4)
    assert tuple(bernoulli_to_modified_geometric(torch.rand(7, 11, 3, device=DEV)).shape) == (7, 11, 4)


def test_modified_geometric_known_answers_exact():
    """test/prior_test.py:100-120 (assert_array_equal: exact)"""
    f = lambda p: bernoulli_to_modified_geometric(_t(p)).cpu().numpy()
    assert_array_equal(f([0., 0., 0.]), [1., 0., 0., 0.])
    assert_array_equal(f([1., 0., 0.]), [0., 1., 0., 0.])
    assert_array_equal(f([1., 1., 0.]), [0., 0., 1., 0.])
    assert_array_equal(f([1., 1., 1.]), [0., 0., 0., 1.])
    assert_array_equal(f([.5, .5, .5]), [.5, .5 ** 2, .5 ** 3, .5 ** 3])


def test_modified_geometric_matches_oracle_bitwise():
    g = torch.Generator().manual_seed(0)
    for T in (1, 2, 3, 5, 8):
        p = torch.rand(257, T, generator=g)
        p[0] = 0.
        p[1] = 1.
        assert torch.equal(bernoulli_to_modified_geometric(p.to(DEV)).cpu(), O.bernoulli_to_modified_geometric(p))


def test_kl_posterior_prior_stress():
    """test/prior_test.py:141-205: prior geometric_prior(.005, 3); KL finite and positive for random posteriors and
    for [.5, 0, 0]."""
    prior = geometric_prior(.005, 3, device=DEV)
    rng = np.random.RandomState(1)
    for i in range(_N_STRESS_ITER + 1):
        p = np.asarray([.5, 0., 0.]) if i == 0 else rng.rand(3)
        post = bernoulli_to_modified_geometric(_t(p))
        kl = tabular_kl(post.reshape(1, 4), prior).cpu().numpy()
        assert np.isfinite(kl).all() and kl.sum() > 0.


def test_num_steps_distribution_gather_exact():
    """prior.py:119-151: prob(n) gathers joint[b, int(n_b)] exactly; log_prob clips at 1e-32."""
    g = torch.Generator().manual_seed(2)
    B, T = 1000, 3
    probs = torch.rand(B, T, generator=g)
    probs[:10] = 0.
    d = NumStepsDistribution(probs.to(DEV))
    joint = d.prob().cpu()
    assert torch.equal(joint, O.bernoulli_to_modified_geometric(probs))
    n = torch.randint(0, T + 1, (B,), generator=g).float()
    got = d.prob(n.to(DEV)).cpu()
    assert torch.equal(got, torch.gather(joint, 1, n.long()[:, None])[:, 0])
    lp = d.log_prob(n.to(DEV)).cpu()
    ref = O.num_steps_log_prob(joint, n)
    assert torch.allclose(lp, ref, rtol=1e-6, atol=1e-6)
    assert torch.isfinite(lp).all() and float(lp.min()) >= np.log(1e-32) - 1e-3
    s = d.sample().cpu()
    assert ((s >= 0) & (s <= T) & (s == s.round())).all()
