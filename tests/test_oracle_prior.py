"""Pins the oracle's step-count algebra on every known answer the reference's own tests hold
(/root/reference/test/prior_test.py; cited per test).  CPU only."""
import numpy as np
import torch
from numpy.testing import assert_array_almost_equal, assert_array_equal

from oracle import air_oracle as O

N_STRESS = 100


def test_geometric_prior_known_answer():
    # prior_test.py:15-24
    prob, n_steps = .75, 10
    expected = (1. - prob) * prob ** np.arange(n_steps + 1)
    p = O.geometric_prior(prob, n_steps).numpy()
    assert p.dtype == np.float32
    assert_array_almost_equal(p, expected)


def test_tabular_kl_same():
    # prior_test.py:40-44
    p = torch.full((1, 4), .25)
    kl = O.tabular_kl(p, p, 0.).numpy()
    assert kl.shape == (1, 4)
    assert kl.sum() == 0.


def test_tabular_kl_zero_and_one():
    # prior_test.py:46-60
    p = torch.tensor([[0., .25, .25, .5]])
    q = torch.tensor([[.25] * 4])
    assert O.tabular_kl(p, q).sum() > 0
    p = torch.tensor([[0., 1., 0., 0.]])
    q = torch.tensor([[1. - 1e-7, 1e-7, 0., 0.]])
    kl = O.tabular_kl(p, q)
    assert kl.sum() > 0 and torch.isfinite(kl).all()


def test_tabular_kl_positive_on_random():
    # prior_test.py:62-74
    rng = np.random.RandomState(0)
    for _ in range(N_STRESS):
        p = abs(rng.rand(1, 4)); p /= p.sum()
        q = abs(rng.rand(1, 4)); q /= q.sum()
        assert O.tabular_kl(torch.tensor(p, dtype=torch.float32), torch.tensor(q, dtype=torch.float32)).sum() > 0


def test_modified_geometric_shapes():
    # prior_test.py:86-98
    for shp in [(3,), (7, 3), (7, 11, 3)]:
        out = O.bernoulli_to_modified_geometric(torch.rand(*shp))
        assert tuple(out.shape) == shp[:-1] + (4,)
        assert out.dtype == torch.float32


def test_modified_geometric_obvious():
    # prior_test.py:100-115 (assert_array_equal: exact)
    cases = {(0., 0., 0.): [1., 0., 0., 0.], (1., 0., 0.): [0., 1., 0., 0.],
             (1., 1., 0.): [0., 0., 1., 0.], (1., 1., 1.): [0., 0., 0., 1.]}
    for p, want in cases.items():
        assert_array_equal(O.bernoulli_to_modified_geometric(torch.tensor(p)).numpy(), want)


def test_modified_geometric_geom():
    # prior_test.py:117-120
    p = O.bernoulli_to_modified_geometric(torch.tensor([.5, .5, .5])).numpy()
    assert_array_equal(p, [.5, .5 ** 2, .5 ** 3, .5 ** 3])


def _kl_and_grad(x, free):
    x = x.clone().requires_grad_(True)
    prior = O.geometric_prior(.005, 3)
    post = x if free else O.bernoulli_to_modified_geometric(x)
    kl = O.tabular_kl(post, prior, 0.)
    g, = torch.autograd.grad(kl.sum(), x)
    return kl.detach(), g


def test_num_steps_kl_free_stress():
    # prior_test.py:160-172
    rng = np.random.RandomState(1)
    for _ in range(N_STRESS):
        p = abs(rng.rand(1, 4)); p /= p.sum()
        kl, g = _kl_and_grad(torch.tensor(p, dtype=torch.float32), True)
        assert kl.sum() > 0 and torch.isfinite(kl).all() and torch.isfinite(g).all()


def test_num_steps_kl_posterior_stress():
    # prior_test.py:174-186
    rng = np.random.RandomState(2)
    for _ in range(N_STRESS):
        kl, g = _kl_and_grad(torch.tensor(rng.rand(1, 3), dtype=torch.float32), False)
        assert kl.sum() > 0 and torch.isfinite(kl).all() and torch.isfinite(g).all()


def test_num_steps_kl_posterior_zeros():
    # prior_test.py:188-205
    kl, g = _kl_and_grad(torch.tensor([[.5, 0., 0.]]), False)
    assert kl.sum() > 0 and torch.isfinite(kl).all() and torch.isfinite(g).all()


def test_log_prob_gather_and_clip():
    # prior.py:103-116,148-151; ops.py:67-76
    joint = torch.tensor([[.5, .25, .125, .125], [0., 1., 0., 0.]], requires_grad=True)
    n = torch.tensor([2., 0.])
    lp = O.num_steps_log_prob(joint, n)
    assert_array_almost_equal(lp.detach().numpy(), [np.log(.125), np.log(np.float32(1e-32))])
    g, = torch.autograd.grad(lp.sum(), joint)
    assert g[0, 2] == 8.0 and torch.isfinite(g).all()       # clip is identity in the backward pass
