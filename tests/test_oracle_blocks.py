"""Independent cross-checks of the parts of the oracle the reference's tests do not pin:
STN read / inverse vs torch grid_sample, LSTM vs torch.nn.LSTMCell, Normal KL / log-prob vs torch.distributions,
anneal schedule vs closed form, REINFORCE [B,B] broadcast identity.  CPU only."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import air_oracle as O


def _rand_where(B, g):
    sx = torch.rand(B, 1, generator=g) * 0.9 + 0.2
    sy = torch.rand(B, 1, generator=g) * 0.9 + 0.2
    tx = torch.rand(B, 1, generator=g) * 1.6 - 0.8
    ty = torch.rand(B, 1, generator=g) * 1.6 - 0.8
    return torch.cat([sx, tx, sy, ty], 1)


def test_stn_read_vs_grid_sample():
    g = torch.Generator().manual_seed(0)
    B = 16
    img = torch.rand(B, 50, 50, generator=g)
    where = _rand_where(B, g)
    crop = O.stn_read(img, where, (20, 20))
    theta = torch.zeros(B, 2, 3)
    theta[:, 0, 0], theta[:, 0, 2], theta[:, 1, 1], theta[:, 1, 2] = where[:, 0], where[:, 1], where[:, 2], where[:, 3]
    grid = F.affine_grid(theta, (B, 1, 20, 20), align_corners=True)
    ref = F.grid_sample(img[:, None], grid, mode="bilinear", padding_mode="zeros", align_corners=True)[:, 0]
    assert (crop - ref).abs().max() < 2e-5


def test_stn_paint_vs_grid_sample():
    g = torch.Generator().manual_seed(1)
    B = 16
    gl = torch.randn(B, 20, 20, generator=g)
    where = _rand_where(B, g)
    inv = O.stn_paint(gl, where, (50, 50))
    theta = torch.zeros(B, 2, 3)
    theta[:, 0, 0] = 1 / where[:, 0]
    theta[:, 0, 2] = -where[:, 1] / where[:, 0]
    theta[:, 1, 1] = 1 / where[:, 2]
    theta[:, 1, 2] = -where[:, 3] / where[:, 2]
    grid = F.affine_grid(theta, (B, 1, 50, 50), align_corners=True)
    ref = F.grid_sample(gl[:, None], grid, mode="bilinear", padding_mode="zeros", align_corners=True)[:, 0]
    assert (inv - ref).abs().max() < 1e-4


def test_stn_roundtrip_identity():
    # where = (1,0,1,0) on equal sizes is the identity for both directions
    img = torch.rand(3, 20, 20)
    where = torch.tensor([[1., 0., 1., 0.]]).repeat(3, 1)
    assert torch.allclose(O.stn_read(img, where, (20, 20)), img, atol=1e-5)
    assert torch.allclose(O.stn_paint(img, where, (20, 20)), img, atol=1e-5)


def test_resample_zero_outside():
    data = torch.ones(1, 4, 4)
    x = torch.tensor([[-1.0, -0.5, 3.5, 4.0, 1.0]])
    y = torch.tensor([[1.0, 1.0, 1.0, 1.0, -1.0]])
    out = O.resample(data, x, y)
    assert torch.allclose(out, torch.tensor([[0.0, 0.5, 0.5, 0.0, 0.0]]))


def test_lstm_vs_torch_lstmcell():
    g = torch.Generator().manual_seed(2)
    nin, nh, B = 7, 5, 4
    w = torch.randn(nin + nh, 4 * nh, generator=g) * 0.3
    b = torch.randn(4 * nh, generator=g) * 0.1
    x, h, c = (torch.randn(B, n, generator=g) for n in (nin, nh, nh))
    h1, c1 = O.lstm_step(x, h, c, w, b, forget_bias=1.0)
    cell = torch.nn.LSTMCell(nin, nh)
    # sonnet gate order (i, j, f, o) -> torch (i, f, g, o); forget bias folded into b_f
    perm = torch.cat([torch.arange(0, nh), torch.arange(2 * nh, 3 * nh), torch.arange(nh, 2 * nh),
                      torch.arange(3 * nh, 4 * nh)])
    bb = b.clone(); bb[2 * nh:3 * nh] += 1.0
    with torch.no_grad():
        cell.weight_ih.copy_(w[:nin, perm].t()); cell.weight_hh.copy_(w[nin:, perm].t())
        cell.bias_ih.copy_(bb[perm]); cell.bias_hh.zero_()
        h2, c2 = cell(x, (h, c))
    assert torch.allclose(h1, h2, atol=1e-6) and torch.allclose(c1, c2, atol=1e-6)


def test_normal_kl_and_logprob_vs_torch_distributions():
    g = torch.Generator().manual_seed(3)
    mu = torch.randn(8, 5, generator=g); s = torch.rand(8, 5, generator=g) + 0.1
    ref = torch.distributions.kl_divergence(torch.distributions.Normal(mu, s), torch.distributions.Normal(0.3, 1.7))
    assert torch.allclose(O.normal_kl(mu, s, 0.3, 1.7), ref, atol=1e-6)
    cfg = O.AirConfig(H=6, W=5)
    obs = torch.rand(3, 6, 5, generator=g); mean = torch.randn(3, 6, 5, generator=g)
    ref = -torch.distributions.Normal(mean, cfg.output_std).log_prob(obs).sum((1, 2))
    assert torch.allclose(O.rec_loss(cfg, obs, mean), ref, rtol=1e-6)


def test_softplus_elu_match_torch():
    x = torch.linspace(-30, 30, 601)
    assert torch.allclose(O.softplus(x), F.softplus(x), atol=1e-6, rtol=1e-6)
    assert torch.allclose(O.elu(x), F.elu(x), atol=1e-6)


def test_anneal_weight_closed_form():
    # the python-float constants enter the float64 island through float32 (tf.cast -> convert_to_tensor), see
    # tests/golden/reference_loss.npz: init 1 - 1e-15 -> 1.0, final 1e-7 -> float32(1e-7)
    import numpy as np
    final = float(np.float32(1e-7))
    pc = O.PriorConfig()
    assert float(O.steps_prior_success_prob(pc, 0)) == 1.0
    assert float(O.steps_prior_success_prob(pc, 1000)) == 1.0                  # hold_init
    s = float(O.steps_prior_success_prob(pc, 51000))
    assert abs(s - final ** 0.5) < 1e-12
    assert float(O.steps_prior_success_prob(pc, 10 ** 6)) == final              # floor at final
    # geometric_prior still clips to 1 - 1e-15 in float64 (prior.py:28), so the step-0 prior is finite
    p0 = O.geometric_prior(O.steps_prior_success_prob(pc, 0), 3)
    assert p0.dtype == torch.float64 and torch.isfinite(torch.log(p0)).all() and float(p0[0]) > 0


def test_reinforce_broadcast_identity():
    # [B] - [B,1] -> [B,B]; the mean equals mean_j((rec_j - mean_i b_i) * logq_j)
    g = torch.Generator().manual_seed(4)
    B = 6
    joint = O.bernoulli_to_modified_geometric(torch.rand(B, 3, generator=g))
    n = torch.randint(0, 4, (B,), generator=g).float()
    rec = torch.randn(B, generator=g) * 100
    base = torch.randn(B, 1, generator=g)
    rl, iw, lp, _ = O.reinforce(joint, n, rec, base)
    assert iw.shape == (B, B)
    want = ((rec - base.mean()) * lp).mean()
    assert torch.allclose(rl, want, rtol=1e-5)


def test_param_count_matches_survey():
    assert O.param_count(O.AirConfig()) == 1782525
    assert O.param_count(O.AirConfig(H=100, W=100, h=28, w=28, T=5)) == 3899517


def test_forward_shapes_and_gradients():
    cfg = O.AirConfig(H=12, W=12, h=6, w=6, T=3, na=5, nh=16, enc_hidden=(16,), glenc_hidden=(16,),
                      dec_hidden=(16,), where_hidden=(16,), steps_hidden=(8,))
    B = 5
    params = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, 0).items()}
    img, _ = O.synthetic_multi_mnist(B, 12, 12, seed=0)
    ew, ea, up = O.make_noise(cfg, B, 0)
    r = O.forward(cfg, O.PriorConfig(), params, img, ew, ea, up)
    assert r["outs"]["canvas"].shape == (3, B, 144) and r["num_steps_posterior"].shape == (B, 4)
    assert r["loss_per_sample"].shape == (B,)
    assert torch.allclose(r["loss_per_sample"].mean(), r["loss"], rtol=1e-5)
    r["opt_loss"].backward()
    for k, v in params.items():
        assert v.grad is not None and torch.isfinite(v.grad).all(), k
    pres = r["outs"]["presence"]
    assert ((pres == 0) | (pres == 1)).all() and (pres[1:] <= pres[:-1]).all()      # monotone, cell.py:148
