"""Shared helpers for the parity tests: build the same problem for the CPU oracle and for the CUDA library."""
import torch

import attend_infer_repeat_b200 as air
from oracle import air_oracle as O

SCRIPT = dict(H=50, W=50, h=20, w=20, T=3, na=50, nh=256)                      # scripts/multi_mnist.py
CONFIG_D = dict(H=100, W=100, h=28, w=28, T=5, na=50, nh=256)                  # BASELINE.json configs[3]
TINY = dict(H=3, W=3, h=2, w=2, T=3, na=10, nh=8, enc_hidden=(5,), glenc_hidden=(7,), dec_hidden=(11,),
            where_hidden=(13,), steps_hidden=(17,))                             # test/cell_test.py:9-29 widths


def oracle_cfg(**kw) -> O.AirConfig:
    return O.AirConfig(**kw)


def cell_cfg(ocfg: O.AirConfig, precision=air.AIR_PREC_FP32) -> air.CellConfig:
    d = {k: getattr(ocfg, k) for k in ("H", "W", "h", "w", "na", "nh", "enc_hidden", "glenc_hidden", "dec_hidden",
                                       "where_hidden", "steps_hidden", "output_std", "output_multiplier",
                                       "explore_eps", "scale_bias", "step_bias", "what_scale_offset", "forget_bias",
                                       "max_crop_size", "discrete_steps")}
    return air.CellConfig(precision=precision, **d)


def make_problem(ocfg, B, seed=0, weight_gain=1.0, random_bias=True):
    """Seeded weights / images / noise shared verbatim by the oracle and the kernel.  Biases and the
    trainable LSTM initial state get small random values so that every parameter influences the result."""
    params = O.init_params(ocfg, seed)
    g = torch.Generator().manual_seed(seed + 77)
    for k, v in params.items():
        if k.endswith(".w"):
            v.mul_(weight_gain)
        elif random_bias:
            v.copy_(0.1 * torch.randn(v.shape, generator=g))
    if min(ocfg.H, ocfg.W) >= 8:
        img, nums = O.synthetic_multi_mnist(B, ocfg.H, ocfg.W, seed=seed)
    else:   # the 3x3 images of test/cell_test.py:25,55 are np.random.rand
        img, nums = torch.rand(B, ocfg.H, ocfg.W, generator=g), torch.zeros(3, B, 1)
    noise = O.make_noise(ocfg, B, seed)
    return params, img, nums, noise


def prior_struct(pc: O.PriorConfig, global_step=0):
    s = O.steps_prior_success_prob(pc, global_step)
    is64 = pc.steps_anneal is not None
    return air.make_prior(
        dict(loc=pc.what_loc, scale=pc.what_scale),
        dict(loc=pc.where_scale_loc, scale=pc.where_scale_scale),
        dict(scale=pc.where_shift_scale) if pc.where_shift_loc is None else
        dict(loc=pc.where_shift_loc, scale=pc.where_shift_scale),
        float(s), is64, pc.steps_weight, pc.analytic, pc.use_prior, pc.use_reinforce)


def run_cuda(ocfg, params, img, noise, pc=None, global_step=0, baseline=None, precision=air.AIR_PREC_FP32,
             device="cuda"):
    B = img.shape[0]
    eng = air.Engine(cell_cfg(ocfg, precision), B, ocfg.T, device=device)
    flat = O.flatten_params(ocfg, params).to(device)
    ew, ea, u = (n.to(device).contiguous() for n in noise)
    pr = prior_struct(pc, global_step) if pc is not None else None
    bl = None if baseline is None else baseline.reshape(-1).to(device).contiguous()
    out = eng.forward(flat, img.to(device).contiguous(), ew, ea, u, pr, bl)
    torch.cuda.synchronize()
    eng.check_range()
    res = {k: (None if v is None else v.detach().cpu().clone()) for k, v in out.items()}
    eng.close()
    return res


def assert_close(a, b, atol=1e-4, rtol=1e-4, name=""):
    a, b = a.reshape(-1).double(), b.reshape(-1).double()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), f"{name}: max err {err.max():.3e} (tol {tol[err.argmax()]:.3e}), {int(bad.sum())} bad of {a.numel()}"


def presence_mismatches(pres_cuda, p_oracle, u, margin=1e-5):
    """Bit-exact rule for presence (SURVEY 7 'hard parts'): wherever the uniform draw is farther than `margin` from
    the oracle's probability at every step up to t, the cumulative presence must match exactly."""
    T = p_oracle.shape[0]
    z = (u < p_oracle).float()
    safe_step = (u - p_oracle).abs() > margin
    pres_o = torch.cumprod(z, 0)
    safe = torch.cumprod(safe_step.float(), 0).bool()
    mism = (pres_cuda.reshape(T, -1) != pres_o.reshape(T, -1)) & safe.reshape(T, -1)
    return int(mism.sum()), int((~safe).sum())


# ---- golden vectors of the reference's own AIRCell / AIRModel source (tools/make_golden.py: cell_vectors) ------------------
CELL_GOLDEN_CASES = ("script", "odd", "soft")


def load_cell_golden(case):
    """-> (oracle config, params dict, img, noise tuple, dict of the reference's outputs as torch tensors)."""
    import json
    import os
    import numpy as np
    from tests.golden_recipe import golden_tensor
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cell_%s.npz" % case))
    c = json.loads(str(g["cfg_json"]))
    ocfg = O.AirConfig(H=c["H"], W=c["W"], h=c["h"], w=c["w"], T=c["T"], na=c["na"], nh=c["nh"],
                       enc_hidden=tuple(c["enc"]), glenc_hidden=tuple(c["glenc"]), dec_hidden=tuple(c["dec"]),
                       where_hidden=tuple(c["where"]), steps_hidden=tuple(c["steps"]), output_std=c["output_std"],
                       output_multiplier=c["output_multiplier"], explore_eps=c["explore_eps"],
                       scale_bias=c["transform_var_bias"], step_bias=c["step_bias"], discrete_steps=c["discrete_steps"])
    spec = O.param_spec(ocfg)
    # the variables the reference's graph created are exactly the entries of the flat parameter buffer
    created = {str(n): tuple(int(x) for x in s) for n, s in zip(g["param_names"], g["param_shapes"])}
    two_d = lambda s: (1, int(s[0])) if len(s) == 1 else (int(s[0]), int(s[1]))      # biases / h0 / c0 are rows
    assert created == {n: two_d(s) for n, s in spec}, (created, spec)
    params = {n: torch.from_numpy(golden_tensor(n, two_d(s), c["seed"])).reshape(tuple(s)) for n, s in spec}
    noise = tuple(torch.from_numpy(g[k]) for k in ("eps_where", "eps_what", "u_pres"))
    forward_keys = ("what", "what_loc", "what_scale", "where", "where_loc", "where_scale", "presence_prob", "presence",
                    "canvas", "glimpse", "final_canvas", "num_step_per_sample", "final_h", "final_c",
                    "num_steps_posterior", "rec_loss_per_sample")
    ref = {k: torch.from_numpy(np.asarray(g[k])) for k in forward_keys}
    return ocfg, params, torch.from_numpy(g["img"]), noise, ref


def load_train_golden(case):
    """The AIRModel.train_step part of the vectors: (PriorConfig, global_step, l2_weight, g) with g the raw npz."""
    import json
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cell_%s.npz" % case))
    tc = json.loads(str(g["train_cfg_json"]))
    pc = O.PriorConfig(analytic=tc["analytic"], where_shift_loc=0.0 if tc["shift_has_loc"] else None,
                       where_shift_scale=1.0 if tc["shift_has_loc"] else 0.8)
    return pc, tc["global_step"], tc["l2_weight"], g


def golden_baseline_params(g, seed):
    """BaselineMLP parameters of the script case (re-drawn from the recipe), keyed 'baseline.<i|out>.<w|b>'."""
    from tests.golden_recipe import golden_tensor
    out = {}
    for n, s in zip(g["baseline_param_names"], g["baseline_param_shapes"]):
        t = torch.from_numpy(golden_tensor(str(n), (int(s[0]), int(s[1])), seed))
        out[str(n)] = t.reshape(-1) if str(n).endswith(".b") else t
    return out


def compare_with_golden_gradient(prefix, name, got, g, rel=2e-4):
    """`got`: full gradient tensor; the vectors hold its entries at golden_subset(name), its L2 norm and max magnitude."""
    import numpy as np
    from tests.golden_recipe import golden_subset
    flat = got.detach().reshape(-1).double().cpu().numpy()
    ref = g[prefix + name].astype(np.float64)
    norm, gmax = (float(x) for x in g[prefix + "stats:" + name])
    idx = golden_subset(name, flat.size)
    err = float(np.abs(flat[idx] - ref).max()) if ref.size else 0.0
    tol = rel * gmax + 1e-7
    assert err <= tol, f"{name}: max |g - g_ref| {err:.3e} > {tol:.3e} (max |g_ref| {gmax:.3e})"
    got_norm = float(np.sqrt((flat ** 2).sum()))
    assert abs(got_norm - norm) <= 5 * rel * norm + 1e-7, f"{name}: |g| {got_norm:.6e} vs {norm:.6e}"
