"""A sampled where-scale of EXACTLY zero.

where = loc + scale * eps (cell.py:130-133) cancels to 0.0 in fp32 about once per 5e7 draws -- every few hundred training
steps at B = 4096.  The inverse transformer (modules.py:100-102) divides by it: the glimpse then covers no canvas pixel, the
forward pass stays finite, but d rec / d where is 0 / 0 if the formula is taken literally, and every parameter upstream of
`where` turns into NaN (found with tools/train_nan_probe.py).  The backward pass returns 0 for that draw's gradient through
the painted canvas.  This test constructs such a draw through the public API and checks that the gradient is finite and
equal to the gradient of the neighbouring draw (scale ~ 1e-8, where the literal formula is finite)."""
import numpy as np
import pytest
import torch

import attend_infer_repeat_b200 as air
from oracle import air_oracle as O
from tests import util as U


def _cancelling_eps(loc, sc):
    """An fp32 eps with fl(fl(eps * sc) + loc) == 0 -- the kernel's arithmetic: one rounded product, one rounded sum
    (where_read_kernel) -- or None if none of the 65 candidates around -loc / sc cancels exactly."""
    loc, sc = np.float32(loc), np.float32(sc)
    if not np.isfinite(loc) or not np.isfinite(sc) or sc == 0 or loc == 0:
        return None
    e = np.float32(-np.float64(loc) / np.float64(sc))
    cands, lo, hi = [e], e, e
    for _ in range(32):
        lo = np.nextafter(lo, np.float32(-np.inf))
        hi = np.nextafter(hi, np.float32(np.inf))
        cands += [lo, hi]
    for c in cands:
        if np.float32(np.float32(c * sc) + loc) == np.float32(0.0):
            return float(c)
    return None


def test_cancelling_eps_search_on_the_host():
    """The host-side search itself (runs without a GPU)."""
    rng = np.random.default_rng(0)
    hits = 0
    for _ in range(200):
        loc, sc = np.float32(rng.uniform(0.05, 0.95)), np.float32(rng.uniform(0.3, 1.5))
        e = _cancelling_eps(loc, sc)
        if e is not None:
            hits += 1
            assert np.float32(np.float32(np.float32(e) * sc) + loc) == 0.0
    assert hits > 20


@pytest.mark.gpu
def test_zero_scale_draw_gives_finite_gradients():
    ocfg = U.oracle_cfg(**U.SCRIPT)
    pc = O.PriorConfig()
    B, T = 64, ocfg.T
    params, img, nums, noise = U.make_problem(ocfg, B, seed=3)
    dev = "cuda"
    eng = air.Engine(U.cell_cfg(ocfg), B, T, device=dev)
    eng.train_enable(True)
    flat = O.flatten_params(ocfg, params).to(dev)
    ew, ea, u = (n.to(dev).contiguous() for n in noise)
    img_d = img.to(dev).contiguous()
    pr = U.prior_struct(pc, 20000)

    out = eng.forward(flat, img_d, ew, ea, u, pr)
    torch.cuda.synchronize()
    # loc / scale of the where posterior do not depend on the where draws (nothing feeds back into the LSTM, cell.py:126-133)
    loc = out["where_loc"].reshape(T, B, 4).cpu().numpy().copy()
    sc = out["where_scale"].reshape(T, B, 4).cpu().numpy().copy()
    pres = out["presence"].reshape(T, B).cpu().numpy().copy()
    found = None
    for want_present in (True, False):
        for t in range(T):
            for b in range(B):
                if want_present and pres[t, b] != 1.0:
                    continue
                for k in (0, 2):                 # s_x, s_y
                    e = _cancelling_eps(loc[t, b, k], sc[t, b, k])
                    if e is not None and found is None:
                        found = (t, b, k, e)
        if found is not None:
            break
    if found is None:
        eng.close()
        pytest.skip("no exactly cancelling where draw exists for this problem")
    t, b, k, e = found

    ew0 = ew.clone()
    ew0[t, b, k] = e
    out = eng.forward(flat, img_d, ew0, ea, u, pr)
    torch.cuda.synchronize()
    w = float(out["where"].reshape(T, B, 4)[t, b, k])
    if w != 0.0:
        eng.close()
        pytest.skip(f"the draw did not cancel exactly on the device (where = {w})")
    for name in ("canvas", "glimpse", "loss_per_sample", "scalars"):
        assert bool(torch.isfinite(out[name]).all()), f"forward output {name} is not finite at a zero scale"
    g0 = eng.backward(flat, img_d, ew0, ea, pr).clone()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(g0).all()), f"{int((~torch.isfinite(g0)).sum())} non-finite gradient entries at a zero scale"

    # a neighbouring draw: scale ~ 1e-7, still no canvas pixel inside the glimpse, literal formula finite
    e1, nz = np.float32(e), np.float32(0.0)
    for _ in range(8):
        e1 = np.nextafter(e1, np.float32(np.inf))
        nz = np.float32(np.float32(e1 * np.float32(sc[t, b, k])) + np.float32(loc[t, b, k]))
        if nz != 0.0:
            break
    ew1 = ew0.clone()
    ew1[t, b, k] = float(e1)
    out = eng.forward(flat, img_d, ew1, ea, u, pr)
    g1 = eng.backward(flat, img_d, ew1, ea, pr).clone()
    torch.cuda.synchronize()
    w1 = float(out["where"].reshape(T, B, 4)[t, b, k])
    if w1 != 0.0 and abs(w1) < 1e-5 and bool(torch.isfinite(g1).all()):
        assert float((g0 - g1).abs().max()) <= 1e-3 * float(g1.abs().max()) + 1e-7
    eng.close()
