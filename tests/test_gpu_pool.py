"""EnginePool: several independent batches in flight on separate handles / CUDA streams.  The pool is plumbing -- every
batch must come out bit-identical to the same batch run alone on one engine, and in order."""
import pytest
import torch

import attend_infer_repeat_b200 as air
from oracle import air_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("n_streams", [2, 3])
def test_pool_forward_is_bit_identical_to_one_engine(n_streams):
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B, T, n = 256, 3, 7
    cfg = U.cell_cfg(ocfg, air.AIR_PREC_TC_SPLIT)
    params = O.flatten_params(ocfg, O.init_params(ocfg, 0)).to(DEV)
    pr = U.prior_struct(O.PriorConfig(), 20000)
    g = torch.Generator(device=DEV).manual_seed(5)
    data = [(torch.rand(B, 50, 50, device=DEV, generator=g), torch.randn(T, B, 4, device=DEV, generator=g),
             torch.randn(T, B, cfg.na, device=DEV, generator=g), torch.rand(T, B, 1, device=DEV, generator=g)) for _ in range(n)]
    eng = air.Engine(cfg, B, T, device=DEV)
    keys = ("canvas", "what", "where", "presence", "loss_per_sample", "scalars")
    ref = []
    for d in data:
        out = eng.forward(params, *d, pr)
        ref.append({k: out[k].clone() for k in keys})
    eng.close()
    pool = air.EnginePool(cfg, B, T, n_streams=n_streams, device=DEV)
    assert len(pool) == n_streams
    got = []
    for i, d in enumerate(data):
        e, out = pool.forward(params, *d, pr)
        assert e is pool.engines[i % n_streams]
        with torch.cuda.stream(pool.streams[i % n_streams]):
            got.append({k: out[k].clone() for k in keys})          # before the engine's next batch overwrites its outputs
    pool.join()
    torch.cuda.synchronize()
    assert not torch.equal(ref[0]["loss_per_sample"], ref[1]["loss_per_sample"])
    for r, o in zip(ref, got):
        for k in keys:
            assert torch.equal(r[k], o[k]), k
    pool.close()


@pytest.mark.parametrize("n_batches", [0, 1, 2, 3, 4, 11])
def test_pool_host_stream_equals_synchronous_calls_in_order(n_batches):
    """EnginePool.stream_host_u8 (every handle double-buffers its own feed; batch i on handle i % n) against the synchronous
    air_forward_host_u8_rng call on the same batch and seed: bit-identical, in order, for fewer batches than handles, a
    partial last round and several rounds."""
    from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B, T = 192, 3
    cfg = U.cell_cfg(ocfg, air.AIR_PREC_TC_SPLIT)
    params = O.flatten_params(ocfg, O.init_params(ocfg, 0)).to(DEV)
    pr = U.prior_struct(O.PriorConfig(), 20000)
    batches = [torch.from_numpy(synthetic_multi_mnist_u8(B, 50, 50, seed=30 + i)[0]).pin_memory() for i in range(n_batches)]
    eng = air.Engine(cfg, B, T, device=DEV)
    sc, lps = torch.empty(16).pin_memory(), torch.empty(B).pin_memory()
    ref = []
    for i, b in enumerate(batches):
        eng.forward_host_u8_rng(params, b, 100 + i, pr, sc, lps)
        ref.append((sc.clone(), lps.clone()))
    eng.close()
    pool = air.EnginePool(cfg, B, T, n_streams=3, device=DEV)
    got = [(s.clone(), l.clone()) for s, l in pool.stream_host_u8(params, iter(batches), pr, seed0=100)]
    assert len(got) == n_batches
    for (s0, l0), (s1, l1) in zip(ref, got):
        assert torch.equal(s0, s1) and torch.equal(l0, l1)
    pool.close()


def test_glimpse_viz_on_request_equals_the_fused_output_and_the_oracle():
    """model.py:90 `presence * sigmoid(glimpse)`: air_glimpse_viz on request is bit-identical to what the paint kernel writes
    when outs->glimpse_viz is given, and AIRModel.glimpse (lazy) matches the oracle."""
    import attend_infer_repeat_b200.functional as F
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B, T = 96, 3
    cfg = U.cell_cfg(ocfg, air.AIR_PREC_TC_SPLIT)
    params = O.flatten_params(ocfg, O.init_params(ocfg, 0)).to(DEV)
    pr = U.prior_struct(O.PriorConfig(), 20000)
    g = torch.Generator(device=DEV).manual_seed(11)
    d = (torch.rand(B, 50, 50, device=DEV, generator=g), torch.randn(T, B, 4, device=DEV, generator=g),
         torch.randn(T, B, cfg.na, device=DEV, generator=g), torch.rand(T, B, 1, device=DEV, generator=g))
    eng = air.Engine(cfg, B, T, device=DEV)                               # materialise_viz=True
    out = eng.forward(params, *d, pr)
    viz = F.glimpse_viz(out["glimpse"], out["presence"])
    assert torch.equal(viz, out["glimpse_viz"])
    assert float(viz.abs().max()) > 0
    lean = air.Engine(cfg, B, T, device=DEV, materialise_viz=False)
    out2 = lean.forward(params, *d, pr)
    assert out2["glimpse_viz"] is None
    for k in ("canvas", "glimpse", "loss_per_sample", "scalars"):
        assert torch.equal(out[k], out2[k]), k
    eng.close()
    lean.close()


def test_launch_overlap_switch_changes_scheduling_only():
    """air_set_launch_overlap (programmatic dependent launch on / off): bit-identical outputs either way; the pool turns it off
    for its engines when more than one batch is in flight and leaves it on for a pool of one."""
    ocfg = U.oracle_cfg(**U.SCRIPT)
    B, T = 128, 3
    cfg = U.cell_cfg(ocfg, air.AIR_PREC_TC_SPLIT)
    params = O.flatten_params(ocfg, O.init_params(ocfg, 0)).to(DEV)
    pr = U.prior_struct(O.PriorConfig(), 20000)
    g = torch.Generator(device=DEV).manual_seed(3)
    d = (torch.rand(B, 50, 50, device=DEV, generator=g), torch.randn(T, B, 4, device=DEV, generator=g),
         torch.randn(T, B, cfg.na, device=DEV, generator=g), torch.rand(T, B, 1, device=DEV, generator=g))
    eng = air.Engine(cfg, B, T, device=DEV)
    keys = ("canvas", "glimpse", "what", "where", "presence", "loss_per_sample", "scalars")
    a = {k: v.clone() for k, v in eng.forward(params, *d, pr).items() if k in keys}
    eng.set_launch_overlap(False)
    b = {k: v.clone() for k, v in eng.forward(params, *d, pr).items() if k in keys}
    eng.set_launch_overlap(True)
    c = {k: v.clone() for k, v in eng.forward(params, *d, pr).items() if k in keys}
    torch.cuda.synchronize()
    for k in keys:
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), k
    eng.close()
    with pytest.raises(Exception):
        air._lib.check(air._lib.lib().air_set_launch_overlap(None, 1), "air_set_launch_overlap")
