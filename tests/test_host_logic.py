"""CPU-side tests (no GPU needed): the C-ABI library builds, loads and exports every symbol include/air_b200.h
declares; the Python mirror of the reference's class surface lowers to the configuration / parameter layout the oracle
uses; host-only helpers (Loss, clip_preserve, anneal schedule, prior packing)."""
import ctypes
import os
import re

import pytest
import torch

import attend_infer_repeat_b200 as air
from attend_infer_repeat_b200 import _lib
from oracle import air_oracle as O
from tests import util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    path = _lib.build()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "air_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(air_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    handle = ctypes.CDLL(path)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in air_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.lib().air_abi_version() == 1


def test_struct_layouts_match_header_sizes():
    # air_config: 8 ints + 5 * (1 + AIR_MAX_HIDDEN) ints + 8 floats + 2 ints ; all 4-byte fields -> no padding
    assert ctypes.sizeof(_lib.air_config) == 4 * (8 + 5 * (1 + _lib.AIR_MAX_HIDDEN) + 8 + 2)
    assert ctypes.sizeof(_lib.air_outputs) == 8 * len(_lib.OUTPUT_FIELDS)
    # air_prior: 6 floats, int, (pad), double, int, float, 3 ints, 2 floats with 8-byte alignment
    assert ctypes.sizeof(_lib.air_prior) == 72
    assert _lib.air_prior.steps_success_prob.offset == 32


def test_anneal_weight_matches_oracle():
    for step in (0, 500, 1000, 1500, 20000, 101000, 500000):
        a = air.functional.anneal_weight(1 - 1e-15, 1e-7, "exp", step, 1e5, 1e3, 1e4)
        b = float(O.anneal_weight(1 - 1e-15, 1e-7, "exp", step, 1e5, 1e3, 1e4))
        assert abs(a - b) <= 1e-15 * max(1.0, abs(b)) or abs(a - b) / abs(b) < 1e-12, (step, a, b)
        a = air.functional.anneal_weight(0.9, 0.1, "linear", step, 1e5, 1e3, 1.0)
        b = float(O.anneal_weight(0.9, 0.1, "linear", step, 1e5, 1e3, 1.0))
        assert abs(a - b) < 1e-14
    with pytest.raises(NotImplementedError):
        air.functional.anneal_weight(1, 0, "cosine", 0, 1)


def test_param_layout_matches_oracle():
    for kw in (U.SCRIPT, U.CONFIG_D, U.TINY):
        ocfg = U.oracle_cfg(**kw)
        ccfg = U.cell_cfg(ocfg)
        a = [(n, int(r * c)) for n, (r, c) in air.param_spec(ccfg)]
        b = [(n, int(torch.tensor(s).prod())) for n, s in O.param_spec(ocfg)]
        assert a == b
    assert air.param_count(air.CellConfig()) == 1782525
    assert air.param_count(U.cell_cfg(U.oracle_cfg(**U.CONFIG_D))) == 3899517


def test_module_descriptors_lower_to_script_config():
    """The factories of mnist_model.py:32-41 / multi_mnist.py:82-94 produce the configuration the fused path expects."""
    from functools import partial
    te = partial(air.StochasticTransformParam, [256, 256], scale_bias=.5)(4)
    assert te._n_param == 8 and te._scale_bias == .5 and te._n_hidden == [256, 256]
    assert air.Encoder(5)._n_hidden == [5]                         # test/cell_test.py passes ints
    d = air.Decoder([256, 256], (20, 20))
    assert d.mlp.output_size == 400
    lstm = air.LSTM(256)
    assert lstm.output_size[0] == 256 and lstm.state_size == (256, 256)
    with pytest.raises(RuntimeError):
        air.MLP([4])(torch.zeros(1, 4))                           # unbound module: no silent fallback


def test_make_prior_packing():
    p = air.make_prior(dict(loc=0.1, scale=2.0), dict(loc=0.2, scale=3.0), dict(scale=4.0), 0.25, True, 0.5, False,
                       True, False)
    assert (round(p.what_loc, 6), p.what_scale, p.where_shift_has_loc, p.where_shift_scale) == (0.1, 2.0, 0, 4.0)
    assert (p.steps_success_prob, p.steps_prob_is_f64, p.steps_weight, p.analytic, p.use_prior, p.use_reinforce) == \
        (0.25, 1, 0.5, 0, 1, 0)
    p = air.make_prior(where_shift_prior=dict(loc=0.0, scale=1.0))
    assert p.where_shift_has_loc == 1


def test_loss_accumulator_and_clip_preserve():
    l = air.Loss()
    assert float(l.value) == 0.0
    l.add(torch.tensor(2.0), torch.tensor([1.0, 3.0]))
    inner = air.Loss()
    inner.add(torch.tensor(1.0), torch.tensor([0.5, 1.5]), weight=2.0)
    l.add(inner, weight=0.5)
    assert float(l.value) == 3.0 and l.per_sample.tolist() == [1.5, 4.5]
    with pytest.raises(AssertionError):
        l.add(torch.tensor(1.0), torch.tensor([1.0, 2.0, 3.0]))    # ops.py:26 shape assert
    x = torch.tensor([1e-40, 0.5], requires_grad=True)
    y = air.clip_preserve(x, 1e-32, 1.0)
    y.sum().backward()
    assert y[0].item() == pytest.approx(1e-32) and x.grad.tolist() == [1.0, 1.0]


def test_no_cuda_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(air.AirError):
        air.Engine(air.CellConfig(), 4, 3)
    with pytest.raises(air.AirError):
        air.functional.stn_read(torch.rand(1, 4, 4), torch.rand(1, 4), (2, 2))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "attend_infer_repeat_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_data_pickle_round_trip_and_loaders(tmp_path):
    """The reference's dataset format (data.py:35-118): uint8 pickle -> load_data float32 / 255; tensors_from_data draws
    with replacement when shuffling and (as written, data.py:136-139) always returns the first batch otherwise."""
    import numpy as np
    from attend_infer_repeat_b200 import data as D
    imgs, nums = D.synthetic_multi_mnist_u8(40, 50, 50, seed=3)
    path = os.path.join(str(tmp_path), "mnist_train.pickle")
    D.save_data(path, imgs, nums)
    raw = D.load_raw("mnist_train.pickle", str(tmp_path))
    assert raw["imgs"].dtype == np.uint8 and raw["imgs"].shape == (40, 50, 50) and raw["nums"].shape == (3, 40, 1)
    data = D.load_data("mnist_train.pickle", str(tmp_path))
    assert data["imgs"].dtype == np.float32 and float(data["imgs"].max()) <= 1.0
    np.testing.assert_array_equal(data["imgs"], imgs.astype(np.float32) / 255.)
    axes = {'imgs': 0, 'labels': 0, 'nums': 1}
    t = D.tensors_from_data(data, 8, axes, shuffle=True, seed=0)
    b = t["next_batch"]()
    assert b["imgs"].shape == (8, 50, 50) and b["nums"].shape == (3, 8, 1) and b["labels"].shape[0] == 8
    t = D.tensors_from_data(data, 8, axes, shuffle=False)
    b1, b2 = t["next_batch"](), t["next_batch"]()
    np.testing.assert_array_equal(b1["imgs"], data["imgs"][:8])
    np.testing.assert_array_equal(b2["imgs"], data["imgs"][:8])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle on the host cores; the only arm that runs without a GPU) prints ONE JSON line
    on stdout with the keys the driver reads, and the same metric / unit / config as the CUDA arm."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--batch", "64"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "cell-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "configs[1]" in d["config"]["workload"]


def test_row_kernel_schedule_replays_cleanly_on_the_host():
    """air_row_schedule_check (no GPU): the producer / MMA / epilogue programs of the row kernel (csrc/row_tc.cuh) for the
    script configuration, BASELINE configs[3] and an odd-shaped one are deadlock- and hazard-free under the mbarrier
    protocol; configurations the kernel does not cover are refused with a reason (the engine then uses chain_kernel)."""
    import attend_infer_repeat_b200 as air
    assert air.row_schedule_check(air.CellConfig(precision=air.AIR_PREC_TC_SPLIT)) == ""
    assert air.row_schedule_check(air.CellConfig(H=100, W=100, h=28, w=28, precision=air.AIR_PREC_TC_SPLIT), T=5) == ""
    odd = air.CellConfig(H=9, W=14, h=4, w=6, na=7, nh=256, enc_hidden=(40,), glenc_hidden=(24, 200), dec_hidden=(30,),
                         where_hidden=(20,), steps_hidden=(10,), precision=air.AIR_PREC_TC_SPLIT)
    assert air.row_schedule_check(odd, T=4) == ""
    assert "wider than 256" in air.row_schedule_check(air.CellConfig(glenc_hidden=(300,), precision=air.AIR_PREC_TC_SPLIT))
    assert "what head" in air.row_schedule_check(air.CellConfig(na=80, precision=air.AIR_PREC_TC_SPLIT))


def test_rect_stn_bbox_matches_the_reference_formula():
    """evaluation.py:23-28: x = W (1 - sx + tx) / 2, y = H (1 - sy + ty) / 2, bbox = [y - .5, x - .5, H sy, W sx]."""
    from attend_infer_repeat_b200.evaluation import rect_stn_bbox
    assert rect_stn_bbox(50, 50, (1., 0., 1., 0.)) == [-.5, -.5, 50., 50.]           # identity: the whole image
    b = rect_stn_bbox(50, 40, (.5, .2, .25, -.4))
    assert b == [40 * (1 - .25 - .4) / 2 - .5, 50 * (1 - .5 + .2) / 2 - .5, 40 * .25, 50 * .5]


def test_bind_host_to_device_never_raises_without_a_gpu():
    """sharding.bind_host_to_device: no NVML / no GPU here -> None, and the process keeps its CPU set."""
    import os
    from attend_infer_repeat_b200.sharding import bind_host_to_device
    before = os.sched_getaffinity(0)
    got = bind_host_to_device(0)
    assert got is None or set(got) <= before
    assert os.sched_getaffinity(0) == (before if got is None else set(got))
