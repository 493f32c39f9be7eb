"""scripts/train_multi_mnist.py (the reference's training entry point, scripts/multi_mnist.py:82-147) end to end on the
device: it trains, logs the evaluation.py:68-92 scalars, checkpoints, and a run resumed from a checkpoint continues
bit-identically (parameters, both optimisers' slots, global_step, and the noise / minibatch / validation streams are all
part of the checkpoint)."""
import importlib.util
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _script():
    spec = importlib.util.spec_from_file_location("train_multi_mnist", os.path.join(ROOT, "scripts", "train_multi_mnist.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_script_defaults_are_the_reference_hyper_parameters():
    """multi_mnist.py:22-26: learning_rate 1e-4 (the baseline optimiser then runs at 1e-3), batch 64."""
    import argparse
    S = _script()
    src = open(os.path.join(ROOT, "scripts", "train_multi_mnist.py")).read()
    assert '"--learning-rate", type=float, default=1e-4' in src
    assert '"--batch-size", type=int, default=64' in src
    assert callable(S.main) and isinstance(argparse.ArgumentParser(), argparse.ArgumentParser)


def test_train_script_checkpoint_resume_is_bit_identical(tmp_path, monkeypatch):
    # the tensor-core weight-gradient GEMMs split K over CTAs and combine with fp32 atomics (order varies run to run, 1 ulp);
    # ten REINFORCE steps amplify that, so the bit-identity check runs on the deterministic fp32 SIMT gradient GEMMs
    monkeypatch.setenv("AIR_NO_TC_BWD", "1")
    S = _script()
    common = ["--batch-size", "32", "--n-synthetic", "512", "--log-every", "10", "--save-every", "10", "--precision", "tc",
              "--iters", "20"]
    log = tmp_path / "log.jsonl"
    m1 = S.main(common + ["--checkpoint-dir", str(tmp_path / "a"), "--log-json", str(log)])
    p20 = m1.params.clone()
    b20 = m1.baseline_module.params.clone()
    assert m1.global_step == 20
    assert os.path.exists(tmp_path / "a" / "model-10.pt") and os.path.exists(tmp_path / "a" / "model-20.pt")
    # the logged scalars are the reference's (evaluation.py:68-92) and finite
    lines = [json.loads(l) for l in open(log)]
    assert [l["step"] for l in lines] == [0, 10, 20]
    for l in lines:
        for split in ("train", "test"):
            assert set(l[split]) >= {"loss", "rec_loss", "num_step_acc", "num_step", "prior_loss", "kl_num_steps", "kl_what",
                                     "kl_where", "baseline_loss", "reinforce_loss"}
            assert all(v == v and abs(v) < 1e9 for v in l[split].values()), l
    del m1
    m2 = S.main(common + ["--checkpoint-dir", str(tmp_path / "b"), "--resume", str(tmp_path / "a" / "model-10.pt")])
    assert m2.global_step == 20
    assert torch.equal(m2.params, p20), float((m2.params - p20).abs().max())
    assert torch.equal(m2.baseline_module.params, b20)
    ck_a, ck_b = torch.load(tmp_path / "a" / "model-20.pt"), torch.load(tmp_path / "b" / "model-20.pt")
    for k in ("mg", "ms", "mom"):
        assert torch.equal(ck_a["slots"][k], ck_b["slots"][k]), k


def test_progress_figure_data_and_logger(tmp_path):
    """evaluation.py:31-108 on the device: make_fig's arrays / rectangles for the current batch and make_logger's scalar set."""
    import numpy as np
    import attend_infer_repeat_b200 as air
    from attend_infer_repeat_b200.data import ResidentDataset, synthetic_multi_mnist_u8
    from attend_infer_repeat_b200.evaluation import make_fig, make_logger, rect_stn_bbox
    dev = torch.device("cuda", 0)
    ds = ResidentDataset(*synthetic_multi_mnist_u8(256, 50, 50, seed=0), device=dev, seed=0)
    B = 16
    imgs, nums = ds.gather(ds.next_indices(B))
    model = air.AIRonMNIST(imgs, nums, max_steps=3, explore_eps=1e-3, inpt_encoder_hidden=[256, 256],
                           glimpse_encoder_hidden=[256, 256], glimpse_decoder_hidden=[256, 256],
                           transform_estimator_hidden=[256, 256], steps_pred_hidden=[128, 64], baseline_hidden=[256, 128],
                           transform_var_bias=.5, step_bias=.75, output_multiplier=.5, precision=air.AIR_PREC_TC_SPLIT)
    pr = dict(loc=0., scale=1.)
    nsp = dict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3, analytic=True)
    train_op, _ = model.train_step(1e-4, 0., pr, pr, pr, nsp)
    train_op()
    d = make_fig(model, str(tmp_path), 1, n_samples=10)
    assert os.path.exists(tmp_path / "progress_fig_1.npz")
    assert d["obs"].shape == (10, 50, 50) and d["canvas"].shape == (3, 10, 50, 50) and d["glimpse"].shape == (3, 10, 20, 20)
    assert d["prob"].shape == (10, 3) and d["bbox"].shape == (3, 10, 4)
    pres, where = model.presence[:, :10, 0].cpu().numpy(), model.where[:, :10].cpu().numpy()
    for i in range(3):
        for j in range(10):
            if pres[i, j] > .5:
                np.testing.assert_allclose(d["bbox"][i, j], rect_stn_bbox(50, 50, where[i, j]), rtol=1e-6)
            else:
                assert np.isnan(d["bbox"][i, j]).all()
    lines = []
    log = make_logger(model, lambda: ds.gather(ds.next_indices(B)), 2, lambda: ds.gather(ds.next_indices(B)), 2, out=lines.append)
    res = log(7)
    assert len(lines) == 2 and lines[0].startswith("Step 7, Data train loss = ") and "Data test" in lines[1]
    assert set(res["train"]) == {"loss", "rec_loss", "num_step_acc", "num_step", "prior_loss", "kl_num_steps", "kl_what",
                                 "kl_where", "baseline_loss", "reinforce_loss", "imp_weight"}
    assert all(v == v for v in res["test"].values())
