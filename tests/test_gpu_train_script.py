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
