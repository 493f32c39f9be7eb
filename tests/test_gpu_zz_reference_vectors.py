"""The CUDA library against the vectors produced by executing the reference's own cell.py / modules.py / neural.py /
model.py / mnist_model.py (tools/make_golden.py: cell_vectors, train_vectors; tests/golden/reference_cell_*.npz): the
unrolled forward pass and the gradient of the training step, through the C ABI.

(Named to sort after the other GPU files: the driver runs `pytest -x`, and these tests were written after the round's
GPU budget was spent -- they run for the first time on the round-end box.)"""
import numpy as np
import pytest
import torch

import attend_infer_repeat_b200 as air

pytestmark = pytest.mark.gpu
DEV = "cuda"


# the tensor-core engine on the script configuration (what it is built for); the fp32 engine on every case
CELL_RUNS = [("script", air.AIR_PREC_FP32), ("script", air.AIR_PREC_TC_SPLIT), ("odd", air.AIR_PREC_FP32),
             ("soft", air.AIR_PREC_FP32)]


@pytest.mark.parametrize("case,precision", CELL_RUNS)
def test_unrolled_forward_vs_reference_cell_vectors(case, precision):
    """air_forward against the vectors the reference's own AIRCell / AIRModel / AIRonMNIST source produced
    (tools/make_golden.py: cell_vectors): 1e-4 absolute + 1e-4 relative on every tensor model.py:86-104 exposes, exact
    presence / step counts away from ties, 1e-4 of the batch's mean magnitude on the reconstruction loss."""
    from oracle import air_oracle as O
    from tests import util as U
    ocfg, params, img, noise, ref = U.load_cell_golden(case)
    T, B = ocfg.T, img.shape[0]
    out = U.run_cuda(ocfg, params, img, noise, O.PriorConfig(), 0, precision=precision, device=DEV)
    for k in ("what", "what_loc", "what_scale", "where", "where_loc", "where_scale", "presence_prob"):
        U.assert_close(out[k].reshape(ref[k].shape), ref[k], atol=1e-4, rtol=1e-4, name=k)
    if ocfg.discrete_steps:
        bad, unsafe = U.presence_mismatches(out["presence"], ref["presence_prob"].reshape(T, B), noise[2].reshape(T, B))
        assert bad == 0, f"{bad} presence mismatches away from ties"
        same = (out["presence"].reshape(T, B) == ref["presence"].reshape(T, B)).all(0)
        assert bool(same.any())
        assert torch.equal(out["num_step_per_sample"].reshape(-1)[same], ref["num_step_per_sample"].reshape(-1)[same])
    else:
        U.assert_close(out["presence"].reshape(ref["presence"].shape), ref["presence"], atol=1e-4, name="presence")
        same = torch.ones(B, dtype=torch.bool)
    U.assert_close(out["canvas"].reshape(T, B, -1)[:, same], ref["canvas"].reshape(T, B, -1)[:, same], atol=1e-4,
                   rtol=1e-4, name="canvas")
    U.assert_close(out["glimpse_viz"].reshape(T, B, -1)[:, same], ref["glimpse"].reshape(T, B, -1)[:, same], atol=1e-4,
                   name="glimpse")
    U.assert_close(out["final_h"], ref["final_h"], atol=1e-4, name="final_h")
    U.assert_close(out["final_c"], ref["final_c"], atol=1e-4, name="final_c")
    U.assert_close(out["num_steps_posterior"], ref["num_steps_posterior"], atol=1e-5, rtol=1e-4, name="q(n)")
    scale = max(1.0, float(ref["rec_loss_per_sample"].abs().mean()))
    U.assert_close(out["rec_loss_per_sample"][same], ref["rec_loss_per_sample"][same], atol=1e-4 * scale, rtol=1e-4,
                   name="rec_loss_per_sample")


@pytest.mark.parametrize("case,precision", CELL_RUNS)     # "soft" = discrete_steps False (cell.py:150-151)
def test_backward_vs_reference_train_step_vectors(case, precision):
    """air_backward against d opt_loss / d (model variables) as the reference's own AIRModel.train_step computes it
    (tools/make_golden.py: train_vectors; autograd through the reference's loss assembly): 5e-4 of each tensor's max |g| on
    the stored entries, 2.5e-3 on each tensor's norm.  (The 2e-4 bar against the oracle is tests/test_gpu_backward.py; the
    oracle itself sits 3.4e-5 from these vectors.)  The script case carries the BaselineMLP output of AIRonMNIST."""
    from tests import util as U
    from tests.test_gpu_backward import cuda_grads
    from oracle import air_oracle as O
    ocfg, params, img, noise, ref = U.load_cell_golden(case)
    pc, gstep, l2, g = U.load_train_golden(case)
    baseline = torch.from_numpy(np.asarray(g["baseline_out"])) if "baseline_out" in g.files else None
    out, grad = cuda_grads(ocfg, pc, params, img, noise, gstep, baseline=baseline, l2_weight=l2, precision=precision)
    idx = air._lib.SCALAR_INDEX
    for k in ("loss", "rec_loss", "prior_loss", "kl_num_steps", "kl_what", "kl_where", "reinforce_loss"):
        U.assert_close(out["scalars"][idx[k]], torch.as_tensor(np.asarray(g["train:" + k]), dtype=torch.float32),
                       atol=2e-4, rtol=2e-4, name=k)
    off = 0
    for name, shape in O.param_spec(ocfg):
        n = int(np.prod(shape))
        U.compare_with_golden_gradient("grad:", name, grad[off:off + n], g, rel=5e-4)
        off += n
    assert off == grad.numel()
