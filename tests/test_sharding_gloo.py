"""world_size-2 `gloo` tests (CPU) of the batch-sharding host logic: contiguous ragged shards, the
16-float scalar exchange that re-forms the whole-batch loss / REINFORCE terms, and the weighted gradient all-reduce.
The per-shard numbers are produced by the oracle (the checker), exactly as the kernel would produce them per rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from attend_infer_repeat_b200 import sharding
from attend_infer_repeat_b200._lib import AIR_N_SCALARS, SCALAR_INDEX
from oracle import air_oracle as O
from tests import util as U


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_scalars(res, baseline, use_reinforce=True):
    """What elbo_scalars_kernel writes for one shard (batch means over the shard)."""
    s = torch.zeros(AIR_N_SCALARS)
    iw, lq = res["rec_loss_per_sample"], res["num_steps_log_prob"]
    s[SCALAR_INDEX["rec_loss"]] = res["rec_loss_per_sample"].mean()
    s[SCALAR_INDEX["kl_num_steps"]] = res["kl_num_steps_per_sample"].mean()
    s[SCALAR_INDEX["kl_what"]] = res["kl_what_per_sample"].mean()
    s[SCALAR_INDEX["kl_where"]] = res["kl_where_per_sample"].mean()
    s[SCALAR_INDEX["num_step"]] = res["num_step_per_sample"].mean()
    s[SCALAR_INDEX["mean_iw_logq"]] = (iw * lq).mean()
    s[SCALAR_INDEX["mean_logq"]] = lq.mean()
    s[SCALAR_INDEX["mean_baseline"]] = baseline.mean() if baseline is not None else 0.0
    return s


def _worker(rank, world, port, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        ocfg = U.oracle_cfg(**U.TINY)
        pc = O.PriorConfig()
        params, img, nums, noise = U.make_problem(ocfg, B, seed=3)
        baseline = torch.linspace(-1.0, 2.0, B).reshape(B, 1)
        # -- the rank's shard (ragged when B is odd) ------------------------------------------------------------
        a, b = sharding.shard_range(B, rank, world)
        img_s = sharding.shard(img, rank, world, 0)
        noise_s = tuple(sharding.shard(n, rank, world, 1) for n in noise)
        assert img_s.shape[0] == b - a and noise_s[0].shape[1] == b - a
        p_local = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        res = O.forward(ocfg, pc, p_local, img_s, *noise_s, global_step=20000)
        scal = _local_scalars({k: v.detach() for k, v in res.items() if torch.is_tensor(v)}, baseline[a:b])
        sharding.combine_scalars(scal, b - a, pc.steps_weight, pc.use_prior, pc.use_reinforce)
        # gradient of the shard's mean loss -> weighted all-reduce = gradient of the global mean loss
        res["loss"].backward()
        flat = O.flatten_params(ocfg, {k: v.grad for k, v in p_local.items()}).clone()
        sharding.allreduce_gradient(flat, b - a, B)
        bmean = sharding.global_baseline_mean(baseline[a:b].reshape(-1), b - a)
        if rank == 0:
            ret["scalars"] = scal
            ret["grad"] = flat
            ret["bmean"] = bmean
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_two_rank_shards_reproduce_the_single_device_batch(B):
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, B, ret), nprocs=world, join=True)
        scal, grad, bmean = ret["scalars"], ret["grad"], ret["bmean"]
    # single-device reference: the oracle on the whole batch
    ocfg = U.oracle_cfg(**U.TINY)
    pc = O.PriorConfig()
    params, img, nums, noise = U.make_problem(ocfg, B, seed=3)
    baseline = torch.linspace(-1.0, 2.0, B).reshape(B, 1)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    res = O.forward(ocfg, pc, p, img, *noise, global_step=20000, baseline=baseline)
    for name in ("rec_loss", "kl_num_steps", "kl_what", "kl_where", "prior_loss", "loss", "reinforce_loss", "opt_loss"):
        a, b = float(scal[SCALAR_INDEX[name]]), float(res[name].detach())
        assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (name, a, b)
    assert abs(float(scal[SCALAR_INDEX["num_step"]]) - float(res["num_step"])) < 1e-6
    assert abs(float(bmean) - float(baseline.mean())) < 1e-6
    res["loss"].backward()
    ref = O.flatten_params(ocfg, {k: v.grad for k, v in p.items()})
    U.assert_close(grad, ref, atol=1e-5, rtol=1e-4, name="sharded gradient")


def test_shard_ranges_cover_the_batch_without_overlap():
    for n in (0, 1, 5, 8, 4096, 32768):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a0, b0), (a1, b1) in zip(edges, edges[1:]):
                assert b0 == a1 and b0 - a0 >= b1 - a1 >= 0
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)
