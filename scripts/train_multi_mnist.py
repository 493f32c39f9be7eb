#!/usr/bin/env python
"""B200 counterpart of the reference's training entry point (scripts/train_multi_mnist.sh -> scripts/multi_mnist.py).

Same flow and hyper-parameters (multi_mnist.py:24-94): AIRonMNIST on 50x50 multi-MNIST, max_steps = 3, 20x20 glimpses,
RMSProp(momentum=.9, centered=True), annealed geometric step prior, NVIL baseline at 10x lr, log every --log-every
iterations (the scalars of evaluation.py:68-92), checkpoint every --save-every (params + optimiser slots + global_step).
The dataset is the reference's pickle (data.py:35-107: uint8 imgs [N,50,50], nums [3,N,1]) held resident in HBM; with
no --data-path a synthetic multi-MNIST-shaped set is generated (no network in this image).  Multi-GPU:
    python -m torch.distributed.run --nproc-per-node N scripts/train_multi_mnist.py ...
shards every batch across the ranks (one gradient all-reduce per step).
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import attend_infer_repeat_b200 as air                                   # noqa: E402
from attend_infer_repeat_b200.data import ResidentDataset, synthetic_multi_mnist_u8   # noqa: E402


def B_valid(args):
    return args.batch_size


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-path", default=None, help="directory with mnist_train.pickle / mnist_validation.pickle")
    ap.add_argument("--batch-size", type=int, default=64, help="per GPU (multi_mnist.py:26)")
    ap.add_argument("--iters", type=int, default=300000, help="multi_mnist.py:131")
    ap.add_argument("--learning-rate", type=float, default=1e-4, help="multi_mnist.py:24 (the baseline trains at 10x)")
    ap.add_argument("--l2-weight", type=float, default=0.0)
    ap.add_argument("--log-every", type=int, default=10000)
    ap.add_argument("--save-every", type=int, default=5000)
    ap.add_argument("--checkpoint-dir", default="checkpoints")
    ap.add_argument("--resume", default=None)
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--n-synthetic", type=int, default=20000, help="synthetic canvases when no --data-path is given")
    ap.add_argument("--log-json", default=None, help="append one JSON line per log point to this file")
    args = ap.parse_args(argv)

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
    # every rank draws its own where / what / presence noise (AIRCell.draw_noise uses torch's CUDA generator)
    torch.cuda.manual_seed(args.seed * 1000003 + rank)

    if args.data_path:
        train = ResidentDataset.from_pickle("mnist_train.pickle", args.data_path, device=dev, seed=args.seed + rank)
        valid = ResidentDataset.from_pickle("mnist_validation.pickle", args.data_path, device=dev, seed=1)
    else:
        train = ResidentDataset(*synthetic_multi_mnist_u8(args.n_synthetic, 50, 50, seed=0), device=dev,
                                seed=args.seed + rank)
        valid = ResidentDataset(*synthetic_multi_mnist_u8(max(B_valid(args), args.n_synthetic // 10), 50, 50, seed=1),
                                device=dev, seed=1)

    B = args.batch_size
    idx = train.next_indices(B)
    imgs, nums = train.gather(idx)
    n_hiddens = [256, 256]
    model = air.AIRonMNIST(imgs, nums, max_steps=3, explore_eps=1e-3, inpt_encoder_hidden=n_hiddens,
                           glimpse_encoder_hidden=n_hiddens, glimpse_decoder_hidden=n_hiddens,
                           transform_estimator_hidden=n_hiddens, steps_pred_hidden=[128, 64], baseline_hidden=[256, 128],
                           transform_var_bias=.5, step_bias=.75, output_multiplier=.5, seed=args.seed,
                           precision=air.AIR_PREC_TC_SPLIT if args.precision == "tc" else air.AIR_PREC_FP32)
    prior = dict(loc=0., scale=1.)
    num_steps_prior = dict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3,
                           analytic=True)
    train_op, global_step = model.train_step(args.learning_rate, args.l2_weight, prior, prior, prior, num_steps_prior)

    if args.resume:
        ck = torch.load(args.resume, map_location=dev)
        model.params.copy_(ck["params"])
        for k in ("mg", "ms", "mom"):
            model._slots[k].copy_(ck["slots"][k])
        if ck.get("baseline") is not None and model.baseline_module is not None:
            model.baseline_module.params.copy_(ck["baseline"]["params"])
            for k in ("mg", "ms", "mom"):
                model.baseline_module.slots[k].copy_(ck["baseline"]["slots"][k])
        model.global_step = int(ck["global_step"])
        if "rng" in ck:     # noise stream, minibatch stream and validation stream continue where the checkpoint left them
            torch.cuda.set_rng_state(ck["rng"]["cuda"].cpu(), dev)
            train._gen.set_state(ck["rng"]["train"].cpu())
            valid._gen.set_state(ck["rng"]["valid"].cpu())

    def scalars():
        names = ["loss", "rec_loss", "num_step_acc", "num_step", "prior_loss", "kl_num_steps", "kl_what", "kl_where",
                 "baseline_loss", "reinforce_loss"]
        vals = [model.loss.value, model.rec_loss, model.num_step_accuracy, model.num_step, model.prior_loss.value,
                model.kl_num_steps, model.kl_what, model.kl_where, model.baseline_loss, model.reinforce_loss]
        return {n: float(v) for n, v in zip(names, vals)}

    def log(itr):
        out = {"train": scalars()}
        imgs_v, nums_v = valid.gather(valid.next_indices(B))
        model.forward(imgs_v, nums_v)
        out["test"] = scalars()
        if rank == 0:
            for k, v in out.items():
                print(f"Step {itr}, Data {k} " + ", ".join(f"{n} = {x:.4f}" for n, x in v.items()), flush=True)
            if args.log_json:
                import json
                with open(args.log_json, "a") as f:
                    f.write(json.dumps(dict(step=itr, **out)) + "\n")
        return out

    def save(itr):
        if rank != 0:
            return
        os.makedirs(args.checkpoint_dir, exist_ok=True)
        bm = model.baseline_module
        torch.save(dict(params=model.params, slots=model._slots, global_step=itr,
                        baseline=None if bm is None or bm.params is None else dict(params=bm.params, slots=bm.slots),
                        rng=dict(cuda=torch.cuda.get_rng_state(dev), train=train._gen.get_state(),
                                 valid=valid._gen.get_state())),
                   os.path.join(args.checkpoint_dir, f"model-{itr}.pt"))

    itr = global_step()
    if rank == 0:
        print(f"Starting training at iter = {itr}", flush=True)
    if itr == 0:
        log(0)
    t0, n0 = time.time(), itr
    while itr < args.iters:
        imgs, nums = train.gather(train.next_indices(B))
        train_op(imgs, nums)
        itr = global_step()
        if itr % args.log_every == 0:
            torch.cuda.synchronize()
            if rank == 0:
                print(f"{(itr - n0) * B * world * 3 / (time.time() - t0):.0f} cell-steps/s", flush=True)
            log(itr)
        if itr % args.save_every == 0:
            save(itr)
            if rank == 0:     # progress figure next to the checkpoint (evaluation.py:31-65; multi_mnist.py:145-146)
                air.evaluation.make_fig(model, args.checkpoint_dir, itr)
    save(itr)
    return model


if __name__ == "__main__":
    _model = main()
    # a captured training step holds NCCL kernels under torchrun: the graphs go before the process group does
    _model.release_graphs()
    import torch.distributed as _dist
    if _dist.is_available() and _dist.is_initialized():
        _dist.destroy_process_group()
