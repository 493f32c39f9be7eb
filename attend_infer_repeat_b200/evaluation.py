"""Mirror of the reference's evaluation.py: the progress figure and the scalar logger of the training script
(evaluation.py:15-65 rect / rect_stn / make_fig, :68-108 make_logger, :110-150 make_expr_logger).

matplotlib is not part of this image and TF summaries are out of scope, so ``make_fig`` produces the figure's DATA -- every
array and rectangle the reference draws, per sample and step -- and writes it as ``progress_fig_<step>.npz`` (or renders the
PNG when matplotlib is importable); ``make_logger`` averages the reference's scalar set over a number of batches and prints
the reference's log line.  Host logic only: the numbers come from the model's device buffers.
"""
from __future__ import annotations

import os
import time
from typing import Callable, Dict, Optional

import numpy as np
import torch


def rect_stn_bbox(width, height, stn_params):
    """evaluation.py:23-28 rect_stn: the glimpse rectangle of where = (sx, tx, sy, ty) on a width x height image as the
    reference's bbox [y - .5, x - .5, height * sy, width * sx] (top, left, height, width in pixel-edge coordinates)."""
    sx, tx, sy, ty = (float(v) for v in stn_params)
    x = width * (1. - sx + tx) / 2
    y = height * (1. - sy + ty) / 2
    return [y - .5, x - .5, height * sy, width * sx]


def figure_data(air, n_samples=10) -> Dict[str, np.ndarray]:
    """What make_fig fetches and draws (evaluation.py:31-56) for the model's CURRENT batch: obs [bs,H,W], canvas [T,bs,H,W],
    glimpse [T,bs,h,w], prob = q(n)[..., 1:] [bs,T], presence [T,bs], where [T,bs,4], and per (step, sample) the rectangle
    the reference overlays when presence > .5 (NaN otherwise)."""
    T = air.max_steps
    bs = min(int(n_samples), air.batch_size)
    height, width = air.img_size
    obs = air.obs[:bs].detach().cpu().numpy()
    canvas = air.canvas[:, :bs].detach().cpu().numpy()
    glimpse = air.glimpse[:, :bs].detach().cpu().numpy()
    prob = air.num_steps_distrib.prob()[:bs, 1:].detach().cpu().numpy()
    pres = air.presence[:, :bs, 0].detach().cpu().numpy()
    where = air.where[:, :bs].detach().cpu().numpy()
    bbox = np.full((T, bs, 4), np.nan, dtype=np.float32)
    for i in range(T):
        for j in range(bs):
            if pres[i, j] > .5:
                bbox[i, j] = rect_stn_bbox(width, height, where[i, j])
    titles = np.array([["{:d} with p({:d}) = {:.02f}".format(int(pres[i, j]), i + 1, float(prob[j, i])) for j in range(bs)]
                       for i in range(T)])
    return dict(obs=obs, canvas=canvas, glimpse=glimpse, prob=prob, presence=pres, where=where, bbox=bbox, titles=titles)


def make_fig(air, checkpoint_dir=None, global_step=None, n_samples=10):
    """evaluation.py:31-65.  Returns the figure data; with a checkpoint_dir also writes progress_fig_<global_step>.npz (and
    the PNG, laid out like the reference's, when matplotlib is available)."""
    d = figure_data(air, n_samples)
    if checkpoint_dir is None:
        return d
    os.makedirs(checkpoint_dir, exist_ok=True)
    np.savez_compressed(os.path.join(checkpoint_dir, "progress_fig_{}.npz".format(global_step)), **d)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        from matplotlib.patches import Rectangle
    except ImportError:
        return d
    T, bs = d["canvas"].shape[:2]
    scale = 1.5
    fig, axes = plt.subplots(2 * T + 1, bs, figsize=scale * np.asarray((bs, 2 * T + 1)), squeeze=False)
    for j, ax in enumerate(axes[0]):
        ax.imshow(d["obs"][j], cmap="gray", vmin=0, vmax=1)
    for i in range(T):
        for j in range(bs):
            ax = axes[1 + i][j]
            ax.imshow(d["canvas"][i, j], cmap="gray", vmin=0, vmax=1)
            if not np.isnan(d["bbox"][i, j, 0]):
                b = d["bbox"][i, j]
                ax.add_patch(Rectangle((b[1], b[0]), b[3], b[2], linewidth=3, edgecolor="r", facecolor="none"))
            ax2 = axes[1 + T + i][j]
            ax2.imshow(d["glimpse"][i, j], cmap="gray")
            ax2.set_title(str(d["titles"][i, j]), fontsize=4 * scale)
    for ax in axes.flatten():
        ax.xaxis.set_visible(False)
        ax.yaxis.set_visible(False)
    fig.savefig(os.path.join(checkpoint_dir, "progress_fig_{}.png".format(global_step)), dpi=300)
    plt.close("all")
    return d


def logged_expressions(air) -> Dict[str, Callable[[], torch.Tensor]]:
    """The scalar set of make_logger (evaluation.py:69-92), as callables evaluated after each forward."""
    exprs = {"loss": lambda: air.loss.value, "rec_loss": lambda: air.rec_loss,
             "num_step_acc": lambda: air.num_step_accuracy, "num_step": lambda: air.num_step}
    if air.use_prior:
        exprs["prior_loss"] = lambda: air.prior_loss.value
        if air.num_steps_prior is not None:
            exprs["kl_num_steps"] = lambda: air.kl_num_steps
        if air.what_prior is not None:
            exprs["kl_what"] = lambda: air.kl_what
            exprs["kl_where"] = lambda: air.kl_where
    if air.use_reinforce:
        if air.baseline is not None:
            exprs["baseline_loss"] = lambda: air.baseline_loss
        exprs["reinforce_loss"] = lambda: air.reinforce_loss
        # tf.reduce_mean over the [B] - [B,1] broadcast = mean(iw) - mean(baseline): two entries of the scalar block
        exprs["imp_weight"] = lambda: air.engine.scalar("mean_iw") - air.engine.scalar("mean_baseline")
    return exprs


def make_expr_logger(air, num_batches, exprs, name, next_batch: Optional[Callable] = None, measure_time=True, out=print):
    """evaluation.py:110-150: average every expression over `num_batches` evaluations of the model (each on the batch
    `next_batch()` returns -- (imgs, nums) -- or on the model's current batch) and print the reference's log line."""
    def logger(itr=0, num_batches_to_eval=None, write=True):
        n = int(num_batches if num_batches_to_eval is None else num_batches_to_eval)
        acc = {k: 0. for k in exprs}
        start = time.time()
        for _ in range(max(n, 1)):
            if next_batch is not None:
                air.forward(*next_batch())
            else:
                air.forward()
            for k, f in exprs.items():
                acc[k] += float(f())
        acc = {k: v / max(n, 1) for k, v in acc.items()}
        line = "Step {}, Data {} ".format(itr, name) + ", ".join("{} = {:.4f}".format(k, v) for k, v in acc.items())
        if measure_time:
            line += ", eval time = {:.4}s".format(time.time() - start)
        if write and out is not None:
            out(line)
        return acc
    return logger


def make_logger(air, train_batch: Callable, train_batches, test_batch: Callable, test_batches, out=print):
    """evaluation.py:68-108: log(train_itr) evaluates the scalar set on `train_batches` training batches and `test_batches`
    validation batches (callables returning (imgs, nums) device tensors) and prints both lines; returns both dicts."""
    exprs = logged_expressions(air)
    train_log = make_expr_logger(air, train_batches, exprs, "train", train_batch, out=out)
    test_log = make_expr_logger(air, test_batches, exprs, "test", test_batch, out=out)

    def log(train_itr, num_batches_to_eval=None):
        return dict(train=train_log(train_itr, num_batches_to_eval), test=test_log(train_itr, num_batches_to_eval))
    return log
