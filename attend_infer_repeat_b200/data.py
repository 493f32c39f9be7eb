"""Synthetic multi-MNIST-shaped batches in the reference's data format (data/data.py:35-118).

The reference pickles ``imgs uint8 [N,50,50]``, ``labels``, ``nums uint8 [3,N,1]`` and feeds float32 / 255 images
(data.py:116).  There is no MNIST in this image (no network), so benchmarks and tests use stroke-like blobs with the
same statistics: 0..2 objects in tight ~20x20 boxes, no overlap, background exactly 0, foreground = uint8 / 255.

Also here: the reference's data API (``load_data`` / ``tensors_from_data``, data.py:110-158) and a device-resident
variant of it (``ResidentDataset``): the uint8 dataset lives in HBM, a minibatch is a vector of indices, and gather +
/255 + operand preparation run inside the CUDA library (air_forward_dataset_u8 / air_gather_u8), so the host leaves the
training loop.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch


def synthetic_multi_mnist_u8(n: int, H: int = 50, W: int = 50, seed: int = 0, max_objects: int = 2):
    """Returns (imgs uint8 [n,H,W], nums uint8 [max_objects+1, n, 1]) like create_mnist (data.py:35-107)."""
    rng = np.random.default_rng(seed)
    imgs = np.zeros((n, H, W), dtype=np.uint8)
    nums = np.zeros((max_objects + 1, n, 1), dtype=np.uint8)
    s = max(4, int(round(20 * H / 50)))
    yy, xx = np.mgrid[0:s, 0:s].astype(np.float32)
    for b in range(n):
        k = int(rng.integers(0, max_objects + 1))
        boxes = []
        for _ in range(50):
            if len(boxes) == k:
                break
            y0, x0 = int(rng.integers(0, H - s + 1)), int(rng.integers(0, W - s + 1))
            if any(abs(y0 - py) < s and abs(x0 - px) < s for py, px in boxes):
                continue
            boxes.append((y0, x0))
            pts = rng.uniform(0.15 * s, 0.85 * s, size=(4, 2)).astype(np.float32)
            d = np.full((s, s), 1e9, dtype=np.float32)
            for j in range(3):
                p, v = pts[j], pts[j + 1] - pts[j]
                tt = np.clip(((xx - p[0]) * v[0] + (yy - p[1]) * v[1]) / max(float(v @ v), 1e-6), 0, 1)
                d = np.minimum(d, np.hypot(xx - (p[0] + tt * v[0]), yy - (p[1] + tt * v[1])))
            blob = (np.clip(1.6 - d / (0.06 * s), 0, 1) * 255).astype(np.uint8)
            imgs[b, y0:y0 + s, x0:x0 + s] = np.maximum(imgs[b, y0:y0 + s, x0:x0 + s], blob)
        nums[:len(boxes), b, 0] = 1
    return imgs, nums


def synthetic_multi_mnist(n: int, H: int = 50, W: int = 50, seed: int = 0, max_objects: int = 2):
    """float32 images in [0,1] (load_data, data.py:110-118) and float32 nums [max_objects+1, n, 1]."""
    imgs, nums = synthetic_multi_mnist_u8(n, H, W, seed, max_objects)
    return torch.from_numpy(imgs.astype(np.float32) / 255.0), torch.from_numpy(nums.astype(np.float32))


# ------------------------------------------------------------------------------------------------------------
# the reference's pickle format and loaders (data/data.py:35-158)
# ------------------------------------------------------------------------------------------------------------
def save_data(path, imgs_u8, nums_u8, labels=None):
    """Write the dict create_mnist returns (data.py:35-107): imgs uint8 [N,H,W], labels uint8 [N,max_objects],
    nums uint8 [max_objects+1,N,1] -- pickled with protocol 2 like the Python-2 reference."""
    imgs_u8 = np.ascontiguousarray(imgs_u8, dtype=np.uint8)
    nums_u8 = np.ascontiguousarray(nums_u8, dtype=np.uint8)
    if labels is None:
        labels = np.zeros((imgs_u8.shape[0], nums_u8.shape[0] - 1), dtype=np.uint8)
    with open(path, "wb") as f:
        pickle.dump(dict(imgs=imgs_u8, labels=np.asarray(labels, dtype=np.uint8), nums=nums_u8), f, protocol=2)


def load_raw(path, data_path=None):
    """The pickled dict as stored (uint8); Python-2 pickles of the reference load with encoding='latin1'."""
    if data_path is not None:
        path = os.path.join(data_path, path)
    with open(path, "rb") as f:
        try:
            return pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            return pickle.load(f, encoding="latin1")


def load_data(path, data_path=None):
    """data.py:110-118: imgs -> float32 / 255, nums -> float32 (host numpy arrays, as in the reference)."""
    data = dict(load_raw(path, data_path))
    data["imgs"] = data["imgs"].astype(np.float32) / 255.
    data["nums"] = data["nums"].astype(np.float32)
    return data


def tensors_from_data(data_dict, batch_size, axes=None, shuffle=False, seed=None):
    """data.py:121-158 without the TF py_func: returns a dict of zero-argument callables sharing one index draw per
    ``next_batch()``; ``tensors['next_batch']()`` returns {key: minibatch}.  shuffle=True: np.random.choice with
    replacement (data.py:131-132).  shuffle=False reproduces the reference as written: ``next(itertools.cycle(...))``
    builds a fresh cycle on every call, so the FIRST batch is returned every time (data.py:136-139)."""
    keys = list(data_dict.keys())
    if axes is None:
        axes = {k: 0 for k in keys}
    n_entries = data_dict[keys[0]].shape[axes[keys[0]]]
    rng = np.random.RandomState(seed)

    def idx_fun():
        if shuffle:
            return rng.choice(n_entries, batch_size)
        return np.arange(0, batch_size)

    def next_batch():
        idx = idx_fun()
        return {k: data_dict[k].take(idx, axes[k]) for k in keys}

    return dict(next_batch=next_batch, keys=keys, n_entries=n_entries)


class ResidentDataset:
    """The uint8 dataset of data.py:35-107 resident in HBM.  ``next_indices()`` draws the minibatch (with replacement,
    like data.py:131-132) on the device; ``gather()`` / ``Engine.forward_dataset_u8`` turn indices into float32 images
    inside the CUDA library.  50x50 multi-MNIST: 60,000 canvases = 150 MB of the 180 GB."""

    def __init__(self, imgs_u8, nums_u8=None, device="cuda", seed=0):
        imgs = torch.as_tensor(np.ascontiguousarray(imgs_u8))
        if imgs.dtype != torch.uint8 or imgs.dim() != 3:
            raise ValueError("imgs must be uint8 [N,H,W] (the reference's pickle format)")
        self.imgs = imgs.to(device).contiguous()
        self.nums = None if nums_u8 is None else torch.as_tensor(np.ascontiguousarray(nums_u8)).to(device)
        self.n, self.H, self.W = self.imgs.shape
        self.device = self.imgs.device
        self._gen = torch.Generator(device=self.device).manual_seed(seed)

    @classmethod
    def from_pickle(cls, path, data_path=None, **kw):
        raw = load_raw(path, data_path)
        return cls(raw["imgs"], raw.get("nums"), **kw)

    def next_indices(self, batch_size):
        return torch.randint(0, self.n, (batch_size,), device=self.device, dtype=torch.int32, generator=self._gen)

    def gather(self, idx):
        """float32 [B,H,W] = imgs[idx] / 255 (one kernel); nums [n_max+1,B,1] float32 alongside when present."""
        from . import _lib
        from ._lib import check, current_stream_ptr, ptr
        idx = idx.to(torch.int32).contiguous()
        out = torch.empty(idx.numel(), self.H, self.W, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(_lib.lib().air_gather_u8(ptr(self.imgs), ptr(idx), ptr(out), idx.numel(), self.H * self.W,
                                           current_stream_ptr()), "air_gather_u8")
        return out, self.gather_nums(idx)

    def gather_nums(self, idx):
        if self.nums is None:
            return None
        return self.nums.index_select(1, idx.long()).to(torch.float32)
