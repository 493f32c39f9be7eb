"""Synthetic multi-MNIST-shaped batches in the reference's data format (data/data.py:35-118).

The reference pickles ``imgs uint8 [N,50,50]``, ``labels``, ``nums uint8 [3,N,1]`` and feeds float32 / 255 images
(data.py:116).  There is no MNIST in this image (no network), so benchmarks and tests use stroke-like blobs with the
same statistics: 0..2 objects in tight ~20x20 boxes, no overlap, background exactly 0, foreground = uint8 / 255.
Host-side numpy only; nothing here is on the timed path.
"""
from __future__ import annotations

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
