"""Weights of the reference-cell golden vectors (tests/golden/reference_cell_*.npz), shared by the generator
(tools/make_golden.py) and the tests: the 1.78 M parameters of the script configuration are not stored, they are
re-drawn from this recipe.  Pure numpy; each tensor is seeded by its own name, so the order of creation does not matter
(np.random.RandomState streams are frozen across numpy versions)."""
import zlib

import numpy as np


def golden_tensor(name, shape, seed):
    """name: canonical parameter name ('input_encoder.0.w', 'lstm.b', 'lstm.h0', ...); shape: 2-D (rows, cols)."""
    rows, cols = int(shape[0]), int(shape[1])
    rs = np.random.RandomState((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 31 - 1))
    x = rs.standard_normal((rows, cols))
    if name.endswith(".w"):
        return (x / np.sqrt(rows)).astype(np.float32)        # fan-in scaling like the effective reference init (App. C2)
    return (0.1 * x).astype(np.float32)                      # biases and the trainable initial state: small, non-zero


def golden_subset(name, numel, n=1024):
    """Flat indices at which a large gradient tensor is stored in the vectors (all of them for small tensors): a fixed
    pseudo-random subset per tensor name.  The vectors also hold every tensor's L2 norm and maximum magnitude."""
    numel = int(numel)
    if numel <= 4 * n:
        return np.arange(numel)
    rs = np.random.RandomState(zlib.crc32(("subset:" + name).encode()) % (2 ** 31 - 1))
    return np.sort(rs.choice(numel, size=n, replace=False))
