"""Mirror of the reference's prior.py: step-count prior / posterior algebra, computed by the CUDA library
(float64 islands included) -- see functional.py for the C-ABI calls."""
from __future__ import annotations

import torch

from . import functional as F
from .functional import bernoulli_to_modified_geometric, geometric_prior, tabular_kl  # noqa: F401  (re-exported)


def sample_from_tensor(tensor, idx):
    """prior.py:103-116: ``tensor[b, int(idx[b])]`` for minibatches (2-D only)."""
    if tensor.dim() > 2:
        raise NotImplementedError
    return F.sample_from_tensor(tensor, idx)


class NumStepsDistribution:
    """prior.py:119-151: turns Bernoulli 'take another step' probabilities [B,T] into p(n), n = 0..T."""

    def __init__(self, steps_probs, joint=None):
        self._steps_probs = steps_probs
        self._joint = joint if joint is not None else bernoulli_to_modified_geometric(steps_probs)

    def sample(self, n=None):
        shape = tuple(self._steps_probs.shape) if n is None else (n,) + tuple(self._steps_probs.shape)
        u = torch.rand(shape, device=self._steps_probs.device)
        s = (u < self._steps_probs).to(torch.float32)
        s = torch.cumprod(s, dim=-1)
        return s.sum(-1)

    def prob(self, samples=None):
        if samples is None:
            return self._joint
        return sample_from_tensor(self._joint, samples)

    def log_prob(self, samples):
        return F.num_steps_log_prob(self._joint, samples)
