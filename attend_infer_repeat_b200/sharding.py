"""Batch sharding of the AIR path across the GPUs of one node.

Every canvas is independent through the whole unroll (no batch-norm, no cross-sample op): rank r owns a contiguous
slice of the batch, parameters and optimiser state are replicated, and there is NO data-path collective.  The only
cross-sample couplings are batch means (model.py:103,151,184,212,248,322) and the REINFORCE [B,B] broadcast
(model.py:242-248), which factors into three means:

    mean_{i,j}((iw_j - baseline_i) * logq_j) = mean(iw * logq) - mean(baseline) * mean(logq)

so the exchange is one all-reduce of 16 floats per pass, plus (training) one all-reduce of the flat gradient buffer.
Nothing here computes on the path: these are host-side helpers around torch.distributed (NCCL on the GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import AIR_N_SCALARS, SCALAR_INDEX

# scalars that are plain batch means of per-sample terms (everything else is derived from them)
_MEAN_SLOTS = ("rec_loss", "kl_num_steps", "kl_what", "kl_where", "num_step", "mean_iw_logq", "mean_logq",
               "mean_baseline", "mean_iw", "mean_iw2", "mean_baseline2")


def shard_range(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [start, stop) of a batch of n_global canvases owned by `rank` (remainder to the low ranks)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(int(n_global), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int, dim: int = 0) -> torch.Tensor:
    """The rank's slice of a batch-major (dim=0) or time-major (dim=1) tensor."""
    a, b = shard_range(t.shape[dim], rank, world)
    return t.narrow(dim, a, b - a)


def combine_scalars(scalars: torch.Tensor, n_local: int, steps_weight: float = 1.0, use_prior: bool = True,
                    use_reinforce: bool = True, group=None, nvil_shift: float = 0.0, nvil_scale: float = 0.0) -> torch.Tensor:
    """Turn the per-shard scalar block of air_forward (16 floats, batch means over the LOCAL shard) into the scalars
    of the whole batch, in place, on every rank.  Means are re-weighted by the shard size (shards may be ragged); the
    derived entries (prior_loss, loss, reinforce_loss, opt_loss; elbo_scalars_kernel) are re-formed from the global
    means exactly as the single-device kernel forms them."""
    import torch.distributed as dist
    assert scalars.numel() == AIR_N_SCALARS
    dev = scalars.device
    # a handful of vector operations (this runs every training step between the forward and the backward pass: one tiny
    # launch per scalar would cost more than the all-reduce)
    mean_idx = _index_tensor(tuple(SCALAR_INDEX[n] for n in _MEAN_SLOTS), dev)
    wire = torch.float32 if dev.type == "cuda" else torch.float64          # NCCL path: fp32 on the wire (16 floats)
    buf = torch.zeros(AIR_N_SCALARS, dtype=wire, device=dev)
    buf[mean_idx] = (scalars[mean_idx].double() * float(n_local)).to(wire)
    buf[AIR_N_SCALARS - 1:].fill_(float(n_local))     # (fill_, not item assignment: no host-to-device copy, capturable)
    dist.all_reduce(buf, group=group)
    m = buf / buf[AIR_N_SCALARS - 1]                                       # global means at the mean slots
    I = SCALAR_INDEX
    prior_loss = m[I["kl_num_steps"]] * steps_weight + m[I["kl_what"]] + m[I["kl_where"]]
    loss = m[I["rec_loss"]] + prior_loss * (1.0 if use_prior else 0.0)
    nv_scale, nv_shift = (nvil_scale, nvil_shift) if nvil_scale != 0.0 else (1.0, 0.0)
    reinforce = (nv_scale * (m[I["mean_iw_logq"]] - (m[I["mean_baseline"]] + nv_shift) * m[I["mean_logq"]])
                 if use_reinforce else torch.zeros_like(loss))
    out = torch.zeros(AIR_N_SCALARS, dtype=m.dtype, device=dev)
    out[mean_idx] = m[mean_idx]
    derived = torch.stack([prior_loss, loss, reinforce, loss + reinforce])
    out[_index_tensor((I["prior_loss"], I["loss"], I["reinforce_loss"], I["opt_loss"]), dev)] = derived
    scalars.copy_(out.to(scalars.dtype))
    return scalars


_INDEX_CACHE = {}


def _index_tensor(idx: tuple, device) -> torch.Tensor:
    key = (idx, str(device))
    t = _INDEX_CACHE.get(key)
    if t is None:
        t = _INDEX_CACHE[key] = torch.tensor(idx, dtype=torch.long, device=device)
    return t


def allreduce_gradient(flat_grad: torch.Tensor, n_local: int, n_global: Optional[int] = None, group=None,
                       async_op: bool = False):
    """The ONE data-path collective of a training step: sum over ranks of the flat gradient buffer.

    Each rank's buffer holds the gradient of ITS shard's mean loss; the gradient of the global mean is the shard-size
    weighted average.  With equal shards (the benchmark configuration) that is all-reduce(sum) / world; ragged shards
    pre-scale by n_local / n_global.  In place; returns the work handle when async_op."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if n_global is None:
        n_global = n_local * world
    flat_grad.mul_(float(n_local) / float(n_global))
    return dist.all_reduce(flat_grad, group=group, async_op=async_op)


def global_baseline_mean(baseline_local: torch.Tensor, n_local: int, group=None) -> torch.Tensor:
    """mean(baseline) over the WHOLE batch (needed by the backward of the REINFORCE term on every rank)."""
    import torch.distributed as dist
    buf = torch.stack([baseline_local.double().sum(), torch.tensor(float(n_local), dtype=torch.float64,
                                                                    device=baseline_local.device)])
    if buf.device.type == "cuda":
        buf = buf.float()
    dist.all_reduce(buf, group=group)
    return (buf[0] / buf[1]).to(baseline_local.dtype)


def bind_host_to_device(device_index: int, min_cpus: int = 4) -> Optional[list]:
    """Bind the calling thread (and the threads it starts afterwards: NCCL's proxy thread, torch's pinned-memory
    allocator) to the CPUs NVML reports as local to GPU ``device_index``, so that the pinned host buffers of the feed
    path are first-touched on the GPU's own NUMA node and its H2D copies do not cross the socket interconnect (round-1
    review: eight ranks feeding through one node lost 22 % end to end).  Call it once per rank before allocating pinned
    memory.  Returns the CPU list, or None when there is nothing sensible to bind to (no NVML, a cgroup that already
    restricts the process to fewer than ``min_cpus`` of those CPUs, ...): never raises."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, ((os.cpu_count() or 64) + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus = sorted(local & os.sched_getaffinity(0))
        if len(cpus) < min_cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
