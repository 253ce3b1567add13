"""Device helpers mirroring src/deeperwin/utils/utils.py:28-30, 54-112 for a one-process-per-GPU world:
`pmean`/`psum` are NCCL all-reduces over torch.distributed (identity when not initialised)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def psum(x: torch.Tensor) -> torch.Tensor:
    """jax.lax.psum over the "devices" axis (utils.py:30)."""
    if world_size() > 1:
        x = x.clone()
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
    return x


def pmean(x: torch.Tensor) -> torch.Tensor:
    """jax.lax.pmean over the "devices" axis (utils.py:29)."""
    if world_size() > 1:
        x = x.clone()
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
        x = x / world_size()
    return x


def all_gather_batch(x: torch.Tensor) -> torch.Tensor:
    """merge_from_devices (utils.py:100-112): the reference emulates this gather with a psum of a
    zero-padded array; here it is a real all-gather along the walker axis."""
    if world_size() == 1:
        return x
    xs = x.contiguous()
    if xs.dtype == torch.uint32:
        out = torch.empty((world_size() * xs.shape[0],) + tuple(xs.shape[1:]), dtype=torch.int32, device=xs.device)
        dist.all_gather_into_tensor(out, xs.view(torch.int32))
        return out.view(torch.uint32)
    out = torch.empty((world_size() * xs.shape[0],) + tuple(xs.shape[1:]), dtype=xs.dtype, device=xs.device)
    dist.all_gather_into_tensor(out, xs)
    return out


def flat_allreduce_mean(tensors):
    """One flat-buffer all-reduce for gradients / KFAC factors (optimizers.py:133; custom_kfac_jax
    optimizer.py:1151, curvature_blocks.py:1187-1199): packs, reduces once, unpacks in place."""
    if world_size() == 1 or not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world_size()
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors
