"""Instance sharding across the GPUs of one box.

Every slice is an independent optimisation (src/test/test_immoco.py:45-72 loops slice by slice and
src/models/immoco.py:134 builds a fresh model per call), so slices are dealt round-robin to ranks
with NO data-path collective; the only exchange is one gather of the corrected images at the end
(NCCL over NVLink on GPUs; the same code runs on gloo for the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_slices: int, rank: int, world: int) -> List[int]:
    """Slices s with s % world == rank (round-robin keeps ranks within one slice of each other)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    return list(range(rank, n_slices, world))


def gather_images(local: torch.Tensor, n_slices: int, rank: int, world: int, dst: int = 0) -> Optional[torch.Tensor]:
    """local: (S_local, H, W) complex64 results of this rank's shard, in shard order.
    Returns the (n_slices, H, W) stack in ORIGINAL slice order on rank ``dst``, None elsewhere."""
    if world == 1:
        return local
    h, w = local.shape[-2:]
    per_rank = (n_slices + world - 1) // world
    padded = torch.zeros((per_rank, h, w, 2), dtype=torch.float32, device=local.device)
    if local.shape[0]:
        padded[: local.shape[0]] = torch.view_as_real(local.to(torch.complex64))
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_slices, h, w, 2), dtype=torch.float32, device=local.device)
    for r in range(world):
        idx = shard_indices(n_slices, r, world)
        if idx:
            out[idx] = bufs[r][: len(idx)]
    return torch.view_as_complex(out)


def reconstruct_slices(kspaces: Sequence[torch.Tensor], masks: Sequence[torch.Tensor], *, iters: int = 200,
                       learning_rate: float = 1e-2, lambda_ge: float = 1e-2,
                       reconstruct_fn: Optional[Callable] = None, in_flight: Optional[int] = None,
                       deterministic: Optional[bool] = None, batch: Optional[int] = None) -> Optional[torch.Tensor]:
    """Reconstruct every slice of a stack, sharded over the ranks of the default process group.

    By default each rank runs its shard through ``reconstruct_batch`` (several slices in flight per
    GPU).  ``reconstruct_fn(kspace, masks, iters, learning_rate, lambda_ge, False) -> (image,
    kspace_fwd)`` replaces the per-slice fit (the gloo tests inject a CPU stand-in).  All slices must
    share (H, W)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = len(kspaces)
    mine = shard_indices(n, rank, world)
    if reconstruct_fn is None:
        from .batch import DEFAULT_BATCH, DEFAULT_IN_FLIGHT, reconstruct_batch
        outs = reconstruct_batch([kspaces[s] for s in mine], [masks[s] for s in mine], iters, learning_rate,
                                 lambda_ge, in_flight=in_flight or DEFAULT_IN_FLIGHT, deterministic=deterministic,
                                 batch=batch or DEFAULT_BATCH) if mine else []
    else:
        outs = [reconstruct_fn(kspaces[s], masks[s], iters, learning_rate, lambda_ge, False)[0] for s in mine]
    if outs:
        local = torch.stack([o.detach() for o in outs])
    else:
        h, w = kspaces[0].shape[-2:]
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        local = torch.zeros((0, h, w), dtype=torch.complex64, device=dev)
    return gather_images(local, n, rank, world)
