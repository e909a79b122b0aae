"""Clip-parallel multi-GPU plumbing: one process per GPU, independent clips sharded across
ranks, weights broadcast ONCE at init (NCCL over NVLink 5 / NVSwitch), no collective in the step loop.

The reference's inference is single-process (scripts/sampling/sampling_tv2v.py:106); `num_samples` only replicates
prompts into a serial loop (:174-178, :289-291).  Clips (and samples of a clip) never interact - every norm is
per-sample / per-frame - so rank r of W simply takes the chunk indices i with i % W == r.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, local_rank, world).  world == 1 without the env: nothing is initialised."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_clips(n_clips: int, rank: int, world: int) -> List[int]:
    """Indices of the clips (chunks of prompts_chunk, sampling_tv2v.py:289-291) rank `rank` processes."""
    return list(range(rank, n_clips, world))


def broadcast_weights(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 256 << 20) -> int:
    """Broadcast every parameter and buffer of `module` from rank `src` (the only collective of the whole job).
    Tensors are coalesced into flat buckets of ~bucket_bytes so the broadcast is bandwidth- not latency-bound
    (1 550 tensors for tv2v).  Returns the number of bytes broadcast.  No-op when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    total = 0
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    for (dtype, device), group in by_dtype.items():
        bucket, size = [], 0
        for t in group + [None]:
            if t is not None:
                bucket.append(t)
                size += t.numel() * t.element_size()
            if bucket and (t is None or size >= bucket_bytes):
                flat = torch.cat([b.reshape(-1) for b in bucket])
                dist.broadcast(flat, src=src)
                off = 0
                for b in bucket:
                    b.copy_(flat[off:off + b.numel()].view_as(b))
                    off += b.numel()
                total += size
                bucket, size = [], 0
    # the copies above go through `.data` (no version bump): drop packed fp16 copies / captured graphs explicitly
    if hasattr(module, "invalidate"):
        module.invalidate()
    elif hasattr(module, "invalidate_packed"):
        module.invalidate_packed()
    return total


def max_over_ranks(value: float, device) -> float:
    """Max of a per-rank scalar (device time of the timed region) over all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def gather_results(local: Sequence, world: int):
    """Gather per-rank python objects (e.g. output file names) on every rank, in clip order."""
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return list(local)
    parts = [None] * world
    dist.all_gather_object(parts, list(local))
    n = sum(len(p) for p in parts)
    out = [None] * n
    for r, p in enumerate(parts):
        for j, item in enumerate(p):
            out[r + j * world] = item
    return out
