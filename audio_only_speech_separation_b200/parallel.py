"""Host-side data-parallel plumbing (one process per GPU, torch.distributed).

The path shards by utterance: every utterance is independent through the whole network (GroupNorm statistics are
per sample) and the loss is a batch mean, so inference needs no collective at all and training needs exactly one:
a SUM all-reduce of the flat gradient buffer, scaled by 1/world (the reference's Lightning ``DDPStrategy``,
audio_train.py:22,126).  These helpers are backend-agnostic so the logic is testable with ``gloo`` on CPU.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced shard ``[lo, hi)`` of ``n_items`` utterances for ``rank`` (sizes differ by at most 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a flat gradient buffer followed by the 1/world scale."""
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / dist.get_world_size(group))
    return flat


def max_over_ranks(value: float, device=None, group=None) -> float:
    """MAX all-reduce of a host scalar (timing is reported as the slowest rank)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def init_from_env(backend: str = "nccl"):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local
