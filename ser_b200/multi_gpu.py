"""One process per GPU: rendezvous, timing reductions and result gathering.

The data path has no collective: every rank extracts the features of its own contiguous clip
range (``sharding.shard_bounds``).  ``torch.distributed`` is plumbing only -- the barrier and
max-over-ranks of the timing contract, and the gather of the small per-clip feature rows to
rank 0.  The backend is NCCL on GPUs and gloo on CPU (tests), chosen by the caller.
"""

from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from .sharding import shard_bounds


@dataclass(frozen=True)
class RankInfo:
    rank: int
    local_rank: int
    world: int

    @property
    def distributed(self) -> bool:
        return self.world > 1


def rank_info() -> RankInfo:
    """RANK / LOCAL_RANK / WORLD_SIZE as torchrun exports them (single process when absent)."""
    return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
                    int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(info: RankInfo, backend: str, device=None) -> None:
    if not info.distributed:
        return
    import torch.distributed as dist

    if dist.is_initialized():
        return
    kwargs = {}
    if device is not None and backend == "nccl":
        kwargs["device_id"] = device
    dist.init_process_group(backend, rank=info.rank, world_size=info.world, **kwargs)


def destroy_process_group(info: RankInfo) -> None:
    if not info.distributed:
        return
    import torch.distributed as dist

    if dist.is_initialized():
        dist.destroy_process_group()


def barrier(info: RankInfo) -> None:
    if info.distributed:
        import torch.distributed as dist

        dist.barrier()


def max_over_ranks(info: RankInfo, value: float, device="cpu") -> float:
    """The slowest rank's figure: what the timing contract reports."""
    if not info.distributed:
        return float(value)
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def my_clip_range(info: RankInfo, lengths: np.ndarray) -> tuple[int, int]:
    """This rank's contiguous [lo, hi) slice of a globally known clip list (strong scaling)."""
    return shard_bounds(lengths, info.world)[info.rank]


def gather_rows(info: RankInfo, local_rows: np.ndarray, n_total: int) -> np.ndarray | None:
    """Concatenates every rank's (n_i, dim) block in rank order on rank 0 (None elsewhere)."""
    local_rows = np.ascontiguousarray(local_rows)
    if not info.distributed:
        assert local_rows.shape[0] == n_total
        return local_rows
    import torch.distributed as dist

    blocks: list = [None] * info.world if info.rank == 0 else None
    dist.gather_object(local_rows, blocks, dst=0)
    if info.rank != 0:
        return None
    out = np.concatenate(blocks, axis=0)
    if out.shape[0] != n_total:
        raise RuntimeError(f"gathered {out.shape[0]} rows, expected {n_total}")
    return out
