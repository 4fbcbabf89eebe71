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


def bind_host_to_gpu(local_rank: int) -> dict | None:
    """Moves the calling thread (and the threads it starts later) onto the CPU cores NVML names as
    closest to GPU ``local_rank``, so that the pinned buffers this rank allocates afterwards -- and
    the staging threads of ``csrc/host_stage.h`` -- sit on the GPU's own NUMA node.  With eight
    ranks each pulling ~0.5 GB per step out of host memory, copies that cross the socket
    interconnect are what bends the end-to-end scaling curve.  ``SERB_NUMA_BIND=0`` turns it off.
    Returns what was done (for the bench line), or None when disabled."""
    if os.environ.get("SERB_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_getaffinity"):
        return None
    before = sorted(os.sched_getaffinity(0))
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        if all(hasattr(props, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        else:
            visible = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(visible[local_rank]) if visible else local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
    except Exception as exc:  # no NVML, a cpuset that excludes the ideal cores, ...: stay where we are
        return {"bound": False, "why": f"{type(exc).__name__}: {exc}"[:120], "cpus": len(before)}
    after = sorted(os.sched_getaffinity(0))
    return {"bound": after != before, "cpus": len(after), "cpus_before": len(before),
            "first_cpu": after[0] if after else None, "restore": before}


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


def gather_rows(info: RankInfo, local_rows: np.ndarray, n_total: int, *, counts: list[int] | None = None,
                device=None) -> np.ndarray | None:
    """Concatenates every rank's (n_i, dim) block in rank order on rank 0 (None elsewhere).

    One tensor collective (``dist.gather`` of blocks padded to the largest rank's row count), not an
    object collective: no pickling, and on NCCL the blocks travel GPU to GPU.  ``counts`` (rows per rank)
    is known to callers that sharded the work themselves; when absent it is exchanged first.
    ``device`` is where the collective runs: "cuda" for NCCL, "cpu" for gloo (default: by backend)."""
    local_rows = np.ascontiguousarray(local_rows)
    if not info.distributed:
        assert local_rows.shape[0] == n_total
        return local_rows
    import torch
    import torch.distributed as dist

    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if counts is None:
        counts = [0] * info.world
        dist.all_gather_object(counts, int(local_rows.shape[0]))
    if sum(counts) != n_total:
        raise RuntimeError(f"ranks hold {sum(counts)} rows, expected {n_total}")
    most = max(counts)
    block = torch.zeros((most,) + local_rows.shape[1:], dtype=torch.from_numpy(local_rows[:0]).dtype, device=device)
    block[: local_rows.shape[0]].copy_(torch.from_numpy(local_rows), non_blocking=False)
    blocks = [torch.empty_like(block) for _ in range(info.world)] if info.rank == 0 else None
    dist.gather(block, blocks, dst=0)
    if info.rank != 0:
        return None
    return np.concatenate([b[:c].cpu().numpy() for b, c in zip(blocks, counts)], axis=0)


class RowGatherer:
    """Double-buffered gather of per-rank row blocks to rank 0, off the critical path.

    ``submit(rows)`` starts the gather of this step's rows (pinned staging -> device block -> one
    ``dist.gather`` -> one device-to-host copy on rank 0, all on a side stream) and returns a ticket;
    ``collect(ticket)`` waits for it and returns the concatenated rows on rank 0 (None elsewhere).
    A caller that collects step i after submitting step i + 1 overlaps the exchange -- and rank 0's
    wait for the slowest rank -- with the next step's compute; every row still reaches rank 0 inside
    the caller's timed region as long as the last ticket is collected before the clock stops.
    On gloo (CPU tests) the same protocol runs on ``async_op`` work handles.  Counts (rows per rank)
    and the row shape are fixed at construction; two tickets may be in flight."""

    def __init__(self, info: RankInfo, counts: list[int], row_shape: tuple[int, ...], dtype=np.float64,
                 device: str | None = None) -> None:
        self.info = info
        self.counts = [int(c) for c in counts]
        self.row_shape = tuple(int(d) for d in row_shape)
        self.dtype = np.dtype(dtype)
        self._submitted = 0
        self._slots: list[dict] = []
        if not info.distributed:
            self._slots = [{"rows": None}, {"rows": None}]
            return
        import torch
        import torch.distributed as dist

        if device is None:
            device = "cuda" if dist.get_backend() == "nccl" else "cpu"
        self.device = device
        self.cuda = device != "cpu"
        tdtype = torch.from_numpy(np.zeros(0, dtype=self.dtype)).dtype
        most = max(self.counts)
        shape = (most,) + self.row_shape
        self.side = torch.cuda.Stream() if self.cuda else None
        for _ in range(2):
            slot = {
                "host_in": torch.zeros(shape, dtype=tdtype, pin_memory=self.cuda),
                "block": torch.zeros(shape, dtype=tdtype, device=device),
                "all": torch.zeros((info.world,) + shape, dtype=tdtype, device=device) if info.rank == 0 else None,
                "host_out": torch.zeros((info.world,) + shape, dtype=tdtype, pin_memory=self.cuda) if info.rank == 0 else None,
                "event": torch.cuda.Event() if self.cuda else None,
                "work": None,
                "busy": False,
            }
            self._slots.append(slot)

    def submit(self, local_rows: np.ndarray) -> int:
        ticket = self._submitted
        self._submitted += 1
        slot = self._slots[ticket & 1]
        local_rows = np.ascontiguousarray(local_rows, dtype=self.dtype)
        if not self.info.distributed:
            slot["rows"] = local_rows
            return ticket
        import torch
        import torch.distributed as dist

        if slot["busy"]:
            raise RuntimeError("RowGatherer: collect the ticket two steps back before submitting again")
        n = self.counts[self.info.rank]
        if local_rows.shape != (n,) + self.row_shape:
            raise ValueError(f"rank {self.info.rank} submits {local_rows.shape}, expected {(n,) + self.row_shape}")
        slot["host_in"][:n].copy_(torch.from_numpy(local_rows))
        gather_list = list(slot["all"].unbind(0)) if self.info.rank == 0 else None
        if self.cuda:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                slot["block"].copy_(slot["host_in"], non_blocking=True)
                dist.gather(slot["block"], gather_list, dst=0)
                if self.info.rank == 0:
                    slot["host_out"].copy_(slot["all"], non_blocking=True)
                slot["event"].record(self.side)
        else:
            slot["block"].copy_(slot["host_in"])
            slot["work"] = dist.gather(slot["block"], gather_list, dst=0, async_op=True)
        slot["busy"] = True
        return ticket

    def collect(self, ticket: int) -> np.ndarray | None:
        slot = self._slots[ticket & 1]
        if not self.info.distributed:
            return slot["rows"]
        if not slot["busy"]:
            raise RuntimeError("RowGatherer: ticket already collected")
        if self.cuda:
            slot["event"].synchronize()
        else:
            slot["work"].wait()
            if self.info.rank == 0:
                slot["host_out"].copy_(slot["all"])
        slot["busy"] = False
        if self.info.rank != 0:
            return None
        out = slot["host_out"].numpy()
        return np.concatenate([out[r, :c] for r, c in enumerate(self.counts)], axis=0)
