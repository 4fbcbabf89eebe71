"""Host-side logic of the N > 1 path on CPU: shard bounds, and a world_size-2 gloo run of the
same rendezvous / reduction / gather helpers bench.py uses under NCCL (ser_b200/multi_gpu.py).
The per-clip "features" here are a deterministic host function, because no GPU exists in this
container; what is covered is the partition, the order of the gathered rows and the
max-over-ranks timing reduction."""

from __future__ import annotations

import multiprocessing as mp
import socket

import numpy as np
import pytest

from ser_b200.sharding import shard_bounds


def test_shard_bounds_are_contiguous_and_balanced():
    rng = np.random.default_rng(3)
    lengths = rng.integers(2048, 400000, size=1000)
    for n in (1, 2, 4, 8):
        bounds = shard_bounds(lengths, n)
        assert bounds[0][0] == 0 and bounds[-1][1] == lengths.size
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        loads = [int(lengths[lo:hi].sum()) for lo, hi in bounds]
        assert max(loads) - min(loads) <= 2 * int(lengths.max())
    assert shard_bounds(np.asarray([], dtype=np.int64), 4) == [(0, 0)] * 4
    assert shard_bounds(np.asarray([10, 10]), 4)[-1][1] == 2          # more shards than clips
    with pytest.raises(ValueError):
        shard_bounds(lengths, 0)


def _fake_rows(lengths: np.ndarray, lo: int, hi: int) -> np.ndarray:
    idx = np.arange(lo, hi, dtype=np.float64)
    return np.stack([idx, lengths[lo:hi].astype(np.float64), idx * 0.5], axis=1)


def _worker(rank: int, world: int, port: int, queue) -> None:
    import os

    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    from ser_b200 import multi_gpu

    info = multi_gpu.rank_info()
    multi_gpu.init_process_group(info, "gloo")
    try:
        lengths = np.random.default_rng(11).integers(2048, 200000, size=37)
        lo, hi = multi_gpu.my_clip_range(info, lengths)
        rows = _fake_rows(lengths, lo, hi)
        multi_gpu.barrier(info)
        slowest = multi_gpu.max_over_ranks(info, 10.0 + rank)
        gathered = multi_gpu.gather_rows(info, rows, lengths.size)
        # the double-buffered gatherer: two steps in flight, collected one step late
        bounds = multi_gpu.shard_bounds(lengths, world)
        gatherer = multi_gpu.RowGatherer(info, [b - a for a, b in bounds], rows.shape[1:])
        t0 = gatherer.submit(rows)
        t1 = gatherer.submit(rows * 2.0)
        first = gatherer.collect(t0)
        t2 = gatherer.submit(rows * 3.0)
        second, third = gatherer.collect(t1), gatherer.collect(t2)
        if rank == 0:
            assert np.array_equal(first, gathered) and np.array_equal(second, 2.0 * gathered)
            assert np.array_equal(third, 3.0 * gathered)
        else:
            assert first is None and second is None and third is None
        queue.put((rank, lo, hi, slowest, None if gathered is None else gathered.tolist()))
    finally:
        multi_gpu.destroy_process_group(info)


def test_world_size_2_gloo_partition_gather_and_timing():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = {}
    for _ in procs:
        rank, lo, hi, slowest, gathered = queue.get(timeout=120)
        results[rank] = (lo, hi, slowest, gathered)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lengths = np.random.default_rng(11).integers(2048, 200000, size=37)
    assert results[0][0] == 0 and results[0][1] == results[1][0] and results[1][1] == 37
    assert results[0][2] == 11.0 and results[1][2] == 11.0              # max over ranks on both
    assert results[1][3] is None
    np.testing.assert_array_equal(np.asarray(results[0][3]), _fake_rows(lengths, 0, 37))


def test_single_process_helpers_are_no_ops():
    from ser_b200 import multi_gpu

    info = multi_gpu.RankInfo(0, 0, 1)
    multi_gpu.barrier(info)
    assert multi_gpu.max_over_ranks(info, 3.5) == 3.5
    rows = np.zeros((4, 2))
    assert multi_gpu.gather_rows(info, rows, 4) is rows
    gatherer = multi_gpu.RowGatherer(info, [4], (2,))
    assert np.array_equal(gatherer.collect(gatherer.submit(rows + 1.0)), rows + 1.0)


def test_host_binding_is_optional_and_never_raises(monkeypatch):
    """bind_host_to_gpu moves the process next to its GPU where NVML names such cores; without a GPU (this
    container), without NVML or inside a restrictive cpuset it reports why and leaves the affinity alone."""
    import os

    from ser_b200 import multi_gpu

    before = os.sched_getaffinity(0)
    monkeypatch.setenv("SERB_NUMA_BIND", "0")
    assert multi_gpu.bind_host_to_gpu(0) is None
    monkeypatch.setenv("SERB_NUMA_BIND", "1")
    report = multi_gpu.bind_host_to_gpu(0)
    assert isinstance(report, dict) and "bound" in report and report["cpus"] >= 1
    if not report["bound"]:
        assert os.sched_getaffinity(0) == before
    else:                                     # a GPU box: undo for the tests that follow
        os.sched_setaffinity(0, report["restore"])
