"""Multi-GPU sharding of a clip batch: no data-path collective.

Clips are independent, so the batch is cut into contiguous, sample-count-balanced ranges,
one per device; each device writes its (n_i, dim) block and the host concatenates in the
original order.  Two drivers:

* ``extract_features_sharded`` -- one host thread per device inside one process (the native
  calls release the GIL);
* one process per GPU (``bench.py`` under torchrun) using ``shard_bounds`` for the split.
"""

from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _native
from .config import FeatureFlags, feature_dim, flag_bits


def shard_bounds(lengths: np.ndarray, n_shards: int) -> list[tuple[int, int]]:
    """Contiguous [lo, hi) clip ranges with near-equal total sample counts."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = int(lengths.size)
    if n_shards <= 0:
        raise ValueError("n_shards must be positive")
    if n == 0:
        return [(0, 0)] * n_shards
    cum = np.concatenate(([0], np.cumsum(lengths)))
    total = int(cum[-1])
    bounds = []
    lo = 0
    for shard in range(n_shards):
        if shard == n_shards - 1:
            hi = n
        else:
            target = total * (shard + 1) / n_shards
            hi = int(np.searchsorted(cum, target, side="left"))
            hi = min(max(hi, lo), n)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def extract_features_sharded(wave: np.ndarray, starts: np.ndarray, lengths: np.ndarray, sample_rate: int,
                             *, feature_flags: FeatureFlags | None = None,
                             devices: list[int] | None = None) -> np.ndarray:
    """Ragged-batch features across several GPUs of one box; rows come back in input order."""
    flags = feature_flags if feature_flags is not None else FeatureFlags()
    starts = np.asarray(starts, dtype=np.int64)
    lengths = np.asarray(lengths, dtype=np.int64)
    if devices is None:
        devices = list(range(max(1, _native.device_count())))
    bits = flag_bits(flags)
    out = np.empty((starts.size, feature_dim(flags)), dtype=np.float32)
    bounds = shard_bounds(lengths, len(devices))

    def run(shard: int) -> None:
        lo, hi = bounds[shard]
        if hi <= lo:
            return
        # hand each device only the span of the waveform its clips touch
        first = int(starts[lo:hi].min())
        last = int((starts[lo:hi] + lengths[lo:hi]).max())
        first -= first % 4  # keep 16-byte alignment of clip starts for the TMA path
        ctx = _native.get_context(devices[shard])
        out[lo:hi] = ctx.features_host(wave[first:last], starts[lo:hi] - first, lengths[lo:hi],
                                       sample_rate, bits)

    with ThreadPoolExecutor(max_workers=len(devices)) as pool:
        list(pool.map(run, range(len(devices))))
    return out
