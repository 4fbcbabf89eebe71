"""Temporal pooling operators over encoded frame sequences.

Mirrors ser/_internal/pool/windowing.py:10-71 (`temporal_pooling_windows`, pure index/time
arithmetic, kept on the host) and ser/_internal/pool/stats_pool.py:15-43 (`mean_std_pool`), whose
reduction runs in the library's segmented pooling kernel.  ``overlap_frame_mask`` selects, for
monotone frame timestamps, a contiguous run of frames, so each window becomes a [lo, hi) range.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np
from numpy.typing import NDArray

from . import _native
from .backend import EncodedSequence, PoolingWindow, overlap_frame_mask


def _positive_finite(value: float, name: str) -> float:
    value = float(value)
    if not (value > 0.0 and np.isfinite(value)):
        raise ValueError(f"{name} must be a positive finite float.")
    return value


def temporal_pooling_windows(encoded: EncodedSequence, *, window_size_seconds: float,
                             window_stride_seconds: float) -> list[PoolingWindow]:
    """Pooling windows over the encoded timeline, same values as the reference's generator
    (ser/_internal/pool/windowing.py:10-71) computed in closed form.

    The reference walks a cursor (``cursor += stride``) while a whole window still fits; here the
    window starts are one sequential ``np.add.accumulate`` over ``[t0, stride, stride, ...]`` (the
    same left-to-right float64 sums, hence bit-identical starts), cut where ``start + size`` leaves
    the timeline by more than 1e-9, plus the reference's single tail rule: if the last regular
    window ends short of the timeline, one window of the same size is anchored at its end."""
    size = _positive_finite(window_size_seconds, "window_size_seconds")
    stride = _positive_finite(window_stride_seconds, "window_stride_seconds")
    t0 = float(encoded.frame_start_seconds[0])
    t1 = float(encoded.frame_end_seconds[-1])
    if not t1 - t0 > 0.0:
        raise ValueError("Encoded sequence duration must be positive.")
    span = min(size, t1 - t0)
    if np.isclose(span, t1 - t0):                      # one window spans the whole recording
        return [PoolingWindow(start_seconds=t0, end_seconds=t1)]
    slack = 1e-9
    upper = int(max(0.0, (t1 - t0 - span + slack) / stride)) + 2      # no more starts than this fit
    starts = np.add.accumulate(np.concatenate(([t0], np.full(upper, stride, dtype=np.float64))))
    starts = starts[: int(np.count_nonzero(np.logical_and.accumulate(starts + span <= t1 + slack)))]
    ends = np.minimum(t1, starts + span)
    anchored = (max(t0, t1 - span), t1)
    pairs = list(zip(starts.tolist(), ends.tolist()))
    if not pairs:
        pairs = [anchored]
    elif pairs[-1][1] < t1 - slack and not (np.isclose(pairs[-1][0], anchored[0]) and np.isclose(pairs[-1][1], anchored[1])):
        pairs.append(anchored)
    return [PoolingWindow(start_seconds=a, end_seconds=b) for a, b in pairs]


def frame_ranges(encoded: EncodedSequence, windows: Sequence[PoolingWindow]) -> tuple[NDArray[np.int32], NDArray[np.int32]]:
    """[lo, hi) frame range of every window; same errors as ``overlap_frame_mask``."""
    lo = np.empty(len(windows), dtype=np.int32)
    hi = np.empty(len(windows), dtype=np.int32)
    for i, window in enumerate(windows):
        mask = overlap_frame_mask(encoded, window)          # raises the reference's ValueErrors
        idx = np.flatnonzero(mask)
        if idx[-1] - idx[0] + 1 != idx.size:
            raise ValueError("Pooling window selects a non-contiguous frame set.")
        lo[i], hi[i] = idx[0], idx[-1] + 1
    return lo, hi


def mean_std_pool(encoded: EncodedSequence, windows: Sequence[PoolingWindow], *, device: int = 0) -> NDArray[np.float64]:
    """Mean + std (ddof 0) of the frames overlapping each window -> (len(windows), 2 * dim)."""
    dim = int(encoded.embeddings.shape[1])
    if not windows:
        return np.empty((0, dim * 2), dtype=np.float64)
    lo, hi = frame_ranges(encoded, windows)
    return _native.get_context(device).pool_frames_host(encoded.embeddings, lo, hi, 1)


def mean_pool(encoded: EncodedSequence, windows: Sequence[PoolingWindow], *, device: int = 0) -> NDArray[np.float64]:
    """`embeddings[mask].mean(axis=0)` per window, as HandcraftedBackend.pool computes it."""
    dim = int(encoded.embeddings.shape[1])
    if not windows:
        return np.empty((0, dim), dtype=np.float64)
    lo, hi = frame_ranges(encoded, windows)
    return _native.get_context(device).pool_frames_host(encoded.embeddings, lo, hi, 2)
