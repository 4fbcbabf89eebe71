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


def temporal_pooling_windows(encoded: EncodedSequence, *, window_size_seconds: float,
                             window_stride_seconds: float) -> list[PoolingWindow]:
    """Ordered pooling windows covering the encoded timeline (windowing.py:10-71)."""
    if window_size_seconds <= 0.0 or not np.isfinite(window_size_seconds):
        raise ValueError("window_size_seconds must be a positive finite float.")
    if window_stride_seconds <= 0.0 or not np.isfinite(window_stride_seconds):
        raise ValueError("window_stride_seconds must be a positive finite float.")
    clip_start = float(encoded.frame_start_seconds[0])
    clip_end = float(encoded.frame_end_seconds[-1])
    clip_duration = clip_end - clip_start
    if clip_duration <= 0.0:
        raise ValueError("Encoded sequence duration must be positive.")
    effective_window = min(window_size_seconds, clip_duration)
    if np.isclose(effective_window, clip_duration):
        return [PoolingWindow(start_seconds=clip_start, end_seconds=clip_end)]
    windows: list[PoolingWindow] = []
    epsilon = 1e-9
    cursor = clip_start
    while cursor + effective_window <= clip_end + epsilon:
        end = min(clip_end, cursor + effective_window)
        windows.append(PoolingWindow(start_seconds=cursor, end_seconds=end))
        cursor += window_stride_seconds
    if not windows:
        return [PoolingWindow(start_seconds=max(clip_start, clip_end - effective_window), end_seconds=clip_end)]
    if windows[-1].end_seconds < clip_end - epsilon:
        tail = PoolingWindow(start_seconds=max(clip_start, clip_end - effective_window), end_seconds=clip_end)
        previous = windows[-1]
        if not (np.isclose(previous.start_seconds, tail.start_seconds)
                and np.isclose(previous.end_seconds, tail.end_seconds)):
            windows.append(tail)
    return windows


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
