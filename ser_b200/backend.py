"""Typed contracts of the representation layer (mirror of ser/_internal/repr/backend.py:19-155):
PoolingWindow, EncodedSequence, overlap_frame_mask and the FeatureBackend protocols.
Validation rules and error texts follow the reference so its own tests read the same here.
"""

from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass
from typing import Protocol, runtime_checkable

import numpy as np
from numpy.typing import NDArray


def _reject(rules) -> None:
    """Raises ``ValueError`` with the text of the first rule whose predicate holds.  Rules are
    evaluated lazily and in order: a later predicate may rely on the earlier ones having passed
    (the order, and therefore which text a doubly-invalid value gets, is the reference's)."""
    for violated, text in rules:
        if violated():
            raise ValueError(text)


@dataclass(frozen=True)
class PoolingWindow:
    """Closed-open time range, in seconds, pooled into one row (repr/backend.py:19-33)."""

    start_seconds: float
    end_seconds: float

    def __post_init__(self) -> None:
        lo, hi = self.start_seconds, self.end_seconds
        _reject((
            (lambda: not (np.isfinite(lo) and np.isfinite(hi)), "PoolingWindow bounds must be finite numbers."),
            (lambda: lo < 0.0, "PoolingWindow start_seconds must be non-negative."),
            (lambda: hi <= lo, "PoolingWindow end_seconds must be greater than start_seconds."),
        ))


@dataclass(frozen=True)
class EncodedSequence:
    """Per-window feature rows with their time bounds (repr/backend.py:36-78)."""

    embeddings: NDArray[np.float32]
    frame_start_seconds: NDArray[np.float64]
    frame_end_seconds: NDArray[np.float64]
    backend_id: str

    def __post_init__(self) -> None:
        rows, t0, t1 = self.embeddings, self.frame_start_seconds, self.frame_end_seconds
        stamps = (("frame_start_seconds", t0), ("frame_end_seconds", t1))
        _reject((
            (lambda: not self.backend_id, "EncodedSequence backend_id must be a non-empty string."),
            (lambda: rows.ndim != 2, "EncodedSequence embeddings must be 2D (frames, features)."),
            (lambda: t0.ndim != 1 or t1.ndim != 1, "Frame timestamp arrays must be 1D."),
            (lambda: rows.shape[0] <= 0, "EncodedSequence must contain at least one frame."),
            *((lambda v=v: v.size != rows.shape[0], f"{name} length must match embeddings frame count.")
              for name, v in stamps),
            *((lambda v=v: not np.isfinite(v).all(), f"EncodedSequence {name} contain non-finite values.")
              for name, v in (("embeddings", rows), *stamps)),
            *((lambda v=v: bool((v[1:] < v[:-1]).any()), f"{name} must be non-decreasing.") for name, v in stamps),
            (lambda: bool((t1 <= t0).any()), "Each frame must satisfy end_seconds > start_seconds."),
        ))


def overlap_frame_mask(encoded: EncodedSequence, window: PoolingWindow) -> NDArray[np.bool_]:
    """Frames whose interval intersects ``window`` (ser/_internal/repr/backend.py:81-111)."""
    first_start = float(encoded.frame_start_seconds[0])
    last_end = float(encoded.frame_end_seconds[-1])
    if window.start_seconds < first_start or window.end_seconds > last_end:
        raise ValueError(
            "Pooling window is outside encoded sequence range: "
            f"[{window.start_seconds}, {window.end_seconds}] vs [{first_start}, {last_end}]"
        )
    mask = (encoded.frame_end_seconds > window.start_seconds) & (
        encoded.frame_start_seconds < window.end_seconds
    )
    if not np.any(mask):
        raise ValueError(
            "Pooling window does not overlap any encoded frames: "
            f"[{window.start_seconds}, {window.end_seconds}]"
        )
    return mask


@runtime_checkable
class FeatureBackend(Protocol):
    """Sequence encoding + temporal pooling (ser/_internal/repr/backend.py:114-143)."""

    @property
    def backend_id(self) -> str: ...

    @property
    def feature_dim(self) -> int: ...

    def encode_sequence(self, audio: NDArray[np.float32], sample_rate: int) -> EncodedSequence: ...

    def pool(self, encoded: EncodedSequence, windows: Sequence[PoolingWindow]) -> NDArray[np.float64]: ...


@runtime_checkable
class VectorFeatureBackend(FeatureBackend, Protocol):
    """Adds whole-clip vector extraction (ser/_internal/repr/backend.py:146-155)."""

    def extract_vector(self, audio: NDArray[np.float32], sample_rate: int) -> NDArray[np.float64]: ...
