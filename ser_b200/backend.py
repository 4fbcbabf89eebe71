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


@dataclass(frozen=True)
class PoolingWindow:
    """Closed-open time range, in seconds, pooled into one row."""

    start_seconds: float
    end_seconds: float

    def __post_init__(self) -> None:
        if not np.isfinite(self.start_seconds) or not np.isfinite(self.end_seconds):
            raise ValueError("PoolingWindow bounds must be finite numbers.")
        if self.start_seconds < 0.0:
            raise ValueError("PoolingWindow start_seconds must be non-negative.")
        if self.end_seconds <= self.start_seconds:
            raise ValueError("PoolingWindow end_seconds must be greater than start_seconds.")


@dataclass(frozen=True)
class EncodedSequence:
    """Per-window feature rows with their time bounds."""

    embeddings: NDArray[np.float32]
    frame_start_seconds: NDArray[np.float64]
    frame_end_seconds: NDArray[np.float64]
    backend_id: str

    def __post_init__(self) -> None:
        if not self.backend_id:
            raise ValueError("EncodedSequence backend_id must be a non-empty string.")
        if self.embeddings.ndim != 2:
            raise ValueError("EncodedSequence embeddings must be 2D (frames, features).")
        if self.frame_start_seconds.ndim != 1 or self.frame_end_seconds.ndim != 1:
            raise ValueError("Frame timestamp arrays must be 1D.")
        n_frames = int(self.embeddings.shape[0])
        if n_frames <= 0:
            raise ValueError("EncodedSequence must contain at least one frame.")
        if self.frame_start_seconds.size != n_frames:
            raise ValueError("frame_start_seconds length must match embeddings frame count.")
        if self.frame_end_seconds.size != n_frames:
            raise ValueError("frame_end_seconds length must match embeddings frame count.")
        if not np.all(np.isfinite(self.embeddings)):
            raise ValueError("EncodedSequence embeddings contain non-finite values.")
        if not np.all(np.isfinite(self.frame_start_seconds)):
            raise ValueError("EncodedSequence frame_start_seconds contain non-finite values.")
        if not np.all(np.isfinite(self.frame_end_seconds)):
            raise ValueError("EncodedSequence frame_end_seconds contain non-finite values.")
        if np.any(np.diff(self.frame_start_seconds) < 0.0):
            raise ValueError("frame_start_seconds must be non-decreasing.")
        if np.any(np.diff(self.frame_end_seconds) < 0.0):
            raise ValueError("frame_end_seconds must be non-decreasing.")
        if np.any(self.frame_end_seconds <= self.frame_start_seconds):
            raise ValueError("Each frame must satisfy end_seconds > start_seconds.")


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
