"""File-level feature extraction (boundary B3; mirror of
ser/_internal/features/feature_extractor.py:26-179): read a file, run the GPU backend, wrap
the rows with their timestamps.  ``settings`` may be the reference's AppConfig or anything
exposing ``feature_flags``; ``None`` means default flags (all groups on).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from numpy.typing import NDArray

from . import dsp
from .audio import read_audio_file, read_pcm16_file
from .config import FeatureFlags
from .handcrafted import HandcraftedBackend


@dataclass(frozen=True)
class FeatureFrame:
    """One window's feature vector with explicit time bounds (feature_extractor.py:26-32)."""

    start_seconds: float
    end_seconds: float
    features: NDArray[np.float64]


def _flags_of(settings) -> FeatureFlags:
    flags = getattr(settings, "feature_flags", None) if settings is not None else None
    return flags if flags is not None else FeatureFlags()


def extract_feature_from_signal(audio: NDArray[np.float32], sample_rate: int, *, settings=None) -> NDArray[np.float64]:
    """Compatibility wrapper (feature_extractor.py:106-118)."""
    return dsp.extract_feature_from_signal(audio, sample_rate, feature_flags=_flags_of(settings))


def extract_feature(file: str, *, settings=None, device: int = 0) -> NDArray[np.float64]:
    """Whole-file feature vector (feature_extractor.py:121-136)."""
    backend = HandcraftedBackend(feature_flags=_flags_of(settings), device=device)
    raw = read_pcm16_file(file)
    if raw is not None:                 # 16-bit PCM: decode scaling / mono / peak normalisation on the device
        return backend.extract_vector_pcm16(*raw)
    audio, sample_rate = read_audio_file(file)
    return backend.extract_vector(audio, sample_rate)


def extract_feature_frames(audiofile: str, frame_size: int = 3, frame_stride: int = 1, *,
                           settings=None, device: int = 0) -> list[FeatureFrame]:
    """Sliding-window features with start/end timestamps (feature_extractor.py:164-179)."""
    if frame_size <= 0:
        raise ValueError("frame_size must be greater than zero.")
    if frame_stride <= 0:
        raise ValueError("frame_stride must be greater than zero.")
    backend = HandcraftedBackend(frame_size_seconds=frame_size, frame_stride_seconds=frame_stride,
                                 feature_flags=_flags_of(settings), device=device)
    raw = read_pcm16_file(audiofile)
    if raw is not None:                 # 16-bit PCM: the device prepares the samples (N1)
        encoded = backend.encode_sequence_pcm16(*raw)
    else:
        audio, sample_rate = read_audio_file(audiofile)
        encoded = backend.encode_sequence(audio, sample_rate)
    return [
        FeatureFrame(
            start_seconds=float(encoded.frame_start_seconds[i]),
            end_seconds=float(encoded.frame_end_seconds[i]),
            features=np.asarray(encoded.embeddings[i], dtype=np.float64),
        )
        for i in range(encoded.embeddings.shape[0])
    ]


def extended_extract_feature(audiofile: str, frame_size: int = 3, frame_stride: int = 1, *,
                             settings=None, device: int = 0) -> list[NDArray[np.float64]]:
    """Frame-wise vectors without timestamps (feature_extractor.py:139-161)."""
    return [frame.features for frame in
            extract_feature_frames(audiofile, frame_size, frame_stride, settings=settings, device=device)]
