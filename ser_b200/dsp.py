"""GPU feature extraction behind the reference's DSP seam (boundary B1).

``extract_feature_from_signal(audio, sample_rate, *, feature_flags=None)`` keeps the signature,
argument meaning, output (float64 vector in the order mfcc | chroma | mel | contrast |
tonnetz) and error behaviour of ser/_internal/utils/dsp.py:67-151; the arithmetic runs in
libser_b200's CUDA kernels.  The batched forms underneath (`extract_features_ragged`,
`extract_features_batch`) are what the sliding-window and training callers use: one native
call for any number of clips.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np
from numpy.typing import NDArray

from . import _native
from .config import FeatureFlags, feature_dim, flag_bits

ParameterError = _native.ParameterError


def _validate_signal(audio: NDArray, sample_rate: int) -> None:
    # same checks, same order, same texts as dsp.py:85-95
    if sample_rate <= 0:
        raise ValueError("Sample rate must be a positive integer.")
    if audio.ndim != 1:
        raise ValueError("Audio must be mono (1D array).")
    if audio.size == 0:
        raise ValueError("Audio contains no samples.")


def extract_feature_from_signal(
    audio: NDArray[np.float32],
    sample_rate: int,
    *,
    feature_flags: FeatureFlags | None = None,
    device: int = 0,
) -> NDArray[np.float64]:
    """One feature vector for one in-memory mono clip.

    Raises:
        ValueError: invalid sample rate / shape / empty or non-finite audio (reference texts).
        ParameterError: a librosa parameter check the reference would trip (e.g. contrast
            enabled at sample_rate <= 12800).
        RuntimeError: CUDA failure, missing library or no GPU.
    """
    audio = np.asarray(audio)
    _validate_signal(audio, sample_rate)
    flags = feature_flags if feature_flags is not None else FeatureFlags()
    prepared = np.ascontiguousarray(audio, dtype=np.float32)
    if not bool(np.all(np.isfinite(prepared))):
        raise ValueError("Audio buffer is not finite everywhere.")
    if feature_dim(flags) == 0:
        return np.empty(0, dtype=np.float64)
    ctx = _native.get_context(device)
    rows = ctx.features_host(
        prepared,
        np.zeros(1, dtype=np.int64),
        np.asarray([prepared.size], dtype=np.int64),
        int(sample_rate),
        flag_bits(flags),
    )
    return rows[0].astype(np.float64)


def extract_features_ragged(
    wave: NDArray[np.float32],
    starts: NDArray[np.int64],
    lengths: NDArray[np.int64],
    sample_rate: int,
    *,
    feature_flags: FeatureFlags | None = None,
    device: int = 0,
) -> NDArray[np.float32]:
    """Feature rows for clips ``wave[starts[i] : starts[i] + lengths[i]]`` (clips may overlap).

    Returns float32 of shape (n_clips, dim): the precision the reference's inference path
    keeps (ser/_internal/repr/handcrafted.py:95,103).
    """
    wave = np.asarray(wave)
    if sample_rate <= 0:
        raise ValueError("Sample rate must be a positive integer.")
    if wave.ndim != 1:
        raise ValueError("Audio must be mono (1D array).")
    flags = feature_flags if feature_flags is not None else FeatureFlags()
    ctx = _native.get_context(device)
    return ctx.features_host(wave, starts, lengths, int(sample_rate), flag_bits(flags))


def extract_features_batch(
    clips: Sequence[NDArray[np.float32]],
    sample_rate: int,
    *,
    feature_flags: FeatureFlags | None = None,
    device: int = 0,
) -> NDArray[np.float64]:
    """Whole-clip vectors for a list of clips (the training loop's unit of work,
    ser/_internal/data/data_loader.py:485-529), float64 like ``extract_vector``."""
    if sample_rate <= 0:
        raise ValueError("Sample rate must be a positive integer.")
    arrays = []
    for clip in clips:
        clip = np.asarray(clip)
        _validate_signal(clip, sample_rate)
        arrays.append(np.ascontiguousarray(clip, dtype=np.float32))
    flags = feature_flags if feature_flags is not None else FeatureFlags()
    if not arrays:
        return np.empty((0, feature_dim(flags)), dtype=np.float64)
    # every clip stays in its own array: the library copies them to aligned offsets piecewise
    rows = _native.get_context(device).features_host_clips(arrays, sample_rate, flag_bits(flags))
    return rows.astype(np.float64)


def extract_features_pcm16(
    files: Sequence[NDArray[np.int16]],
    channels: Sequence[int] | int,
    clip_file: NDArray[np.int64],
    clip_starts: NDArray[np.int64],
    clip_lengths: NDArray[np.int64],
    sample_rate: int,
    *,
    feature_flags: FeatureFlags | None = None,
    device: int = 0,
) -> NDArray[np.float32]:
    """Feature rows for clips cut out of raw 16-bit PCM files.

    The device performs ``read_audio_file``'s post-decode preparation
    (ser/_internal/utils/audio_utils.py:28-60: x / 32768, channel mean, whole-file peak
    normalisation) bit-identically, so the host ships 2 bytes per sample and never builds the
    float32 buffer.  ``files[f]`` holds ``frames * channels[f]`` interleaved samples; clip ``i`` is
    frames ``[clip_starts[i], clip_starts[i] + clip_lengths[i])`` of file ``clip_file[i]``
    (non-decreasing).  Returns float32 (n_clips, dim)."""
    if sample_rate <= 0:
        raise ValueError("Sample rate must be a positive integer.")
    flags = feature_flags if feature_flags is not None else FeatureFlags()
    return _native.get_context(device).features_host_pcm16(
        files, channels, clip_file, clip_starts, clip_lengths, int(sample_rate), flag_bits(flags))
