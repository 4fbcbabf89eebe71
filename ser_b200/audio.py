"""Audio file reading for the file-level seams (mirror of ser/_internal/utils/audio_utils.py:28-113).

The reference decodes with librosa/soundfile; this image has neither, so PCM WAV is decoded
with the standard library and anything else is rejected.  Post-decode preparation is the
reference's: NaN/Inf -> 0, channel mean, whole-file peak normalisation to [-1, 1].
"""

from __future__ import annotations

import wave
from pathlib import Path

import numpy as np
from numpy.typing import NDArray

_GIT_LFS_POINTER_PREFIX = b"version https://git-lfs.github.com/spec/v1"


class AudioIntegrityError(OSError):
    """A path holds metadata (e.g. a Git LFS pointer) in place of audio bytes."""


class AudioDecodeError(OSError):
    """A regular media file could not be decoded locally."""


def prepare_audio_buffer(raw_audio: NDArray) -> NDArray[np.float32]:
    """audio_utils.py:53-60: float32, NaN/Inf -> 0, mono mix, peak normalise."""
    prepared = np.asarray(raw_audio, dtype=np.float32)
    prepared = np.nan_to_num(prepared, copy=True, nan=0.0, posinf=0.0, neginf=0.0)
    if prepared.ndim == 2:
        if prepared.shape[1] == 0:
            prepared = np.array([], dtype=np.float32)
        else:
            prepared = np.asarray(np.mean(prepared, axis=1, dtype=np.float32), dtype=np.float32)
    elif prepared.ndim != 1:
        raise OSError(f"Unsupported audio shape: {prepared.shape}")
    if prepared.size == 0:
        raise OSError("Audio file contains no samples.")
    peak = float(np.max(np.abs(prepared)))
    if peak == 0:
        return np.zeros_like(prepared)
    return prepared / peak


def decode_wav(path: str | Path) -> tuple[NDArray[np.float32], int]:
    """PCM WAV -> float32 in soundfile's convention (int16 / 32768), shape (frames[, channels])."""
    try:
        with wave.open(str(path), "rb") as handle:
            sample_rate = handle.getframerate()
            channels = handle.getnchannels()
            width = handle.getsampwidth()
            raw = handle.readframes(handle.getnframes())
    except (wave.Error, EOFError) as err:
        raise AudioDecodeError(f"Could not decode audio file {path}: {err}") from err
    if width == 2:
        data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / np.float32(32768.0)
    elif width == 1:
        data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / np.float32(128.0)
    elif width == 4:
        data = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    else:
        raise AudioDecodeError(f"Unsupported PCM sample width {width} in {path}")
    if channels > 1:
        data = data.reshape(-1, channels)
    return data, int(sample_rate)


def read_audio_file(
    file_path: str,
    *,
    start_seconds: float | None = None,
    duration_seconds: float | None = None,
) -> tuple[NDArray[np.float32], int]:
    """Reads (a segment of) an audio file and normalises it to [-1, 1] (audio_utils.py:63-113)."""
    if start_seconds is not None and start_seconds < 0.0:
        raise ValueError("start_seconds must be >= 0")
    if duration_seconds is not None and duration_seconds <= 0.0:
        raise ValueError("duration_seconds must be > 0")
    path = Path(file_path)
    if not path.exists():
        raise FileNotFoundError(f"Audio file not found: {file_path}")
    if not path.is_file():
        raise OSError(f"Path is not a regular file: {file_path}")
    with path.open("rb") as handle:
        if handle.read(len(_GIT_LFS_POINTER_PREFIX)) == _GIT_LFS_POINTER_PREFIX:
            raise AudioIntegrityError(f"Audio file is an unmaterialized Git LFS pointer: {file_path}.")
    data, sample_rate = decode_wav(path)
    first = int(float(start_seconds or 0.0) * sample_rate)
    if duration_seconds is not None:
        data = data[first : first + int(float(duration_seconds) * sample_rate)]
    elif first:
        data = data[first:]
    return prepare_audio_buffer(data), sample_rate
