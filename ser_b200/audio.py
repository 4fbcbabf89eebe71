"""Audio file reading for the file-level seams (mirror of ser/_internal/utils/audio_utils.py:28-113).

The reference decodes with librosa/soundfile; this image has neither, so PCM WAV is decoded
with the standard library and anything else is rejected.  Post-decode preparation is the
reference's: NaN/Inf -> 0, channel mean, whole-file peak normalisation to [-1, 1].

16-bit PCM WAV -- what RAVDESS and the reference's own synthetic generator ship -- does not need
that host pass at all: ``read_pcm16_file`` hands the raw int16 samples to the device entries
(``serb_features_host_pcm16``), which do the same preparation bit-identically on the GPU
(SURVEY.md section 8f, row N1).  ``prepare_audio_buffer`` remains for float / 8 / 24 / 32-bit input.
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
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        data = (np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float64) / 8388608.0).astype(np.float32)
    else:
        raise AudioDecodeError(f"Unsupported PCM sample width {width} in {path}")
    if channels > 1:
        data = data.reshape(-1, channels)
    return data, int(sample_rate)


def _check_path(file_path: str) -> Path:
    path = Path(file_path)
    if not path.exists():
        raise FileNotFoundError(f"Audio file not found: {file_path}")
    if not path.is_file():
        raise OSError(f"Path is not a regular file: {file_path}")
    with path.open("rb") as handle:
        if handle.read(len(_GIT_LFS_POINTER_PREFIX)) == _GIT_LFS_POINTER_PREFIX:
            raise AudioIntegrityError(f"Audio file is an unmaterialized Git LFS pointer: {file_path}.")
    return path


def read_pcm16_file(
    file_path: str,
    *,
    start_seconds: float | None = None,
    duration_seconds: float | None = None,
) -> tuple[NDArray[np.int16], int, int] | None:
    """Raw samples of a 16-bit PCM WAV file (or segment) for the device-side preparation path
    (``serb_features_host_pcm16``): ``(interleaved int16, channels, sample_rate)``, or ``None``
    when the file is not 16-bit PCM WAV (the caller then takes ``read_audio_file``).  Segment
    bounds follow librosa.load(offset=, duration=): ``int(offset * sr)`` frames skipped,
    ``int(duration * sr)`` frames read (audio_utils.py:104-109).  Same argument and path errors as
    ``read_audio_file``."""
    if start_seconds is not None and start_seconds < 0.0:
        raise ValueError("start_seconds must be >= 0")
    if duration_seconds is not None and duration_seconds <= 0.0:
        raise ValueError("duration_seconds must be > 0")
    path = _check_path(file_path)
    try:
        with wave.open(str(path), "rb") as handle:
            if handle.getsampwidth() != 2 or handle.getcomptype() != "NONE":
                return None
            sample_rate = handle.getframerate()
            channels = handle.getnchannels()
            total = handle.getnframes()
            first = min(int(float(start_seconds or 0.0) * sample_rate), total)
            count = total - first
            if duration_seconds is not None:
                count = min(count, int(float(duration_seconds) * sample_rate))
            handle.setpos(first)
            raw = handle.readframes(max(count, 0))
    except (wave.Error, EOFError):
        return None
    pcm = np.frombuffer(raw, dtype="<i2")
    if channels < 1 or channels > 256 or pcm.size < channels:
        if pcm.size == 0:
            raise OSError("Audio file contains no samples.")
        return None
    return pcm[: pcm.size // channels * channels], int(channels), int(sample_rate)


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
    path = _check_path(file_path)
    data, sample_rate = decode_wav(path)
    first = int(float(start_seconds or 0.0) * sample_rate)
    if duration_seconds is not None:
        data = data[first : first + int(float(duration_seconds) * sample_rate)]
    elif first:
        data = data[first:]
    return prepare_audio_buffer(data), sample_rate
