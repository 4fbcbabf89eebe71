"""Audio file reading for the file-level seams (mirror of ser/_internal/utils/audio_utils.py:28-113).

The reference decodes with librosa/soundfile; this image has neither, so WAV (integer PCM of 8 / 16 /
24 / 32 bits, IEEE float of 32 / 64 bits, plain or WAVE_FORMAT_EXTENSIBLE headers) is decoded here
and anything else is rejected with ``AudioDecodeError``.  Post-decode preparation is the
reference's: NaN/Inf -> 0, channel mean, whole-file peak normalisation to [-1, 1].

16-bit PCM WAV -- what RAVDESS and the reference's own synthetic generator ship -- does not need
that host pass at all: ``read_pcm16_file`` hands the raw int16 samples to the device entries
(``serb_features_host_pcm16``), which do the same preparation bit-identically on the GPU
(SURVEY.md section 8f, row N1).  ``prepare_audio_buffer`` remains for float / 8 / 24 / 32-bit input.
"""

from __future__ import annotations

import struct
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


def _pcm_to_float32(raw: bytes, width: int, path) -> NDArray[np.float32]:
    """Integer PCM -> float32 in soundfile's convention: ``x / 2^(bits - 1)`` (8-bit WAV is unsigned)."""
    if width == 2:
        return np.frombuffer(raw, dtype="<i2").astype(np.float32) / np.float32(32768.0)
    if width == 1:
        return (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / np.float32(128.0)
    if width == 4:
        return (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    if width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        return (np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float64) / 8388608.0).astype(np.float32)
    raise AudioDecodeError(f"Unsupported PCM sample width {width} in {path}")


_WAVE_FORMAT_PCM, _WAVE_FORMAT_IEEE_FLOAT, _WAVE_FORMAT_EXTENSIBLE = 0x0001, 0x0003, 0xFFFE


def _decode_riff(path) -> tuple[NDArray[np.float32], int, int]:
    """RIFF/WAVE chunk walk for what the standard library's ``wave`` refuses: IEEE-float samples
    (format tag 3, 32 or 64 bit: handed on as float32 like soundfile's ``dtype="float32"`` read) and
    WAVE_FORMAT_EXTENSIBLE headers whose sub-format is PCM or IEEE float.  Returns
    ``(flat samples, channels, sample_rate)``."""
    blob = Path(path).read_bytes()
    if len(blob) < 12 or blob[:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise AudioDecodeError(f"Could not decode audio file {path}: not a RIFF/WAVE file")
    fmt = data = None
    pos = 12
    while pos + 8 <= len(blob):
        tag = blob[pos:pos + 4]
        size = struct.unpack_from("<I", blob, pos + 4)[0]
        body = blob[pos + 8:pos + 8 + size]
        if tag == b"fmt " and fmt is None:
            fmt = body
        elif tag == b"data" and data is None:
            data = body                  # a truncated file yields the samples that are there
        pos += 8 + size + (size & 1)     # chunks are word aligned
    if fmt is None or data is None or len(fmt) < 16:
        raise AudioDecodeError(f"Could not decode audio file {path}: missing fmt or data chunk")
    kind, channels, sample_rate, _rate, block_align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if kind == _WAVE_FORMAT_EXTENSIBLE:
        if len(fmt) < 26:
            raise AudioDecodeError(f"Could not decode audio file {path}: short extensible header")
        kind = struct.unpack_from("<H", fmt, 24)[0]      # first two bytes of the sub-format GUID
    if channels < 1 or sample_rate < 1 or bits % 8 or block_align != channels * (bits // 8):
        raise AudioDecodeError(f"Could not decode audio file {path}: inconsistent fmt chunk")
    width = bits // 8
    data = data[: len(data) // block_align * block_align]
    if kind == _WAVE_FORMAT_IEEE_FLOAT and width in (4, 8):
        samples = np.frombuffer(data, dtype="<f4" if width == 4 else "<f8").astype(np.float32)
    elif kind == _WAVE_FORMAT_PCM:
        samples = _pcm_to_float32(data, width, path)
    else:
        raise AudioDecodeError(f"Could not decode audio file {path}: unsupported WAVE format tag {kind:#06x}")
    return samples, int(channels), int(sample_rate)


def decode_wav(path: str | Path) -> tuple[NDArray[np.float32], int]:
    """WAV -> float32 in soundfile's convention (int16 / 32768, floats as they are), shape
    (frames[, channels]).  Integer PCM goes through the standard library's reader; IEEE-float and
    extensible files, which it refuses, through ``_decode_riff``."""
    try:
        with wave.open(str(path), "rb") as handle:
            sample_rate = handle.getframerate()
            channels = handle.getnchannels()
            width = handle.getsampwidth()
            raw = handle.readframes(handle.getnframes())
        data = _pcm_to_float32(raw[: len(raw) // (width * channels) * (width * channels)], width, path)
    except (wave.Error, EOFError, struct.error):
        try:
            data, channels, sample_rate = _decode_riff(path)
        except (struct.error, ValueError) as err:
            raise AudioDecodeError(f"Could not decode audio file {path}: {err}") from err
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
