"""Minimal ``soundfile`` stand-in (PCM WAV via stdlib ``wave``) for running the reference's
host code in this container.  TEST INFRASTRUCTURE (oracle); SURVEY.md Appendix C.

Touched by the reference at ser/_internal/utils/audio_utils.py:135-137 and
ser/_internal/models/training_readiness.py:546-551,1540.
"""

from __future__ import annotations

import wave
from dataclasses import dataclass

import numpy as np

__version__ = "0.13.1+oracle"


@dataclass
class _Info:
    samplerate: int
    channels: int
    frames: int
    format: str = "WAV"
    subtype: str = "PCM_16"

    @property
    def duration(self) -> float:
        return self.frames / float(self.samplerate)


def info(path):
    with wave.open(str(path), "rb") as handle:
        return _Info(handle.getframerate(), handle.getnchannels(), handle.getnframes())


def _read_all(path, dtype="float32", always_2d=False):
    with wave.open(str(path), "rb") as handle:
        sr = handle.getframerate()
        channels = handle.getnchannels()
        width = handle.getsampwidth()
        raw = handle.readframes(handle.getnframes())
    if width != 2:
        raise RuntimeError("oracle soundfile stand-in decodes PCM16 only")
    ints = np.frombuffer(raw, dtype="<i2")
    if dtype == "int16":
        data = ints.copy()
    else:
        data = (ints.astype(np.float32) / np.float32(32768.0)).astype(dtype)
    if channels > 1 or always_2d:
        data = data.reshape(-1, channels)
    return data, sr


def read(path, dtype="float64", always_2d=False, **_kwargs):
    return _read_all(path, dtype=dtype, always_2d=always_2d)


class SoundFile:
    def __init__(self, path, mode="r"):
        self._path = path
        meta = info(path)
        self.samplerate = meta.samplerate
        self.channels = meta.channels
        self.frames = meta.frames

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def read(self, frames=-1, dtype="float64", always_2d=False):
        data, _ = _read_all(self._path, dtype=dtype, always_2d=always_2d)
        return data if frames < 0 else data[:frames]

    def blocks(self, blocksize=65536, dtype="float64", always_2d=False):
        data, _ = _read_all(self._path, dtype=dtype, always_2d=always_2d)
        for start in range(0, data.shape[0], blocksize):
            yield data[start : start + blocksize]
