"""Sliding-window framing on the host (ser_b200/handcrafted.py:frame_bounds, boundary B2): window starts
and ends on the sample grid and their float64 timestamps equal the reference's per-window loop
(ser/_internal/repr/handcrafted.py:78-97) -- restated here as the explicit loop, and, where the reference
checkout is present, run for real with its feature call stubbed out.  Also the argument errors that
must surface before any device work.  CPU only."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

from ser_b200.handcrafted import HandcraftedBackend, frame_bounds

REPO = Path(__file__).resolve().parents[1]
REFERENCE = Path("/root/reference")

CASES = [
    # (n_samples, sample_rate, frame seconds, stride seconds)
    (69937, 16000, 3, 1), (168000, 48000, 3, 1), (57_600_000 // 100, 16000, 3, 1), (1, 16000, 3, 1), (15999, 16000, 3, 1),
    (16000, 16000, 3, 1), (16001, 16000, 3, 1), (48000, 16000, 3, 1), (48001, 16000, 3, 1), (100, 8, 3, 1), (7, 3, 1, 2),
    (22050 * 5 + 17, 22050, 2, 2), (44100 * 3, 44100, 1, 3), (1000, 16000, 0.01, 0.005), (1000, 7, 0.05, 0.05),
    (5000, 11025, 0.3, 0.7), (12345, 48000, 0.25, 0.1),
]


def _loop(n, sr, size, stride):
    length = max(1, int(round(size * sr)))
    step = max(1, int(round(stride * sr)))
    starts, ends = [], []
    for s in range(0, n, step):
        e = min(s + length, n)
        if e - s == 0:
            continue
        starts.append(s)
        ends.append(e)
    return np.asarray(starts, dtype=np.int64), np.asarray(ends, dtype=np.int64)


@pytest.mark.parametrize("n,sr,size,stride", CASES)
def test_frame_bounds_equal_the_window_loop(n, sr, size, stride):
    starts, ends = frame_bounds(n, sr, size, stride)
    want_s, want_e = _loop(n, sr, size, stride)
    assert starts.dtype == np.int64 and ends.dtype == np.int64
    assert np.array_equal(starts, want_s) and np.array_equal(ends, want_e)
    assert np.all(ends > starts)
    if stride <= size:
        assert ends[-1] == n                   # overlapping or abutting windows reach the last sample


def test_sample_wav_shape_gives_the_surveyed_windows():
    starts, ends = frame_bounds(69937, 16000, 3, 1)
    assert (ends - starts).tolist() == [48000, 48000, 37937, 21937, 5937]
    assert (ends.astype(np.float64) / 16000.0).tolist() == [3.0, 4.0, 4.3710625, 4.3710625, 4.3710625]


@pytest.fixture()
def reference_backend(monkeypatch):
    if not (REFERENCE / "ser").exists():
        pytest.skip("reference checkout not present")
    monkeypatch.syspath_prepend(str(REFERENCE))
    monkeypatch.syspath_prepend(str(REPO / "oracle" / "shim"))
    import importlib

    module = importlib.import_module("ser._internal.repr.handcrafted")
    dsp = importlib.import_module("ser._internal.utils.dsp")
    monkeypatch.setattr(dsp, "extract_feature_from_signal", lambda audio, sr, feature_flags=None: np.zeros(193))
    yield module.HandcraftedBackend
    for name in [m for m in sys.modules if m == "ser" or m.startswith("ser.") or m in ("librosa", "soundfile", "colored")
                 or m.startswith("librosa.")]:
        sys.modules.pop(name, None)


def test_timestamps_equal_the_reference_backend(reference_backend):
    for n, sr, size, stride in CASES:
        if n > 600_000:
            continue
        theirs = reference_backend(frame_size_seconds=size, frame_stride_seconds=stride).encode_sequence(
            np.zeros(n, dtype=np.float32), sr)
        starts, ends = frame_bounds(n, sr, size, stride)
        assert np.array_equal(theirs.frame_start_seconds, starts.astype(np.float64) / float(sr)), (n, sr, size, stride)
        assert np.array_equal(theirs.frame_end_seconds, ends.astype(np.float64) / float(sr)), (n, sr, size, stride)


def test_argument_errors_surface_before_any_device_work():
    with pytest.raises(ValueError, match="frame_size_seconds must be greater than zero."):
        HandcraftedBackend(frame_size_seconds=0)
    with pytest.raises(ValueError, match="frame_stride_seconds must be greater than zero."):
        HandcraftedBackend(frame_stride_seconds=-1)
    backend = HandcraftedBackend()
    assert backend.backend_id == "handcrafted" and backend.feature_dim == 193
    assert backend.prepare_runtime() is None
    audio = np.zeros(100, dtype=np.float32)
    with pytest.raises(ValueError, match="sample_rate must be a positive integer."):
        backend.encode_sequence(audio, 0)
    with pytest.raises(ValueError, match=r"audio must be mono \(1D array\)."):
        backend.encode_sequence(np.zeros((2, 50), dtype=np.float32), 16000)
    with pytest.raises(ValueError, match="audio must contain at least one sample."):
        backend.encode_sequence(np.zeros(0, dtype=np.float32), 16000)
    bad = audio.copy()
    bad[3] = np.nan
    with pytest.raises(ValueError, match="Audio buffer is not finite everywhere."):
        backend.encode_sequence(bad, 16000)
    with pytest.raises(ValueError, match="pcm must be a 1-D int16 array of interleaved samples."):
        backend.encode_sequence_pcm16(np.zeros(10, dtype=np.float32), 1, 16000)
    with pytest.raises(ValueError, match="audio must contain at least one sample."):
        backend.encode_sequence_pcm16(np.zeros(1, dtype=np.int16), 2, 16000)
