"""Boundary B4's error taxonomy on a machine without a GPU (ser_b200/fast_inference.py; the reference's
mapping is ser/_internal/runtime/fast_public_boundary.py:181-186,404-411): model problems before any
audio is touched, argument errors passing through, and the missing device arriving as the runtime
boundary's execution error -- never as a silently computed CPU result."""

from __future__ import annotations

import wave

import numpy as np
import pytest

from ser_b200 import _native, fast_inference, feature_extractor
from ser_b200.schema import InferenceRequest


def _wav(path, n=20000, sr=16000):
    t = np.arange(n) / sr
    pcm = (0.3 * np.sin(2 * np.pi * 220.0 * t) * 32767).astype("<i2")
    with wave.open(str(path), "wb") as handle:
        handle.setnchannels(1)
        handle.setsampwidth(2)
        handle.setframerate(sr)
        handle.writeframes(pcm.tobytes())
    return str(path)


def test_incompatible_artifacts_are_refused_before_any_work(tmp_path):
    request = InferenceRequest(file_path=str(tmp_path / "never_opened.wav"))
    for metadata in ({"backend_id": "hf_whisper", "profile": "fast"}, {"backend_id": "handcrafted", "profile": "accurate"}):
        loaded = fast_inference.LoadedModel(model=object(), expected_feature_size=193, artifact_metadata=metadata)
        with pytest.raises(fast_inference.FastModelUnavailableError, match="No compatible fast-profile model artifact"):
            fast_inference.run_fast_inference(request, None, loaded_model=loaded)
    assert issubclass(fast_inference.FastModelUnavailableError, FileNotFoundError)


def test_missing_model_without_the_reference_loader(tmp_path, monkeypatch):
    import sys

    for name in [m for m in sys.modules if m == "ser" or m.startswith("ser.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.setattr(sys, "path", [p for p in sys.path if "reference" not in p and "baseline" not in p])
    with pytest.raises(fast_inference.FastModelUnavailableError, match="loaded_model was not given"):
        fast_inference.run_fast_inference(InferenceRequest(file_path=str(tmp_path / "x.wav")), None)


def test_path_and_argument_errors_pass_through(tmp_path):
    loaded = fast_inference.LoadedModel(model=object(), expected_feature_size=193)
    with pytest.raises(FileNotFoundError, match="Audio file not found"):
        fast_inference.run_fast_inference(InferenceRequest(file_path=str(tmp_path / "absent.wav")), None, loaded_model=loaded)
    path = _wav(tmp_path / "a.wav")
    with pytest.raises(ValueError, match="frame_size must be greater than zero."):
        feature_extractor.extract_feature_frames(path, frame_size=0)
    with pytest.raises(ValueError, match="frame_stride must be greater than zero."):
        feature_extractor.extract_feature_frames(path, frame_stride=0)


def test_missing_device_becomes_the_execution_error(tmp_path):
    if _native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    loaded = fast_inference.LoadedModel(model=object(), expected_feature_size=193)
    with pytest.raises(fast_inference.FastInferenceExecutionError, match="no CPU fallback"):
        fast_inference.run_fast_inference(InferenceRequest(file_path=_wav(tmp_path / "b.wav")), None, loaded_model=loaded)
    assert issubclass(fast_inference.FastInferenceExecutionError, RuntimeError)
