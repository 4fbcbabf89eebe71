"""Config c1 end to end on the GPU: ``run_fast_inference`` on a sample.wav-shaped file (16 kHz,
69 937 samples, five 3 s / 1 s windows) with a real scikit-learn Pipeline as the loaded model.
Timestamps follow handcrafted.py:78-97, labels must equal scikit-learn's own predict on the same
feature rows, segments the run-length merge of fast_path.py:99-144, errors the reference's taxonomy
(fast_public_boundary.py:181-186, 404-411)."""

from __future__ import annotations

import time
import wave
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_wav(path, pcm, sr):
    with wave.open(str(path), "wb") as handle:
        handle.setnchannels(1)
        handle.setsampwidth(2)
        handle.setframerate(sr)
        handle.writeframes(np.asarray(pcm, dtype="<i2").tobytes())


@pytest.fixture(scope="module")
def fitted_model():
    from sklearn.neural_network import MLPClassifier
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    from ser_b200 import synth

    rng = np.random.default_rng(11)
    labels = np.asarray(sorted(synth.RAVDESS_EMOTIONS.values()))
    centers = rng.standard_normal((len(labels), 193)) * 2.0
    y = rng.integers(0, len(labels), size=400)
    x = centers[y] + rng.standard_normal((400, 193))
    model = Pipeline([("scaler", StandardScaler()),
                      ("classifier", MLPClassifier(hidden_layer_sizes=(300,), max_iter=40, random_state=1))])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(x, labels[y])
    return model


def test_run_fast_inference_on_a_sample_shaped_file(tmp_path, fitted_model):
    from ser_b200 import fast_inference, fast_path, synth
    from ser_b200.feature_extractor import extract_feature_frames
    from ser_b200.schema import InferenceRequest

    sr, n = 16000, 69937
    pcm = synth.clip_pcm16(synth.ClipSpec(20, 2, 5), sr, n)
    path = tmp_path / "sample.wav"
    _write_wav(path, pcm, sr)
    loaded = fast_inference.LoadedModel(model=fitted_model, expected_feature_size=193,
                                        artifact_metadata={"backend_id": "handcrafted", "profile": "fast"})
    request = InferenceRequest(file_path=str(path), include_transcript=False)
    result = fast_inference.run_fast_inference(request, None, loaded_model=loaded)
    assert [(f.start_seconds, f.end_seconds) for f in result.frames] == \
           [(0.0, 3.0), (1.0, 4.0), (2.0, 4.3710625), (3.0, 4.3710625), (4.0, 4.3710625)]
    frames = extract_feature_frames(str(path))
    rows = np.vstack([f.features for f in frames])
    assert rows.shape == (5, 193)
    expected = fitted_model.predict(rows).tolist()
    assert [f.emotion for f in result.frames] == expected
    proba = fitted_model.predict_proba(rows)
    np.testing.assert_allclose([f.confidence for f in result.frames], proba.max(axis=1), rtol=0, atol=1e-9)
    merged = fast_path.segment_predictions(result.frames)
    assert [(s.emotion, s.start_seconds, s.end_seconds) for s in result.segments] == \
           [(s.emotion, s.start_seconds, s.end_seconds) for s in merged]
    # the reference's harness semantics (benchmarks.py:21-55): wall time of repeated predictions
    times = []
    for _ in range(5):
        t0 = time.perf_counter()
        fast_inference.run_fast_inference(request, None, loaded_model=loaded)
        times.append(time.perf_counter() - t0)
    print(f"c1 latency: mean {np.mean(times) * 1e3:.2f} ms, max {np.max(times) * 1e3:.2f} ms "
          f"(reference publishes mean 1544 ms, p95 2963 ms on CPU)")


def test_error_taxonomy(tmp_path, fitted_model):
    from ser_b200 import fast_inference
    from ser_b200.schema import InferenceRequest

    loaded = fast_inference.LoadedModel(model=fitted_model, expected_feature_size=187)
    path = tmp_path / "x.wav"
    _write_wav(path, np.zeros(20000, dtype=np.int16), 16000)
    with pytest.raises(ValueError, match="Feature vector size mismatch for loaded model."):
        fast_inference.run_fast_inference(InferenceRequest(file_path=str(path)), None, loaded_model=loaded)
    wrong = fast_inference.LoadedModel(model=fitted_model, expected_feature_size=193,
                                       artifact_metadata={"backend_id": "hf_whisper", "profile": "accurate"})
    with pytest.raises(fast_inference.FastModelUnavailableError):
        fast_inference.run_fast_inference(InferenceRequest(file_path=str(path)), None, loaded_model=wrong)
    good = fast_inference.LoadedModel(model=fitted_model, expected_feature_size=193)
    with pytest.raises(FileNotFoundError):
        fast_inference.run_fast_inference(InferenceRequest(file_path=str(tmp_path / "missing.wav")), None, loaded_model=good)
