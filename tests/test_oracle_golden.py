"""Pins oracle/ser_oracle.py (the restated driver) to the REFERENCE's own host code.

tests/golden/fast_profile_golden.npz was produced by /root/reference's dsp.py, handcrafted.py,
fast_path.py and audio_utils.py running over the oracle's librosa restatement
(tests/golden/make_golden.py).  The restated driver must reproduce it bit for bit.
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle import ser_oracle
from ser_b200 import synth


def _audio(golden, name):
    return synth.decode_pcm16(golden[f"{name}/pcm"]), int(golden[f"{name}/sr"])


def test_feature_order_and_dims(golden):
    # 193 = 40 + 12 + 128 + 7 + 6 (tests/suites/unit/repr/test_handcrafted_backend.py:24)
    assert ser_oracle.feature_dim(ser_oracle.FeatureFlags()) == 193
    assert ser_oracle.feature_dim(ser_oracle.FeatureFlags(tonnetz=False)) == 187
    assert ser_oracle.feature_dim(ser_oracle.FeatureFlags(mfcc=False, mel=False)) == 25


@pytest.mark.parametrize("name", ["c16k_3s", "c22k_2s", "c16k_tail_5937", "c16k_2048", "c16k_short_1500",
                                   "c16k_short_1001", "c16k_short_300", "c48k_short_512", "sine16k_1p5s",
                                   "silence16k"])
def test_driver_matches_reference_host_code(golden, name):
    audio, sr = _audio(golden, name)
    full = ser_oracle.extract_feature_from_signal(audio, sr)
    assert full.dtype == np.float64 and full.shape == (193,)
    np.testing.assert_array_equal(full, golden[f"{name}/features"])
    # spectral contrast is identically zero on the reference path
    assert np.all(full[180:187] == 0.0)


def test_flag_subset_is_a_prefix_slice(golden):
    audio, sr = _audio(golden, "c16k_2048")
    slim = ser_oracle.extract_feature_from_signal(audio, sr, feature_flags=ser_oracle.FeatureFlags(tonnetz=False))
    np.testing.assert_array_equal(slim, golden["c16k_2048/features"][:187])


def test_sequence_matches_reference_backend(golden):
    audio = synth.decode_pcm16(golden["seq/pcm"])
    emb, starts, ends = ser_oracle.encode_sequence(audio, int(golden["seq/sr"]))
    np.testing.assert_array_equal(starts, golden["seq/starts"])
    np.testing.assert_array_equal(ends, golden["seq/ends"])
    np.testing.assert_array_equal(emb, golden["seq/embeddings"])
    assert emb.dtype == np.float32
    # config c1 shape: 5 windows of [48000, 48000, 37937, 21937, 5937] samples
    assert [int(round((e - s) * 16000)) for s, e in zip(starts, ends)] == [48000, 48000, 37937, 21937, 5937]


def _weights(golden):
    return ser_oracle.MlpWeights(
        mean=golden["mlp/mean"], scale=golden["mlp/scale"],
        coefs=(golden["mlp/w1"], golden["mlp/w2"]), intercepts=(golden["mlp/b1"], golden["mlp/b2"]),
        classes=tuple(golden["mlp/classes"].tolist()), out_activation="softmax",
    )


def test_mlp_restatement_matches_sklearn(golden):
    weights = _weights(golden)
    proba = ser_oracle.mlp_predict_proba(weights, golden["mlp/x_eval"])
    np.testing.assert_allclose(proba, golden["mlp/proba"], rtol=1e-12, atol=1e-15)
    assert ser_oracle.mlp_predict(weights, golden["mlp/x_eval"]) == golden["mlp/labels"].tolist()


def test_segment_merge_matches_reference_fast_path(golden):
    weights = _weights(golden)
    rows = golden["fast/frame_rows"]
    frames, segments = ser_oracle.predict_frames(
        weights, golden["mlp/x_eval"][rows], golden["fast/frame_starts"], golden["fast/frame_ends"]
    )
    assert [f.emotion for f in frames] == golden["fast/frame_labels"].tolist()
    assert [s.emotion for s in segments] == golden["fast/seg_labels"].tolist()
    np.testing.assert_array_equal([s.start_seconds for s in segments], golden["fast/seg_starts"])
    np.testing.assert_array_equal([s.end_seconds for s in segments], golden["fast/seg_ends"])
    np.testing.assert_allclose([s.confidence for s in segments], golden["fast/seg_conf"], rtol=1e-12)


def test_binary_logistic_restatement_matches_sklearn():
    from oracle import ser_oracle

    from conftest import binary_model_and_inputs

    model, x_eval = binary_model_and_inputs()
    weights = ser_oracle.mlp_weights_from_sklearn(model)
    assert weights.out_activation == "logistic" and len(weights.classes) == 2
    np.testing.assert_allclose(ser_oracle.mlp_predict_proba(weights, x_eval), model.predict_proba(x_eval), rtol=0, atol=1e-12)
    assert list(ser_oracle.mlp_predict(weights, x_eval)) == model.predict(x_eval).tolist()
