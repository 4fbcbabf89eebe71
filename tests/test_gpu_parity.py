"""GPU parity: the CUDA path, called through the C ABI, against the oracle and the golden vectors.

Tolerances (stated here as the contract):
* STFT magnitudes: <= 3e-6 of the column maximum (float32 FFT vs the reference's float64 FFT
  rounded to complex64).
* pooled features: scaled error <= 1e-4 per group (conftest.group_errors; the raw relative
  error is printed alongside).
* tuning bin, labels, segment boundaries, timestamps: identical.
* probabilities: <= 1e-12 absolute (float64 on both sides, different summation order).
"""

from __future__ import annotations

import numpy as np
import pytest

from conftest import group_errors

pytestmark = pytest.mark.gpu

TOL = 1e-4
MAIN_CASES = ["c16k_3s", "c48k_3p5s", "c22k_2s", "c44k_1s", "c16k_tail_5937", "c16k_2048", "sine16k_1p5s",
              "silence16k"]
SHORT_CASES = ["c16k_short_1500", "c16k_short_1001", "c16k_short_300", "c48k_short_512"]


def _audio(golden, name):
    from ser_b200 import synth

    return synth.decode_pcm16(golden[f"{name}/pcm"]), int(golden[f"{name}/sr"])


def _flags187():
    from ser_b200.config import FeatureFlags

    return FeatureFlags(tonnetz=False)


def test_native_library_is_the_compute_path(gpu_ctx):
    from ser_b200 import _native

    assert _native.device_count() >= 1
    before = gpu_ctx.launch_count
    from ser_b200 import dsp

    dsp.extract_feature_from_signal(np.ones(4096, dtype=np.float32), 16000, feature_flags=_flags187())
    assert gpu_ctx.launch_count > before


@pytest.mark.parametrize("name", ["c16k_3s", "c48k_3p5s", "c22k_2s", "sine16k_1p5s"])
def test_stft_magnitude_matches_oracle(golden, gpu_ctx, name):
    from oracle.shim import librosa

    audio, _ = _audio(golden, name)
    got = gpu_ctx.debug_stft_host(audio)                 # [T][1025]
    ref = np.abs(librosa.stft(audio, n_fft=2048)).T      # oracle: float64 FFT -> complex64 -> |.|
    assert got.shape == ref.shape
    col_max = np.maximum(ref.max(axis=1, keepdims=True), 1e-30)
    err = np.max(np.abs(got - ref) / col_max)
    print(f"{name}: stft max err / column max = {err:.3e}")
    assert err <= 3e-6


@pytest.mark.parametrize("name", MAIN_CASES + SHORT_CASES)
def test_features_match_golden(golden, name):
    from ser_b200 import dsp

    audio, sr = _audio(golden, name)
    got = dsp.extract_feature_from_signal(audio, sr, feature_flags=_flags187())
    assert got.dtype == np.float64 and got.shape == (187,)
    report = group_errors(got, golden[f"{name}/features"][:187])
    print(name, {k: f"{v[0]:.2e} (raw {v[1]:.2e})" for k, v in report.items()})
    for group, (scaled, _raw) in report.items():
        assert scaled <= TOL, f"{name}/{group}: scaled error {scaled:.3e}"
    assert np.all(got[180:187] == 0.0)


@pytest.mark.parametrize("name", MAIN_CASES + SHORT_CASES)
def test_tuning_bin_matches_oracle(golden, gpu_ctx, name):
    from oracle.shim import librosa
    from ser_b200 import dsp

    audio, sr = _audio(golden, name)
    dsp.extract_feature_from_signal(audio, sr, feature_flags=_flags187())
    got = int(gpu_ctx.debug_last_tuning(1)[0])
    padded = audio if audio.size >= 512 else np.pad(audio, (0, 512 - audio.size))
    n_fft = min(padded.size, 2048)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S = np.abs(librosa.stft(padded, n_fft=n_fft))
        ref = librosa.estimate_tuning(S=S, sr=sr, bins_per_octave=12)
    assert np.linspace(-0.5, 0.5, 101)[got] == pytest.approx(ref, abs=1e-12), (got, ref)


def test_flag_subsets_select_groups(golden):
    from ser_b200 import dsp
    from ser_b200.config import FeatureFlags

    audio, sr = _audio(golden, "c16k_3s")
    full = dsp.extract_feature_from_signal(audio, sr, feature_flags=_flags187())
    mel_only = dsp.extract_feature_from_signal(audio, sr, feature_flags=FeatureFlags(mfcc=False, chroma=False, contrast=False, tonnetz=False))
    np.testing.assert_array_equal(mel_only, full[52:180])
    no_mel = dsp.extract_feature_from_signal(audio, sr, feature_flags=FeatureFlags(mel=False, tonnetz=False))
    np.testing.assert_array_equal(no_mel, np.concatenate([full[:52], full[180:187]]))
    assert dsp.extract_feature_from_signal(audio, sr, feature_flags=FeatureFlags(False, False, False, False, False)).shape == (0,)


def test_sliding_windows_match_golden_sequence(golden):
    from ser_b200 import synth
    from ser_b200.handcrafted import HandcraftedBackend

    audio = synth.decode_pcm16(golden["seq/pcm"])
    backend = HandcraftedBackend(feature_flags=_flags187())
    encoded = backend.encode_sequence(audio, int(golden["seq/sr"]))
    # timestamps are exact float64 quotients -> bit identical
    np.testing.assert_array_equal(encoded.frame_start_seconds, golden["seq/starts"])
    np.testing.assert_array_equal(encoded.frame_end_seconds, golden["seq/ends"])
    assert encoded.embeddings.dtype == np.float32 and encoded.embeddings.shape == (5, 187)
    report = group_errors(encoded.embeddings, golden["seq/embeddings"][:, :187])
    print({k: f"{v[0]:.2e} (raw {v[1]:.2e})" for k, v in report.items()})
    for group, (scaled, _raw) in report.items():
        assert scaled <= TOL, group
    vec = backend.extract_vector(audio, int(golden["seq/sr"]))
    for group, (scaled, _raw) in group_errors(vec, golden["seq/vector"][:187]).items():
        assert scaled <= TOL, group


def test_batch_equals_single_calls_bitwise(golden):
    """Ragged batching must not change a single bit of any clip's row (clips are independent)."""
    from ser_b200 import dsp

    names = MAIN_CASES + SHORT_CASES
    sr16 = [n for n in names if int(golden[f"{n}/sr"]) == 16000]
    clips = [_audio(golden, n)[0] for n in sr16]
    batch = dsp.extract_features_batch(clips, 16000, feature_flags=_flags187())
    for row, clip in zip(batch, clips):
        single = dsp.extract_feature_from_signal(clip, 16000, feature_flags=_flags187())
        np.testing.assert_array_equal(row, single)
    again = dsp.extract_features_batch(clips, 16000, feature_flags=_flags187())
    np.testing.assert_array_equal(batch, again)  # deterministic run to run


def test_unaligned_windows_take_the_fallback_loader(golden):
    """Window starts that are not 16-byte aligned bypass the TMA bulk copy; results must not move."""
    from ser_b200 import dsp

    audio, sr = _audio(golden, "c16k_3s")
    base = dsp.extract_features_ragged(audio, np.asarray([4000]), np.asarray([20000]), sr, feature_flags=_flags187())
    shifted = np.concatenate([np.zeros(3, dtype=np.float32), audio])
    moved = dsp.extract_features_ragged(shifted, np.asarray([4003]), np.asarray([20000]), sr, feature_flags=_flags187())
    np.testing.assert_array_equal(base, moved)


def test_error_behaviour_matches_reference():
    from ser_b200 import dsp
    from ser_b200.config import FeatureFlags

    good = np.zeros(4096, dtype=np.float32)
    with pytest.raises(ValueError, match="Sample rate must be a positive integer."):
        dsp.extract_feature_from_signal(good, 0)
    with pytest.raises(ValueError, match=r"Audio must be mono \(1D array\)."):
        dsp.extract_feature_from_signal(np.zeros((2, 100), dtype=np.float32), 16000)
    with pytest.raises(ValueError, match="Audio contains no samples."):
        dsp.extract_feature_from_signal(np.zeros(0, dtype=np.float32), 16000)
    bad = good.copy()
    bad[100] = np.nan
    with pytest.raises(ValueError, match="Audio buffer is not finite everywhere."):
        dsp.extract_feature_from_signal(bad, 16000, feature_flags=_flags187())
    # device-side check of the same condition on the batched entry (no host pre-scan there)
    with pytest.raises(ValueError, match="Audio buffer is not finite everywhere."):
        dsp.extract_features_ragged(bad, np.asarray([0]), np.asarray([4096]), 16000, feature_flags=_flags187())
    # librosa.feature.spectral_contrast rejects sr <= 12800
    with pytest.raises(dsp.ParameterError, match="Nyquist"):
        dsp.extract_feature_from_signal(good, 8000, feature_flags=_flags187())
    ok = dsp.extract_feature_from_signal(good, 8000, feature_flags=FeatureFlags(contrast=False, tonnetz=False))
    assert ok.shape == (180,)
    # a sample rate whose constant-Q plan depends on the tuning estimate: valid for the reference,
    # unsupported here -- NOT a ValueError, and the 187-d groups still work
    from ser_b200 import _native

    with pytest.raises(_native.UnsupportedConfigurationError, match="tuning estimate"):
        dsp.extract_feature_from_signal(good, 20600)
    assert dsp.extract_feature_from_signal(good, 20600, feature_flags=_flags187()).shape == (187,)


def test_mlp_matches_sklearn_golden(golden, gpu_ctx):
    from ser_b200 import mlp

    weights = mlp.MlpWeights(golden["mlp/mean"], golden["mlp/scale"], golden["mlp/w1"], golden["mlp/b1"],
                             golden["mlp/w2"], golden["mlp/b2"], tuple(golden["mlp/classes"].tolist()), 0)
    labels, proba = mlp.predict(weights, golden["mlp/x_eval"])
    assert labels == golden["mlp/labels"].tolist()
    np.testing.assert_allclose(proba, golden["mlp/proba"], rtol=0, atol=1e-12)


def test_fast_path_segments_match_reference(golden):
    import logging

    from ser_b200 import fast_path, mlp
    from ser_b200.feature_extractor import FeatureFrame

    weights = mlp.MlpWeights(golden["mlp/mean"], golden["mlp/scale"], golden["mlp/w1"], golden["mlp/b1"],
                             golden["mlp/w2"], golden["mlp/b2"], tuple(golden["mlp/classes"].tolist()), 0)
    rows = golden["fast/frame_rows"]
    frames = [FeatureFrame(float(s), float(e), golden["mlp/x_eval"][r])
              for s, e, r in zip(golden["fast/frame_starts"], golden["fast/frame_ends"], rows)]
    result = fast_path.predict_emotions_detailed_with_model(
        "unused.wav", model=weights, expected_feature_size=193, output_schema_version="v1",
        extract_feature_frames_fn=lambda _p: frames, logger=logging.getLogger("test"))
    assert [f.emotion for f in result.frames] == golden["fast/frame_labels"].tolist()
    np.testing.assert_allclose([f.confidence for f in result.frames], golden["fast/frame_conf"], atol=1e-12)
    assert [s.emotion for s in result.segments] == golden["fast/seg_labels"].tolist()
    np.testing.assert_array_equal([s.start_seconds for s in result.segments], golden["fast/seg_starts"])
    np.testing.assert_array_equal([s.end_seconds for s in result.segments], golden["fast/seg_ends"])
    np.testing.assert_allclose([s.confidence for s in result.segments], golden["fast/seg_conf"], atol=1e-12)
    with pytest.raises(ValueError, match="Feature vector size mismatch for loaded model."):
        fast_path.predict_emotions_detailed_with_model(
            "unused.wav", model=weights, expected_feature_size=187, output_schema_version="v1",
            extract_feature_frames_fn=lambda _p: frames, logger=logging.getLogger("test"))


def test_prepare_pcm16_matches_reference_normalisation(golden, gpu_ctx):
    from ser_b200 import synth

    pcm = golden["c16k_3s/pcm"]
    np.testing.assert_array_equal(gpu_ctx.prepare_pcm16_host(pcm), synth.decode_pcm16(pcm))
    zeros = np.zeros(1000, dtype=np.int16)
    np.testing.assert_array_equal(gpu_ctx.prepare_pcm16_host(zeros), np.zeros(1000, dtype=np.float32))


def test_full_size_properties_c2_batch(gpu_ctx):
    """Config c2 (1440 clips x 168000 samples @ 48 kHz) through size-independent properties:
    scaling the input by 2 multiplies mel power by exactly 4, leaves chroma unchanged and shifts
    MFCC c0 by 10 log10(4) sqrt(128); every row equals the row of the same clip extracted alone."""
    import torch

    from ser_b200 import _native, synth
    from ser_b200.config import flag_bits

    n_clips, n_samples, sr = 1440, 168000, 48000
    wave = synth.batch_audio_torch(n_clips, sr, n_samples, device="cuda")
    starts = np.arange(n_clips, dtype=np.int64) * n_samples
    lengths = np.full(n_clips, n_samples, dtype=np.int64)
    bits = flag_bits(_flags187())
    out1 = torch.empty((n_clips, 187), dtype=torch.float32, device="cuda")
    out2 = torch.empty_like(out1)
    half = wave * 0.5
    torch.cuda.synchronize()            # inputs were produced on torch's stream
    stream = 0                          # 0 = the context's own stream
    gpu_ctx.features_device(wave.data_ptr(), wave.numel(), starts, lengths, sr, bits, out1.data_ptr(), stream)
    gpu_ctx.features_device(half.data_ptr(), half.numel(), starts, lengths, sr, bits, out2.data_ptr(), stream)
    torch.cuda.synchronize()
    a, b = out1.cpu().numpy(), out2.cpu().numpy()
    assert np.all(np.isfinite(a))
    np.testing.assert_array_equal(a[:, 52:180], 4.0 * b[:, 52:180])           # mel power, exact
    np.testing.assert_array_equal(a[:, 40:52], b[:, 40:52])                   # chroma, exact
    shift = 10.0 * np.log10(4.0) * np.sqrt(128.0)
    np.testing.assert_allclose(a[:, 0] - b[:, 0], shift, rtol=0, atol=2e-3)   # MFCC c0
    np.testing.assert_allclose(a[:, 1:40], b[:, 1:40], rtol=0, atol=2e-3)
    # spot-check independence from batch position
    for idx in (0, 719, 1439):
        single = torch.empty((1, 187), dtype=torch.float32, device="cuda")
        gpu_ctx.features_device(wave.data_ptr(), wave.numel(), starts[idx : idx + 1], lengths[idx : idx + 1], sr,
                                bits, single.data_ptr(), stream)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(single.cpu().numpy()[0], a[idx])
    assert _native.device_count() >= 1


def test_clip_list_entry_errors_and_empty_batch(gpu_ctx):
    """serb_features_host_clips: empty batch, an empty clip (reference text), zero feature groups."""
    from ser_b200 import _native, dsp
    from ser_b200.config import FeatureFlags, flag_bits

    bits = flag_bits(FeatureFlags())
    assert gpu_ctx.features_host_clips([], 16000, bits).shape == (0, 193)
    good = np.ones(3000, dtype=np.float32)
    with pytest.raises(ValueError, match="Audio contains no samples."):
        gpu_ctx.features_host_clips([good, np.zeros(0, dtype=np.float32)], 16000, bits)
    assert gpu_ctx.features_host_clips([good], 16000, 0).shape == (1, 0)
    rows = dsp.extract_features_batch([good, good[:700]], 16000)
    np.testing.assert_array_equal(rows[0], dsp.extract_feature_from_signal(good, 16000))
    np.testing.assert_array_equal(rows[1], dsp.extract_feature_from_signal(good[:700], 16000))
    assert _native.device_count() >= 1


def test_binary_logistic_classifier_matches_sklearn(gpu_ctx):
    """Two emotion classes: scikit-learn's MLPClassifier has a single logistic output unit and
    predict_proba returns [1 - p, p] (fast_path.py:48,181 call the same methods for any label set).
    The fused kernel's SERB_OUT_LOGISTIC branch against the real fitted model."""
    from ser_b200 import mlp
    from conftest import binary_model_and_inputs

    model, x_eval = binary_model_and_inputs()
    labels, proba = mlp.predict(model, x_eval)
    assert proba.shape == (64, 2)
    assert labels == model.predict(x_eval).tolist()
    np.testing.assert_allclose(proba, model.predict_proba(x_eval), rtol=0, atol=1e-12)
    assert set(labels) == {"happy", "sad"}                     # both classes occur in the evaluation set
