"""Row N1: 16-bit PCM files prepared on the device.

The device step must be BIT-IDENTICAL to the reference's ``_prepare_audio_buffer`` after a
soundfile / librosa.load decode (ser/_internal/utils/audio_utils.py:28-60, 104-113): float32
``x / 32768``, channel mean, whole-file peak normalisation, all-zero files stay zero.  The
expected values come from ``ser_b200.audio.prepare_audio_buffer`` -- a numpy restatement whose
equality with the reference's own function is asserted where the golden fixtures are generated
(tests/golden/make_golden.py:72) -- and from numpy directly.  Feature rows through the PCM16
entries must then equal, bit for bit, the rows of the float32 entries fed with that prepared audio.
"""

from __future__ import annotations

import wave

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _numpy_prepare(pcm: np.ndarray, channels: int) -> np.ndarray:
    """audio_utils.py:28-60 in plain numpy, the way librosa.load + _prepare_audio_buffer run it."""
    x = pcm.astype(np.float32) / np.float32(32768.0)
    if channels > 1:
        x = x.reshape(-1, channels)
        x = np.asarray(np.mean(x, axis=1, dtype=np.float32), dtype=np.float32)
    x = np.nan_to_num(x, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    peak = float(np.max(np.abs(x)))
    if peak == 0:
        return np.zeros_like(x)
    return x / peak


def _files():
    rng = np.random.default_rng(42)
    files = [
        (rng.integers(-20000, 20000, size=48000).astype(np.int16), 1),
        (rng.integers(-32768, 32768, size=2 * 30001).astype(np.int16), 2),       # stereo, odd frame count
        (rng.integers(-9000, 9000, size=3 * 5003).astype(np.int16), 3),          # three channels
        (np.zeros(7000, dtype=np.int16), 1),                                     # digital silence
        (np.concatenate([[-32768], rng.integers(-100, 100, size=4098)]).astype(np.int16), 1),   # peak at -1.0
        (rng.integers(-5, 5, size=13).astype(np.int16), 1),                      # shorter than one vector load
        (rng.integers(-3000, 3000, size=6 * 2500).astype(np.int16), 6),
    ]
    return files


def test_device_preparation_is_bit_identical_to_the_reference_recipe(gpu_ctx):
    from ser_b200.audio import prepare_audio_buffer

    files = _files()
    got = gpu_ctx.prepare_pcm16_files_host([f for f, _ in files], [c for _, c in files])
    for (pcm, ch), out in zip(files, got):
        expect = _numpy_prepare(pcm, ch)
        assert out.dtype == np.float32 and out.shape == expect.shape
        np.testing.assert_array_equal(out, expect)
        decoded = pcm.astype(np.float32) / np.float32(32768.0)
        np.testing.assert_array_equal(out, prepare_audio_buffer(decoded.reshape(-1, ch) if ch > 1 else decoded))
    assert gpu_ctx.prepare_pcm16_files_host([], []) == []


def test_pcm16_features_equal_the_float_path_bit_for_bit(gpu_ctx):
    """Ragged clips over several files (mono / stereo / different lengths), all five groups."""
    from ser_b200.config import FeatureFlags, flag_bits

    rng = np.random.default_rng(7)
    sr = 16000
    t = np.arange(70000) / sr
    tone = 0.3 * np.sin(2 * np.pi * 233.0 * t) + 0.05 * rng.standard_normal(t.size)
    files = [
        (np.rint(tone * 20000).astype(np.int16), 1),
        (np.rint(np.stack([tone[:40000], 0.5 * tone[100:40100]], axis=1) * 15000).astype(np.int16).reshape(-1), 2),
        (np.rint(tone[:5001] * 9000).astype(np.int16), 1),
    ]
    clip_file = np.asarray([0, 0, 0, 1, 1, 2, 2], dtype=np.int64)
    clip_starts = np.asarray([0, 16000, 60000, 0, 30000, 0, 4000], dtype=np.int64)
    clip_lengths = np.asarray([48000, 48000, 10000, 40000, 9999, 5001, 700], dtype=np.int64)
    bits = flag_bits(FeatureFlags())
    got = gpu_ctx.features_host_pcm16([f for f, _ in files], [c for _, c in files], clip_file, clip_starts, clip_lengths, sr, bits)
    assert got.shape == (7, 193) and np.all(np.isfinite(got))
    prepared = [_numpy_prepare(f, c) for f, c in files]
    for i in range(clip_file.size):
        wave_f = prepared[clip_file[i]]
        expect = gpu_ctx.features_host(wave_f, clip_starts[i: i + 1], clip_lengths[i: i + 1], sr, bits)
        np.testing.assert_array_equal(got[i], expect[0])


def test_pcm16_entry_validation(gpu_ctx):
    from ser_b200.config import FeatureFlags, flag_bits

    bits = flag_bits(FeatureFlags())
    pcm = np.ones(4000, dtype=np.int16)
    z = np.zeros(1, dtype=np.int64)
    assert gpu_ctx.features_host_pcm16([], [], [], [], [], 16000, bits).shape == (0, 193)
    with pytest.raises(ValueError, match="Sample rate must be a positive integer."):
        gpu_ctx.features_host_pcm16([pcm], 1, z, z, np.asarray([4000]), 0, bits)
    with pytest.raises(ValueError, match="lies outside its file"):
        gpu_ctx.features_host_pcm16([pcm], 1, z, np.asarray([10]), np.asarray([4000]), 16000, bits)
    with pytest.raises(ValueError, match="non-decreasing"):
        gpu_ctx.features_host_pcm16([pcm, pcm], 1, np.asarray([1, 0]), np.asarray([0, 0]), np.asarray([4000, 4000]), 16000, bits)
    with pytest.raises(ValueError, match="Audio contains no samples."):
        gpu_ctx.features_host_pcm16([pcm], 1, z, z, z, 16000, bits)
    with pytest.raises(TypeError):
        gpu_ctx.features_host_pcm16([pcm.astype(np.float32)], 1, z, z, np.asarray([4000]), 16000, bits)
    # an all-zero file is a valid (silent) clip, as in the reference
    rows = gpu_ctx.features_host_pcm16([np.zeros(4000, dtype=np.int16)], 1, z, z, np.asarray([4000]), 16000, bits)
    assert np.all(np.isfinite(rows))


def test_file_level_seams_take_the_pcm16_path(tmp_path, gpu_ctx):
    """extract_feature_frames / extract_feature on a 16-bit WAV file == the float path on the decoded,
    prepared audio; a stereo file and a 24-bit file (float path) included."""
    from ser_b200 import audio
    from ser_b200.feature_extractor import extract_feature, extract_feature_frames
    from ser_b200.handcrafted import HandcraftedBackend

    rng = np.random.default_rng(3)
    sr = 16000
    t = np.arange(69937) / sr
    x = 0.4 * np.sin(2 * np.pi * 311.0 * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 3 * t)) + 0.02 * rng.standard_normal(t.size)
    mono = np.rint(x * 12000).astype("<i2")
    stereo = np.stack([mono, np.roll(mono, 37) // 2], axis=1).astype("<i2")
    for name, data, ch in (("mono.wav", mono, 1), ("stereo.wav", stereo, 2)):
        path = tmp_path / name
        with wave.open(str(path), "wb") as handle:
            handle.setnchannels(ch)
            handle.setsampwidth(2)
            handle.setframerate(sr)
            handle.writeframes(data.tobytes())
        raw = audio.read_pcm16_file(str(path))
        assert raw is not None and raw[1] == ch and raw[2] == sr
        prepared, sr2 = audio.read_audio_file(str(path))
        assert sr2 == sr
        frames = extract_feature_frames(str(path))
        encoded = HandcraftedBackend().encode_sequence(prepared, sr)
        assert [f.start_seconds for f in frames] == encoded.frame_start_seconds.tolist()
        assert [f.end_seconds for f in frames] == encoded.frame_end_seconds.tolist()
        np.testing.assert_array_equal(np.vstack([f.features for f in frames]), encoded.embeddings.astype(np.float64))
        np.testing.assert_array_equal(extract_feature(str(path)), HandcraftedBackend().extract_vector(prepared, sr))
    # segment reads follow librosa.load(offset=, duration=)
    seg_raw = audio.read_pcm16_file(str(tmp_path / "mono.wav"), start_seconds=0.5, duration_seconds=1.25)
    seg_f, _ = audio.read_audio_file(str(tmp_path / "mono.wav"), start_seconds=0.5, duration_seconds=1.25)
    assert seg_raw[0].size == seg_f.size == 20000
    np.testing.assert_array_equal(gpu_ctx.prepare_pcm16_files_host([seg_raw[0]], 1)[0], seg_f)
    # 24-bit PCM is not the int16 fast path
    path24 = tmp_path / "p24.wav"
    with wave.open(str(path24), "wb") as handle:
        handle.setnchannels(1)
        handle.setsampwidth(3)
        handle.setframerate(sr)
        handle.writeframes(b"".join(int(v).to_bytes(3, "little", signed=True) for v in (mono[:4000].astype(np.int32) * 256)))
    assert audio.read_pcm16_file(str(path24)) is None
    assert len(extract_feature_frames(str(path24))) == 1


def test_infer_host_pcm16_equals_infer_host(gpu_ctx):
    """The bench's end-to-end call: labels / probabilities / rows identical to the float32 entry."""
    from ser_b200 import mlp, synth
    from ser_b200.config import FeatureFlags, flag_bits
    from ser_b200.handcrafted import frame_bounds

    sr, n, n_files = 48000, 168000, 6
    bits = flag_bits(FeatureFlags())
    pcm = np.stack([synth.clip_pcm16(s, sr, n) for s in synth.ravdess_specs(n_files)])       # one contiguous buffer
    w_starts, w_ends = frame_bounds(n, sr, 3, 1)
    clip_file = np.repeat(np.arange(n_files, dtype=np.int64), w_starts.size)
    clip_starts = np.tile(w_starts, n_files)
    clip_lengths = np.tile(w_ends - w_starts, n_files)
    rng = np.random.default_rng(0)
    weights = mlp.MlpWeights(mean=rng.standard_normal(193), scale=1.0 + rng.random(193),
                             w1=rng.standard_normal((193, 300)) * 0.1, b1=rng.standard_normal(300) * 0.1,
                             w2=rng.standard_normal((300, 8)) * 0.1, b2=rng.standard_normal(8) * 0.1,
                             classes=tuple(sorted(synth.RAVDESS_EMOTIONS.values())), out_activation=0)
    with mlp.session(weights, 0) as (ctx, _w):
        f16, p16, l16 = ctx.infer_host_pcm16([pcm[i] for i in range(n_files)], 1, clip_file, clip_starts, clip_lengths, sr, bits)
        wave_f = np.concatenate([synth.decode_pcm16(pcm[i]) for i in range(n_files)])
        starts = clip_file * n + clip_starts
        f32, p32, l32 = ctx.infer_host(wave_f, starts, clip_lengths, sr, bits)
    np.testing.assert_array_equal(f16, f32)
    np.testing.assert_array_equal(p16, p32)
    np.testing.assert_array_equal(l16, l32)
