"""Generates tests/golden/*.npz by running the REFERENCE's own host code over the oracle shim.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What runs unchanged from /root/reference: ser/_internal/utils/dsp.py
(extract_feature_from_signal), ser/_internal/repr/handcrafted.py (HandcraftedBackend),
ser/_internal/models/fast_path.py (predict_emotions_detailed_with_model, segment merge),
ser/_internal/utils/audio_utils.py (_prepare_audio_buffer).  Only the third-party
arithmetic (librosa, soundfile) is the oracle's restatement (oracle/shim), because
librosa 0.11.0 is not installable here.  The fixtures therefore pin the
DRIVER logic of oracle/ser_oracle.py and of ser_b200's host mirror to the reference's real
code, and give the GPU parity tests fixed vectors that need no /root/reference at run time.
"""

from __future__ import annotations

import logging
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle" / "shim"))  # `import librosa` -> the oracle restatement
sys.path.insert(0, "/root/reference")

from ser_b200 import synth  # noqa: E402

import ser._internal.utils.dsp as ref_dsp  # noqa: E402
from ser._internal.models import fast_path as ref_fast_path  # noqa: E402
from ser._internal.features import FeatureFrame  # noqa: E402
from ser._internal.repr.handcrafted import HandcraftedBackend  # noqa: E402
from ser._internal.utils.audio_utils import _prepare_audio_buffer  # noqa: E402
from ser.config import FeatureFlags  # noqa: E402

OUT = Path(__file__).resolve().parent


def clip_cases():
    """(name, sample_rate, pcm16) -- every shape class the reference path distinguishes."""
    S = synth.ClipSpec
    cases = [
        ("c16k_3s", 16000, synth.clip_pcm16(S(0, 1, 3), 16000, 48000)),
        ("c48k_3p5s", 48000, synth.clip_pcm16(S(1, 7, 5, 2, 1, 2), 48000, 168000)),
        ("c22k_2s", 22050, synth.clip_pcm16(S(2, 12, 8, 1, 2, 1), 22050, 44100)),
        ("c44k_1s", 44100, synth.clip_pcm16(S(3, 20, 2, 2, 2, 2), 44100, 44100)),
        ("c16k_tail_5937", 16000, synth.clip_pcm16(S(4, 3, 6), 16000, 5937)),
        ("c16k_2048", 16000, synth.clip_pcm16(S(5, 4, 4), 16000, 2048)),
        ("c16k_short_1500", 16000, synth.clip_pcm16(S(6, 5, 7), 16000, 1500)),      # n_fft=1500
        ("c16k_short_1001", 16000, synth.clip_pcm16(S(7, 6, 1), 16000, 1001)),      # odd n_fft
        ("c16k_short_300", 16000, synth.clip_pcm16(S(8, 8, 2), 16000, 300)),        # padded to 512
        ("c48k_short_512", 48000, synth.clip_pcm16(S(9, 9, 3), 48000, 512)),        # empty mel filters
        ("sine16k_1p5s", 16000, synth.pure_sine_pcm16(16000, 1.5, 180.0 + 22.0 * 3 + 7.0 * 2)),
        ("silence16k", 16000, np.zeros(20000, dtype=np.int16)),
    ]
    return cases


def main() -> None:
    logging.disable(logging.CRITICAL)
    import warnings as _warnings

    _warnings.simplefilter("ignore")  # the shim's module names differ from librosa's warning filters
    payload: dict[str, np.ndarray] = {}
    names = []
    for name, sr, pcm in clip_cases():
        audio = _prepare_audio_buffer(pcm.astype(np.float32) / np.float32(32768.0))
        assert np.array_equal(audio, synth.decode_pcm16(pcm)), name
        full = ref_dsp.extract_feature_from_signal(audio, sr, feature_flags=FeatureFlags())
        slim = ref_dsp.extract_feature_from_signal(audio, sr, feature_flags=FeatureFlags(tonnetz=False))
        assert full.shape == (193,) and slim.shape == (187,)
        assert np.array_equal(full[:187], slim)
        payload[f"{name}/pcm"] = pcm
        payload[f"{name}/sr"] = np.asarray(sr)
        payload[f"{name}/features"] = full
        names.append(name)
        print(f"{name:>18s} sr={sr:6d} n={pcm.size:7d} ok")
    payload["names"] = np.asarray(names)

    # sliding-window inference shape of config c1 (sample.wav: 16 kHz, 69 937 samples, 5 windows)
    pcm = synth.clip_pcm16(synth.ClipSpec(20, 2, 5), 16000, 69937)
    audio = synth.decode_pcm16(pcm)
    backend = HandcraftedBackend()
    encoded = backend.encode_sequence(audio, 16000)
    payload["seq/pcm"] = pcm
    payload["seq/sr"] = np.asarray(16000)
    payload["seq/embeddings"] = encoded.embeddings
    payload["seq/starts"] = encoded.frame_start_seconds
    payload["seq/ends"] = encoded.frame_end_seconds
    payload["seq/vector"] = backend.extract_vector(audio, 16000)
    print("sequence", encoded.embeddings.shape, encoded.frame_start_seconds, encoded.frame_end_seconds)

    # classifier: the reference's Pipeline(StandardScaler, MLPClassifier(300)) fitted on synthetic rows
    from sklearn.neural_network import MLPClassifier
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    rng = np.random.default_rng(7)
    labels = np.asarray(sorted(synth.RAVDESS_EMOTIONS.values()))
    centers = rng.standard_normal((len(labels), 193)) * 2.0
    y_train = rng.integers(0, len(labels), size=640)
    x_train = centers[y_train] + rng.standard_normal((640, 193))
    x_train[:, 180:187] = 0.0  # contrast columns are constant zero on the reference path (F5)
    model = Pipeline([
        ("scaler", StandardScaler()),
        ("classifier", MLPClassifier(alpha=0.01, batch_size=256, epsilon=1e-8, hidden_layer_sizes=(300,),
                                     learning_rate="adaptive", max_iter=60, random_state=42)),
    ])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(x_train, labels[y_train])
    clf = model.named_steps["classifier"]
    scaler = model.named_steps["scaler"]
    payload["mlp/mean"] = scaler.mean_
    payload["mlp/scale"] = scaler.scale_
    payload["mlp/w1"] = clf.coefs_[0]
    payload["mlp/b1"] = clf.intercepts_[0]
    payload["mlp/w2"] = clf.coefs_[1]
    payload["mlp/b2"] = clf.intercepts_[1]
    payload["mlp/classes"] = np.asarray(clf.classes_.tolist())
    x_eval = centers[rng.integers(0, len(labels), size=64)] + 1.5 * rng.standard_normal((64, 193))
    x_eval[:, 180:187] = 0.0
    x_eval = x_eval.astype(np.float32).astype(np.float64)  # the inference path's f32 round trip (F9)
    payload["mlp/x_eval"] = x_eval
    payload["mlp/proba"] = model.predict_proba(x_eval)
    payload["mlp/labels"] = np.asarray([str(v) for v in model.predict(x_eval)])

    # fast_path.predict_emotions_detailed_with_model on a crafted 12-frame sequence
    frames = [
        FeatureFrame(start_seconds=float(i), end_seconds=float(min(i + 3, 13.25)), features=x_eval[i // 3])
        for i in range(12)
    ]
    result = ref_fast_path.predict_emotions_detailed_with_model(
        "unused.wav", model=model, expected_feature_size=193, output_schema_version="v1",
        extract_feature_frames_fn=lambda _path: frames, logger=logging.getLogger("golden"),
    )
    payload["fast/frame_rows"] = np.asarray([i // 3 for i in range(12)])
    payload["fast/frame_starts"] = np.asarray([f.start_seconds for f in result.frames])
    payload["fast/frame_ends"] = np.asarray([f.end_seconds for f in result.frames])
    payload["fast/frame_labels"] = np.asarray([f.emotion for f in result.frames])
    payload["fast/frame_conf"] = np.asarray([f.confidence for f in result.frames])
    payload["fast/seg_labels"] = np.asarray([s.emotion for s in result.segments])
    payload["fast/seg_starts"] = np.asarray([s.start_seconds for s in result.segments])
    payload["fast/seg_ends"] = np.asarray([s.end_seconds for s in result.segments])
    payload["fast/seg_conf"] = np.asarray([s.confidence for s in result.segments])
    payload["fast/seg_proba"] = np.asarray(
        [[s.probabilities[str(c)] for c in clf.classes_] for s in result.segments]
    )
    print("segments", [(s.emotion, s.start_seconds, s.end_seconds) for s in result.segments])

    np.savez_compressed(OUT / "fast_profile_golden.npz", **payload)
    print("wrote", OUT / "fast_profile_golden.npz")


if __name__ == "__main__":
    main()
