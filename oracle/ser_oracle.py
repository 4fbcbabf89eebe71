"""CPU restatement of the reference's fast-profile driver.  TEST INFRASTRUCTURE (oracle).

Restates, over ``oracle.shim.librosa`` (the librosa 0.11.0 restatement), the host logic of
the reference path so the oracle travels to the GPU box where /root/reference is absent:

* ``extract_feature_from_signal``  <- ser/_internal/utils/dsp.py:38-45, 67-151
* ``encode_sequence`` / ``extract_vector`` <- ser/_internal/repr/handcrafted.py:65-107, 124-137
* ``prepare_audio_buffer`` <- ser/_internal/utils/audio_utils.py:28-60
* ``mlp_predict`` / ``mlp_predict_proba`` <- sklearn Pipeline(StandardScaler, MLPClassifier)
  as built at ser/_internal/models/training_support.py:87-106 and called at
  ser/_internal/models/fast_path.py:48,181
* ``frame_confidences`` / ``segment_predictions`` <- ser/_internal/models/fast_path.py:19-144

tests/test_oracle_golden.py checks these restatements bit-for-bit against the reference's
own modules run over the same shim (fixtures from tests/golden/make_golden.py).
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass
from statistics import fmean

import numpy as np

from .shim import librosa

FEATURE_DIMS = {"mfcc": 40, "chroma": 12, "mel": 128, "contrast": 7, "tonnetz": 6}


@dataclass(frozen=True)
class FeatureFlags:
    """Mirror of ser.config.FeatureFlags (ser/_internal/config/schema.py:219-227)."""

    mfcc: bool = True
    chroma: bool = True
    mel: bool = True
    contrast: bool = True
    tonnetz: bool = True


def feature_dim(flags: FeatureFlags) -> int:
    """handcrafted.py:46-59."""
    return sum(dim for name, dim in FEATURE_DIMS.items() if getattr(flags, name))


def pad_audio_for_fft(audio: np.ndarray, minimum_window: int = 512) -> np.ndarray:
    """dsp.py:38-45."""
    if audio.size >= minimum_window:
        return audio
    return np.pad(audio, (0, minimum_window - audio.size), mode="constant")


def extract_feature_from_signal(audio, sample_rate, *, feature_flags: FeatureFlags | None = None):
    """dsp.py:67-151: the per-clip feature vector, float64, order mfcc|chroma|mel|contrast|tonnetz."""
    if sample_rate <= 0:
        raise ValueError("Sample rate must be a positive integer.")
    if audio.ndim != 1:
        raise ValueError("Audio must be mono (1D array).")
    if audio.size == 0:
        raise ValueError("Audio contains no samples.")
    flags = feature_flags if feature_flags is not None else FeatureFlags()
    prepared = pad_audio_for_fft(np.asarray(audio, dtype=np.float32))
    if not bool(np.all(np.isfinite(prepared))):
        raise ValueError("Audio buffer is not finite everywhere.")
    n_fft = min(prepared.size, 2048)
    parts: list[np.ndarray] = []
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", message=r"n_fft=\d+ is too large for input signal of length=.*")
        warnings.filterwarnings("ignore", message=r"Trying to estimate tuning from empty frequency set\.")
        stft_magnitude = np.abs(librosa.stft(prepared, n_fft=n_fft))
        stft_power_db = librosa.power_to_db(np.square(stft_magnitude), ref=np.max)
        if flags.mfcc:
            mfccs = np.mean(librosa.feature.mfcc(y=prepared, sr=sample_rate, n_mfcc=40, n_fft=n_fft), axis=1)
            parts.append(np.asarray(mfccs, dtype=np.float64))
        if flags.chroma:
            chroma = np.mean(librosa.feature.chroma_stft(S=stft_magnitude, sr=sample_rate, n_fft=n_fft), axis=1)
            parts.append(np.asarray(chroma, dtype=np.float64))
        if flags.mel:
            mel = np.mean(librosa.feature.melspectrogram(y=prepared, sr=sample_rate, n_fft=n_fft), axis=1)
            parts.append(np.asarray(mel, dtype=np.float64))
        if flags.contrast:
            contrast = np.mean(
                librosa.feature.spectral_contrast(S=stft_power_db, sr=sample_rate, n_fft=n_fft), axis=1
            )
            parts.append(np.asarray(contrast, dtype=np.float64))
        if flags.tonnetz:
            harmonic = librosa.effects.harmonic(prepared)
            tonnetz = np.mean(librosa.feature.tonnetz(y=harmonic, sr=sample_rate), axis=1)
            parts.append(np.asarray(tonnetz, dtype=np.float64))
    if not parts:
        return np.empty(0, dtype=np.float64)
    return np.concatenate(parts).astype(np.float64, copy=False)


def frame_bounds(n_samples: int, sample_rate: int, frame_size_seconds=3, frame_stride_seconds=1):
    """handcrafted.py:78-97: (start_index, end_index) of every inference window."""
    frame_length = max(1, int(round(frame_size_seconds * sample_rate)))
    frame_step = max(1, int(round(frame_stride_seconds * sample_rate)))
    bounds = []
    for start in range(0, n_samples, frame_step):
        end = min(start + frame_length, n_samples)
        if end - start == 0:
            continue
        bounds.append((start, end))
    return bounds


def encode_sequence(audio, sample_rate, *, frame_size_seconds=3, frame_stride_seconds=1,
                    feature_flags: FeatureFlags | None = None):
    """handcrafted.py:65-107 -> (embeddings float32 (W, dim), starts float64, ends float64)."""
    if sample_rate <= 0:
        raise ValueError("sample_rate must be a positive integer.")
    if audio.ndim != 1:
        raise ValueError("audio must be mono (1D array).")
    if audio.size == 0:
        raise ValueError("audio must contain at least one sample.")
    flags = feature_flags if feature_flags is not None else FeatureFlags()
    starts, ends, rows = [], [], []
    for start, end in frame_bounds(audio.size, sample_rate, frame_size_seconds, frame_stride_seconds):
        vec = extract_feature_from_signal(audio[start:end], sample_rate, feature_flags=flags)
        rows.append(np.asarray(vec, dtype=np.float32))
        starts.append(float(start) / float(sample_rate))
        ends.append(float(end) / float(sample_rate))
    if not rows:
        raise ValueError("Could not extract handcrafted features from provided audio.")
    return (
        np.vstack(rows).astype(np.float32, copy=False),
        np.asarray(starts, dtype=np.float64),
        np.asarray(ends, dtype=np.float64),
    )


def extract_vector(audio, sample_rate, *, feature_flags: FeatureFlags | None = None):
    """handcrafted.py:124-137."""
    return np.asarray(
        extract_feature_from_signal(audio, sample_rate, feature_flags=feature_flags), dtype=np.float64
    )


def prepare_audio_buffer(raw_audio):
    """audio_utils.py:28-60: NaN/Inf -> 0, channel mean, whole-file peak normalisation."""
    prepared = np.asarray(raw_audio, dtype=np.float32)
    prepared = np.nan_to_num(prepared, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    if prepared.ndim == 2:
        if prepared.shape[1] == 0:
            prepared = np.array([], dtype=np.float32)
        else:
            prepared = np.asarray(np.mean(prepared, axis=1, dtype=np.float32), dtype=np.float32)
    elif prepared.ndim != 1:
        raise OSError(f"Unsupported audio shape: {prepared.shape}")
    if prepared.size == 0:
        raise OSError("Audio file contains no samples.")
    max_abs = float(np.max(np.abs(prepared)))
    if max_abs == 0:
        return np.zeros_like(prepared)
    return prepared / max_abs


# ----------------------------------------------------------------------------
# scikit-learn Pipeline(StandardScaler, MLPClassifier(hidden=(300,), relu)) forward pass
# ----------------------------------------------------------------------------
@dataclass(frozen=True)
class MlpWeights:
    """What the fused MLP needs from the artifact."""

    mean: np.ndarray       # (F,)  scaler.mean_   (zeros if with_mean=False)
    scale: np.ndarray      # (F,)  scaler.scale_  (ones if with_std=False)
    coefs: tuple           # ((F,H), (H,C')) float64
    intercepts: tuple      # ((H,), (C',))
    classes: tuple         # C labels
    out_activation: str    # "softmax" | "logistic"


def mlp_weights_from_sklearn(model) -> MlpWeights:
    """Pulls the arrays out of a fitted Pipeline([scaler, classifier]) or bare MLPClassifier."""
    scaler = None
    clf = model
    if hasattr(model, "named_steps"):
        scaler = model.named_steps.get("scaler")
        clf = model.named_steps["classifier"]
    n_in = clf.coefs_[0].shape[0]
    mean = np.zeros(n_in)
    scale = np.ones(n_in)
    if scaler is not None:
        if getattr(scaler, "with_mean", True) and scaler.mean_ is not None:
            mean = np.asarray(scaler.mean_, dtype=np.float64)
        if getattr(scaler, "with_std", True) and scaler.scale_ is not None:
            scale = np.asarray(scaler.scale_, dtype=np.float64)
    return MlpWeights(
        mean=mean,
        scale=scale,
        coefs=tuple(np.asarray(c, dtype=np.float64) for c in clf.coefs_),
        intercepts=tuple(np.asarray(b, dtype=np.float64) for b in clf.intercepts_),
        classes=tuple(clf.classes_.tolist()),
        out_activation=str(clf.out_activation_),
    )


def mlp_predict_proba(weights: MlpWeights, X):
    """StandardScaler.transform + MLPClassifier._forward_pass_fast + predict_proba shaping."""
    act = (np.asarray(X, dtype=np.float64) - weights.mean) / weights.scale
    n_layers = len(weights.coefs)
    for i in range(n_layers):
        act = act @ weights.coefs[i]
        act += weights.intercepts[i]
        if i != n_layers - 1:
            np.maximum(act, 0, out=act)
    if weights.out_activation == "softmax":
        tmp = act - act.max(axis=1)[:, np.newaxis]
        np.exp(tmp, out=act)
        act /= act.sum(axis=1)[:, np.newaxis]
        return act
    if weights.out_activation == "logistic":
        act = 1.0 / (1.0 + np.exp(-act))
        if act.shape[1] == 1:
            act = act.ravel()
            return np.vstack([1 - act, act]).T
        return act
    raise ValueError(f"unsupported out_activation {weights.out_activation}")


def mlp_predict(weights: MlpWeights, X):
    """MLPClassifier._predict: LabelBinarizer.inverse_transform of the forward pass."""
    proba = mlp_predict_proba(weights, X)
    if weights.out_activation == "logistic" and len(weights.classes) == 2:
        idx = (proba[:, 1] > 0.5).astype(int)
    else:
        idx = proba.argmax(axis=1)
    return [weights.classes[i] for i in idx]


# ----------------------------------------------------------------------------
# fast_path.py post-processing
# ----------------------------------------------------------------------------
@dataclass(frozen=True)
class FramePrediction:
    start_seconds: float
    end_seconds: float
    emotion: str
    confidence: float
    probabilities: dict | None


@dataclass(frozen=True)
class SegmentPrediction:
    emotion: str
    start_seconds: float
    end_seconds: float
    confidence: float
    probabilities: dict | None = None


def aggregate_probabilities(probabilities):
    """fast_path.py:78-96."""
    if not probabilities or any(item is None for item in probabilities):
        return None
    labels = list(probabilities[0].keys())
    if any(set(item.keys()) != set(labels) for item in probabilities[1:]):
        return None
    return {label: float(fmean([item[label] for item in probabilities])) for label in labels}


def segment_predictions(frames):
    """fast_path.py:99-144: run-length merge of equal adjacent frame labels."""
    if not frames:
        return []
    segments = []
    emotion, start, end = frames[0].emotion, frames[0].start_seconds, frames[0].end_seconds
    confs, probs = [frames[0].confidence], [frames[0].probabilities]

    def _flush():
        segments.append(SegmentPrediction(emotion, start, end, float(fmean(confs)), aggregate_probabilities(probs)))

    for frame in frames[1:]:
        if frame.emotion == emotion:
            end = frame.end_seconds
            confs.append(frame.confidence)
            probs.append(frame.probabilities)
            continue
        _flush()
        emotion, start, end = frame.emotion, frame.start_seconds, frame.end_seconds
        confs, probs = [frame.confidence], [frame.probabilities]
    _flush()
    return segments


def predict_frames(weights: MlpWeights, embeddings, starts, ends):
    """fast_path.py:147-226 on an already-encoded sequence: frames + merged segments."""
    X = np.asarray(embeddings, dtype=np.float64)
    labels = [str(item) for item in mlp_predict(weights, X)]
    proba = mlp_predict_proba(weights, X)
    class_labels = [str(c) for c in weights.classes]
    frames = [
        FramePrediction(
            start_seconds=float(starts[i]),
            end_seconds=float(ends[i]),
            emotion=labels[i],
            confidence=float(np.max(proba[i])),
            probabilities={class_labels[j]: float(proba[i, j]) for j in range(len(class_labels))},
        )
        for i in range(X.shape[0])
    ]
    return frames, segment_predictions(frames)


def tuning_margins(audio, sample_rate):
    """How decisive the two tuning estimates of one clip are: (margin12, margin36).

    ``estimate_tuning`` returns the left edge of the FULLEST of 100 residual bins (Appendix A.6), a
    discontinuous function of the input: when the two fullest bins hold almost the same number of
    pitches, one float32 rounding anywhere upstream moves the arg-max and with it the whole chroma
    filterbank (margin12: chroma_stft on ``|stft(y)|``, dsp.py:113-118) or the whole constant-Q basis
    (margin36: chroma_cqt on ``harmonic(y)``, dsp.py:138-143).  The margin is the count of the fullest
    bin minus the count of the runner-up; parity tests treat clips with margin <= 2 as near-ties
    (tests/test_oracle_sensitivity.py shows the oracle flips on them under 1-ulp input noise).
    """
    prepared = pad_audio_for_fft(np.asarray(audio, dtype=np.float32))
    n_fft = min(prepared.size, 2048)

    def margin(pitch, mag, bins_per_octave):
        mask = pitch > 0
        if not mask.any():
            return 1 << 30
        threshold = np.median(mag[mask])
        freqs = pitch[(mag >= threshold) & mask]
        residual = np.mod(bins_per_octave * librosa.filters.hz_to_octs(freqs), 1.0)
        residual[residual >= 0.5] -= 1.0
        counts, _ = np.histogram(residual, np.linspace(-0.5, 0.5, 101))
        top = np.sort(counts)[::-1]
        return int(top[0] - top[1])

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        magnitude = np.abs(librosa.stft(prepared, n_fft=n_fft))
        pitch, mag = librosa.core.piptrack(S=magnitude, sr=sample_rate, n_fft=n_fft)
        m12 = margin(pitch, mag, 12)
        harmonic = librosa.effects.harmonic(prepared)
        pitch, mag = librosa.core.piptrack(y=harmonic, sr=sample_rate)
        m36 = margin(pitch, mag, 36)
    return m12, m36
