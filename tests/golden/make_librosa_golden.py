"""Drop-in recipe: pin this repo against the REAL librosa 0.11.0 arithmetic.

librosa 0.11.0 / soxr 1.0.0 / soundfile cannot be installed in the build image or on the GPU box
(no wheels, no network: pyproject.toml:30,38, uv.lock:867-868,2175-2176), so every fixture under
tests/golden/ is produced over the oracle's restatement (oracle/shim).  This script is the other
half: run it on ANY machine that has the reference's pinned third-party stack

    pip install "librosa==0.11.0" "soxr==1.0.0" numpy scipy
    python tests/golden/make_librosa_golden.py            # writes tests/golden/librosa_golden.npz

and commit the resulting file.  tests/test_librosa_golden.py then activates by itself:

* `-m "not gpu"`: the oracle (oracle/ser_oracle.py over oracle/shim) against librosa's vectors,
  group by group and stage by stage -- that is the pin DESIGN.md section 2 lists as missing;
* `-m gpu`: the CUDA path against the same vectors (rows <= 1e-4 scaled, tuning bins identical).

It needs neither /root/reference nor this repo's CUDA library: the feature recipe below is the
reference's ser/_internal/utils/dsp.py:100-144 call sequence (same calls, same keyword arguments,
same order), the clip shapes are tests/golden/make_golden.py's twelve (regenerated from
ser_b200/synth.py, which is pure numpy), and the intermediates are the stage outputs the GPU
parity tests already compare with the oracle (tests/test_gpu_tonnetz.py): harmonic signal, both
tuning estimates, constant-Q magnitudes.
"""

from __future__ import annotations

import sys
import warnings
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))          # ser_b200.synth only (numpy); NOT oracle/shim

OUT = Path(__file__).resolve().parent / "librosa_golden.npz"
REQUIRED_LIBROSA = "0.11.0"


def clip_cases():
    """Same twelve shapes as tests/golden/make_golden.py:clip_cases (kept in step by a test)."""
    from ser_b200 import synth

    S = synth.ClipSpec
    return [
        ("c16k_3s", 16000, synth.clip_pcm16(S(0, 1, 3), 16000, 48000)),
        ("c48k_3p5s", 48000, synth.clip_pcm16(S(1, 7, 5, 2, 1, 2), 48000, 168000)),
        ("c22k_2s", 22050, synth.clip_pcm16(S(2, 12, 8, 1, 2, 1), 22050, 44100)),
        ("c44k_1s", 44100, synth.clip_pcm16(S(3, 20, 2, 2, 2, 2), 44100, 44100)),
        ("c16k_tail_5937", 16000, synth.clip_pcm16(S(4, 3, 6), 16000, 5937)),
        ("c16k_2048", 16000, synth.clip_pcm16(S(5, 4, 4), 16000, 2048)),
        ("c16k_short_1500", 16000, synth.clip_pcm16(S(6, 5, 7), 16000, 1500)),
        ("c16k_short_1001", 16000, synth.clip_pcm16(S(7, 6, 1), 16000, 1001)),
        ("c16k_short_300", 16000, synth.clip_pcm16(S(8, 8, 2), 16000, 300)),
        ("c48k_short_512", 48000, synth.clip_pcm16(S(9, 9, 3), 48000, 512)),
        ("sine16k_1p5s", 16000, synth.pure_sine_pcm16(16000, 1.5, 180.0 + 22.0 * 3 + 7.0 * 2)),
        ("silence16k", 16000, np.zeros(20000, dtype=np.int16)),
    ]


def fast_profile_features(librosa, audio: np.ndarray, sr: int) -> dict[str, np.ndarray]:
    """The reference's 193-d recipe (ser/_internal/utils/dsp.py:93-144) with its intermediates."""
    audio = np.asarray(audio, dtype=np.float32)
    if audio.size < 512:                                        # _pad_audio_for_fft, dsp.py:38-45
        audio = np.pad(audio, (0, 512 - audio.size), mode="constant")
    n_fft = min(audio.size, 2048)                               # dsp.py:96
    stft = np.abs(librosa.stft(audio, n_fft=n_fft))             # dsp.py:100
    power_db = librosa.power_to_db(np.square(stft), ref=np.max)  # dsp.py:101-104
    mfcc = np.mean(librosa.feature.mfcc(y=audio, sr=sr, n_mfcc=40, n_fft=n_fft), axis=1)       # :106-111
    chroma = np.mean(librosa.feature.chroma_stft(S=stft, sr=sr, n_fft=n_fft), axis=1)         # :113-118
    mel = np.mean(librosa.feature.melspectrogram(y=audio, sr=sr, n_fft=n_fft), axis=1)        # :121-125
    # :127-136 -- the dB spectrogram goes in as S (constant zeros, or ParameterError
    # when 6400 Hz >= sr / 2); recorded as librosa really answers
    try:
        contrast = np.mean(librosa.feature.spectral_contrast(S=power_db, sr=sr, n_fft=n_fft), axis=1)
        contrast_error = ""
    except Exception as err:  # noqa: BLE001 - the text is the fixture
        contrast = np.full(7, np.nan)
        contrast_error = f"{type(err).__name__}: {err}"
    harmonic = librosa.effects.harmonic(audio)                  # :139
    tonnetz = np.mean(librosa.feature.tonnetz(y=harmonic, sr=sr), axis=1)                     # :140-143
    out = {
        "mfcc": np.asarray(mfcc, dtype=np.float64), "chroma": np.asarray(chroma, dtype=np.float64),
        "mel": np.asarray(mel, dtype=np.float64), "contrast": np.asarray(contrast, dtype=np.float64),
        "contrast_error": np.asarray(contrast_error), "tonnetz": np.asarray(tonnetz, dtype=np.float64),
        "features": np.concatenate([mfcc, chroma, mel, contrast, tonnetz]).astype(np.float64),
        "harmonic": np.asarray(harmonic, dtype=np.float32),
        # what chroma_stft and chroma_cqt estimate internally (librosa/feature/spectral.py)
        "tuning_stft": np.asarray(librosa.estimate_tuning(S=stft, sr=sr, bins_per_octave=12)),
        "tuning_cqt": np.asarray(librosa.estimate_tuning(y=harmonic, sr=sr, bins_per_octave=36)),
    }
    cq = np.abs(librosa.cqt(harmonic, sr=sr, hop_length=512, fmin=None, n_bins=252, bins_per_octave=36,
                            tuning=float(out["tuning_cqt"])))
    out["cqt_mag"] = cq.astype(np.float32)
    out["chroma_cqt"] = np.asarray(librosa.feature.chroma_cqt(y=harmonic, sr=sr), dtype=np.float32)
    # the decimator alone: what the constant-Q recursion feeds its lower octaves with
    out["resample_half"] = np.asarray(
        librosa.resample(harmonic, orig_sr=2, target_sr=1, res_type="soxr_hq", scale=True), dtype=np.float32)
    return out


def main() -> int:
    try:
        import librosa
    except ImportError:
        print("librosa is not installed here; run this where librosa==0.11.0 and soxr==1.0.0 are", file=sys.stderr)
        return 2
    if getattr(librosa, "__file__", "").startswith(str(REPO)):
        print("refusing to run over the oracle shim: this recipe needs the real librosa", file=sys.stderr)
        return 2
    if librosa.__version__ != REQUIRED_LIBROSA:
        print(f"warning: librosa {librosa.__version__}, the reference pins {REQUIRED_LIBROSA}", file=sys.stderr)
    from ser_b200 import synth

    payload: dict[str, np.ndarray] = {"librosa_version": np.asarray(librosa.__version__)}
    try:
        import soxr
        payload["soxr_version"] = np.asarray(soxr.__version__)
    except ImportError:
        payload["soxr_version"] = np.asarray("absent")
    names = []
    warnings.simplefilter("ignore")
    for name, sr, pcm in clip_cases():
        audio = synth.decode_pcm16(pcm)         # == the reference's _prepare_audio_buffer (make_golden.py:72)
        payload[f"{name}/pcm"] = pcm
        payload[f"{name}/sr"] = np.asarray(sr)
        for key, value in fast_profile_features(librosa, audio, sr).items():
            payload[f"{name}/{key}"] = value
        names.append(name)
        print(f"{name:>18s} sr={sr:6d} n={pcm.size:7d} ok")
    payload["names"] = np.asarray(names)
    np.savez_compressed(OUT, **payload)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
