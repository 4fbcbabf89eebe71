"""Minimal librosa-API-compatible oracle module (restates librosa 0.11.0).  TEST INFRASTRUCTURE.

Exposes exactly the entry points the reference's fast-profile path touches
(SURVEY.md Appendix C): load, stft, power_to_db, feature.{mfcc, chroma_stft,
melspectrogram, spectral_contrast, tonnetz}, effects.harmonic -- plus the internals
they are built from.  See oracle/__init__.py for scope and the "parity unpinned" note.
"""

from . import core, effects, feature, filters, util
from .core import (
    cqt,
    estimate_tuning,
    istft,
    load,
    magphase,
    piptrack,
    pitch_tuning,
    power_to_db,
    resample,
    stft,
    vqt,
)
from .filters import fft_frequencies, hz_to_mel, hz_to_octs, mel_frequencies, mel_to_hz
from .util import LibrosaError, ParameterError

decompose = effects  # librosa.decompose.hpss lives beside effects.harmonic here

__version__ = "0.11.0+oracle"

__all__ = [
    "core", "effects", "feature", "filters", "util", "decompose",
    "cqt", "vqt", "estimate_tuning", "istft", "load", "magphase", "piptrack", "pitch_tuning",
    "power_to_db", "resample", "stft", "LibrosaError", "ParameterError",
]
