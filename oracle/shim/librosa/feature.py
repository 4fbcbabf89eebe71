"""Restatement of librosa.feature entry points used by the reference (librosa 0.11.0).

TEST INFRASTRUCTURE (oracle).  Reference call sites in ser/_internal/utils/dsp.py:
mfcc :108, chroma_stft :115, melspectrogram :122, spectral_contrast :129, tonnetz :141.
"""

from __future__ import annotations

import numpy as np
import scipy.fftpack

from . import core, filters, util
from .util import ParameterError


def melspectrogram(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                   window="hann", center=True, pad_mode="constant", power=2.0, **kwargs):
    """librosa.feature.melspectrogram."""
    S, n_fft = core._spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, power=power,
                                 win_length=win_length, window=window, center=center, pad_mode=pad_mode)
    mel_basis = filters.mel(sr=sr, n_fft=n_fft, **kwargs)
    melspec = np.einsum("...ft,mf->...mt", S, mel_basis, optimize=True)
    return melspec


def mfcc(*, y=None, sr=22050, S=None, n_mfcc=20, dct_type=2, norm="ortho", lifter=0,
         mel_norm="slaney", **kwargs):
    """librosa.feature.mfcc."""
    if S is None:
        S = core.power_to_db(melspectrogram(y=y, sr=sr, norm=mel_norm, **kwargs))
    M = scipy.fftpack.dct(S, axis=-2, type=dct_type, norm=norm)[..., :n_mfcc, :]
    if lifter > 0:
        LI = np.sin(np.pi * np.arange(1, 1 + n_mfcc, dtype=M.dtype) / lifter)
        LI = LI.reshape((-1, 1))
        M *= 1 + (lifter / 2) * LI
        return M
    if lifter == 0:
        return M
    raise ParameterError(f"MFCC lifter={lifter} must be a non-negative number")


def chroma_stft(*, y=None, sr=22050, S=None, norm=np.inf, n_fft=2048, hop_length=512,
                win_length=None, window="hann", center=True, pad_mode="constant", tuning=None,
                n_chroma=12, **kwargs):
    """librosa.feature.chroma_stft."""
    S, n_fft = core._spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, power=2,
                                 win_length=win_length, window=window, center=center, pad_mode=pad_mode)
    if tuning is None:
        tuning = core.estimate_tuning(S=S, sr=sr, bins_per_octave=n_chroma)
    chromafb = filters.chroma(sr=sr, n_fft=n_fft, tuning=tuning, n_chroma=n_chroma, **kwargs)
    raw_chroma = np.einsum("cf,...ft->...ct", chromafb, S, optimize=True)
    return util.normalize(raw_chroma, norm=norm, axis=-2)


def chroma_cqt(*, y=None, sr=22050, C=None, hop_length=512, fmin=None, norm=np.inf, threshold=0.0,
               tuning=None, n_chroma=12, n_octaves=7, window=None, bins_per_octave=36, cqt_mode="full"):
    """librosa.feature.chroma_cqt (cqt_mode="full")."""
    if bins_per_octave is None:
        bins_per_octave = n_chroma
    elif np.remainder(bins_per_octave, n_chroma) != 0:
        raise ParameterError(f"bins_per_octave={bins_per_octave} must be an integer multiple of n_chroma={n_chroma}")
    if C is None:
        if y is None:
            raise ParameterError("At least one of C or y must be provided to compute chroma")
        C = np.abs(
            core.cqt(y, sr=sr, hop_length=hop_length, fmin=fmin, n_bins=n_octaves * bins_per_octave,
                     bins_per_octave=bins_per_octave, tuning=tuning)
        )
    cq_to_chr = filters.cq_to_chroma(C.shape[-2], bins_per_octave=bins_per_octave,
                                     n_chroma=n_chroma, fmin=fmin, window=window)
    chroma = np.einsum("cf,...ft->...ct", cq_to_chr, C, optimize=True)
    if threshold is not None:
        chroma[chroma < threshold] = 0.0
    chroma = util.normalize(chroma, norm=norm, axis=-2)
    return chroma


def tonnetz(*, y=None, sr=22050, chroma=None, **kwargs):
    """librosa.feature.tonnetz."""
    if y is None and chroma is None:
        raise ParameterError("Either the audio samples or the chromagram must be passed as an argument.")
    if chroma is None:
        chroma = chroma_cqt(y=y, sr=sr, **kwargs)
    dim_map = np.linspace(0, 12, num=chroma.shape[-2], endpoint=False)
    scale = np.asarray([7.0 / 6, 7.0 / 6, 3.0 / 2, 3.0 / 2, 2.0 / 3, 2.0 / 3])
    V = np.multiply.outer(scale, dim_map)
    V[::2] -= 0.5
    R = np.array([1, 1, 1, 1, 0.5, 0.5])
    phi = R[:, np.newaxis] * np.cos(np.pi * V)
    ton = np.einsum("pc,...ci->...pi", phi, util.normalize(chroma, norm=1, axis=-2), optimize=True)
    return ton


def spectral_contrast(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                      window="hann", center=True, pad_mode="constant", freq=None, fmin=200.0,
                      n_bands=6, quantile=0.02, linear=False):
    """librosa.feature.spectral_contrast.

    The reference feeds a dB spectrogram in [-80, 0] as ``S`` (dsp.py:127-136), so peak and
    valley are non-positive, power_to_db clamps both at amin, and the result is identically 0
    (SURVEY.md F5).  The algorithm is restated in full so that the exceptions it can raise
    (Nyquist check) and its cost are the reference's.
    """
    S, n_fft = core._spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, power=1,
                                 win_length=win_length, window=window, center=center, pad_mode=pad_mode)
    if freq is None:
        freq = filters.fft_frequencies(sr=sr, n_fft=n_fft)
    freq = np.atleast_1d(freq)
    if freq.ndim != 1 or len(freq) != S.shape[-2]:
        raise ParameterError(f"freq.shape mismatch: expected ({S.shape[-2]:d},)")
    if n_bands < 1 or not isinstance(n_bands, (int, np.integer)):
        raise ParameterError("n_bands must be a positive integer")
    if not 0.0 < quantile < 1.0:
        raise ParameterError("quantile must lie in the range (0, 1)")
    if fmin <= 0:
        raise ParameterError("fmin must be a positive number")
    octa = np.zeros(n_bands + 2)
    octa[1:] = fmin * (2.0 ** np.arange(0, n_bands + 1))
    if np.any(octa[:-1] >= 0.5 * sr):
        raise ParameterError("Frequency band exceeds Nyquist. Reduce either fmin or n_bands.")
    shape = list(S.shape)
    shape[-2] = n_bands + 1
    valley = np.zeros(shape)
    peak = np.zeros_like(valley)
    for k, (f_low, f_high) in enumerate(zip(octa[:-1], octa[1:])):
        current_band = np.logical_and(freq >= f_low, freq <= f_high)
        idx = np.flatnonzero(current_band)
        if k > 0:
            current_band[idx[0] - 1] = True
        if k == n_bands:
            current_band[idx[-1] + 1 :] = True
        sub_band = S[..., current_band, :]
        if k < n_bands:
            sub_band = sub_band[..., :-1, :]
        idx = np.rint(quantile * np.sum(current_band))
        idx = int(np.maximum(idx, 1))
        sortedr = np.sort(sub_band, axis=-2)
        valley[..., k, :] = np.mean(sortedr[..., :idx, :], axis=-2)
        peak[..., k, :] = np.mean(sortedr[..., -idx:, :], axis=-2)
    if linear:
        contrast = peak - valley
    else:
        contrast = core.power_to_db(peak) - core.power_to_db(valley)
    return contrast
