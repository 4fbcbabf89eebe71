"""Restatement of librosa.filters / librosa.core.convert pieces (librosa 0.11.0).

TEST INFRASTRUCTURE (oracle).  Reached from the reference at
ser/_internal/utils/dsp.py:106-143 through librosa.feature.{mfcc,melspectrogram,
chroma_stft,tonnetz}.
"""

from __future__ import annotations

import numpy as np
import scipy.signal

from . import util
from .util import ParameterError

# librosa.filters.WINDOW_BANDWIDTHS["hann"]
WINDOW_BANDWIDTHS = {"hann": 1.50018310546875, "ones": 1.0, "boxcar": 1.0}


# ----------------------------------------------------------------------------
# convert
# ----------------------------------------------------------------------------
def fft_frequencies(*, sr=22050, n_fft=2048):
    """librosa.fft_frequencies."""
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def hz_to_mel(frequencies, *, htk=False):
    """librosa.hz_to_mel (Slaney scale when htk=False)."""
    frequencies = np.asanyarray(frequencies)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, *, htk=False):
    """librosa.mel_to_hz."""
    mels = np.asanyarray(mels)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels=128, *, fmin=0.0, fmax=11025.0, htk=False):
    """librosa.mel_frequencies."""
    min_mel = hz_to_mel(fmin, htk=htk)
    max_mel = hz_to_mel(fmax, htk=htk)
    mels = np.linspace(min_mel, max_mel, n_mels)
    return mel_to_hz(mels, htk=htk)


def hz_to_octs(frequencies, *, tuning=0.0, bins_per_octave=12):
    """librosa.hz_to_octs: log2(f / (A440 * 2**(tuning/bpo) / 16)).  Keeps float32 for float32 input."""
    A440 = 440.0 * 2.0 ** (tuning / bins_per_octave)
    return np.log2(np.asanyarray(frequencies) / (float(A440) / 16))


def hz_to_midi(frequencies):
    """librosa.hz_to_midi."""
    return 12 * (np.log2(np.asanyarray(frequencies)) - np.log2(440.0)) + 69


def midi_to_hz(notes):
    """librosa.midi_to_hz."""
    return 440.0 * (2.0 ** ((np.asanyarray(notes) - 69.0) / 12.0))


def note_to_hz_C1():
    """librosa.note_to_hz("C1"): MIDI note 24."""
    return float(midi_to_hz(24))


# ----------------------------------------------------------------------------
# windows
# ----------------------------------------------------------------------------
def get_window(window, Nx, *, fftbins=True):
    """librosa.filters.get_window for string window specs."""
    if callable(window):
        return window(Nx)
    if isinstance(window, (str, tuple)) or np.isscalar(window):
        return scipy.signal.get_window(window, Nx, fftbins=fftbins)
    if isinstance(window, (np.ndarray, list)):
        if len(window) == Nx:
            return np.asarray(window)
        raise ParameterError(f"Window size mismatch: {len(window):d} != {Nx:d}")
    raise ParameterError(f"Invalid window specification: {window!r}")


def _float_window(window_spec):
    """librosa.filters.__float_window: windows of fractional length."""

    def _wrap(n, *args, **kwargs):
        n_min, n_max = int(np.floor(n)), int(np.ceil(n))
        window = get_window(window_spec, n_min)
        if len(window) < n_max:
            window = np.pad(window, [(0, n_max - n_min)], mode="constant")
        window[n_min:] = 0.0
        return window

    return _wrap


def window_bandwidth(window, n=1000):
    """librosa.filters.window_bandwidth for the windows used here."""
    key = window if isinstance(window, str) else getattr(window, "__name__", None)
    if key not in WINDOW_BANDWIDTHS:
        win = get_window(window, n)
        WINDOW_BANDWIDTHS[key] = n * np.sum(win**2) / (np.sum(np.abs(win)) ** 2 + util.tiny(win))
    return WINDOW_BANDWIDTHS[key]


def window_sumsquare(*, window, n_frames, hop_length=512, win_length=None, n_fft=2048,
                     dtype=np.float32, norm=None):
    """librosa.filters.window_sumsquare: float32 accumulator of float64 squared windows."""
    if win_length is None:
        win_length = n_fft
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = get_window(window, win_length)
    win_sq = util.normalize(win_sq, norm=norm) ** 2
    win_sq = util.pad_center(win_sq, size=n_fft)
    # __window_ss_fill: x[sample : sample+n_fft] += win_sq, one frame at a time
    for i in range(n_frames):
        sample = i * hop_length
        x[sample : min(n, sample + n_fft)] += win_sq[: max(0, min(n_fft, n - sample))]
    return x


# ----------------------------------------------------------------------------
# filterbanks
# ----------------------------------------------------------------------------
def mel(*, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney", dtype=np.float32):
    """librosa.filters.mel (Slaney triangles, area-normalised)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = fft_frequencies(sr=sr, n_fft=n_fft)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if isinstance(norm, str):
        if norm == "slaney":
            enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
            weights *= enorm[:, np.newaxis]
        else:
            raise ParameterError(f"Unsupported norm={norm}")
    else:
        weights = util.normalize(weights, norm=norm, axis=-1)
    return weights


def chroma(*, sr, n_fft, n_chroma=12, tuning=0.0, ctroct=5.0, octwidth=2, norm=2, base_c=True,
           dtype=np.float32):
    """librosa.filters.chroma."""
    wts = np.zeros((n_chroma, n_fft))
    frequencies = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    frqbins = n_chroma * hz_to_octs(frequencies, tuning=tuning, bins_per_octave=n_chroma)
    frqbins = np.concatenate(([frqbins[0] - 1.5 * n_chroma], frqbins))
    binwidthbins = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1]))
    D = np.subtract.outer(frqbins, np.arange(0, n_chroma, dtype="d")).T
    n_chroma2 = np.round(float(n_chroma) / 2)
    D = np.remainder(D + n_chroma2 + 10 * n_chroma, n_chroma) - n_chroma2
    wts = np.exp(-0.5 * (2 * D / np.tile(binwidthbins, (n_chroma, 1))) ** 2)
    wts = util.normalize(wts, norm=norm, axis=0)
    if octwidth is not None:
        wts *= np.tile(
            np.exp(-0.5 * (((frqbins / n_chroma - ctroct) / octwidth) ** 2)),
            (n_chroma, 1),
        )
    if base_c:
        wts = np.roll(wts, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(wts[:, : int(1 + n_fft / 2)], dtype=dtype)


def cq_to_chroma(n_input, *, bins_per_octave=12, n_chroma=12, fmin=None, window=None,
                 base_c=True, dtype=np.float32):
    """librosa.filters.cq_to_chroma."""
    n_merge = float(bins_per_octave) / n_chroma
    fmin_ = note_to_hz_C1() if fmin is None else fmin
    if np.mod(n_merge, 1) != 0:
        raise ParameterError("Incompatible CQ merge: input bins must be an integer multiple of output bins.")
    cq_to_ch = np.repeat(np.eye(n_chroma), int(n_merge), axis=1)
    cq_to_ch = np.roll(cq_to_ch, -int(n_merge // 2), axis=1)
    n_octaves = np.ceil(float(n_input) / bins_per_octave)
    cq_to_ch = np.tile(cq_to_ch, int(n_octaves))[:, :n_input]
    midi_0 = np.mod(hz_to_midi(fmin_), 12)
    roll = midi_0 if base_c else midi_0 - 9
    roll = int(np.round(roll * (n_chroma / 12.0)))
    cq_to_ch = np.roll(cq_to_ch, roll, axis=0).astype(dtype)
    if window is not None:
        cq_to_ch = scipy.signal.convolve(cq_to_ch, np.atleast_2d(window), mode="same")
    return cq_to_ch


def _relative_bandwidth(*, freqs):
    """librosa.filters._relative_bandwidth."""
    if len(freqs) <= 1:
        raise ParameterError(f"2 or more frequencies are required to compute bandwidths. Given freqs={freqs}")
    bpo = np.empty_like(freqs)
    logf = np.log2(freqs)
    bpo[0] = 1 / (logf[1] - logf[0])
    bpo[-1] = 1 / (logf[-1] - logf[-2])
    bpo[1:-1] = 2 / (logf[2:] - logf[:-2])
    alpha = (2.0 ** (2 / bpo) - 1) / (2.0 ** (2 / bpo) + 1)
    return alpha


def wavelet_lengths(*, freqs, sr=22050, window="hann", filter_scale=1, gamma=0, alpha=None):
    """librosa.filters.wavelet_lengths."""
    freqs = np.asarray(freqs)
    if filter_scale <= 0:
        raise ParameterError(f"filter_scale={filter_scale} must be positive")
    if gamma is not None and gamma < 0:
        raise ParameterError(f"gamma={gamma} must be non-negative")
    if np.any(freqs <= 0):
        raise ParameterError("frequencies must be strictly positive")
    if len(freqs) > 1 and np.any(freqs[:-1] > freqs[1:]):
        raise ParameterError(f"Frequency array={freqs} must be in strictly ascending order")
    if alpha is None:
        alpha = _relative_bandwidth(freqs=freqs)
    else:
        alpha = np.asarray(alpha)
    gamma_ = alpha * 24.7 / 0.108 if gamma is None else gamma
    Q = float(filter_scale) / alpha
    f_cutoff = max(freqs * (1 + 0.5 * window_bandwidth(window) / Q) + 0.5 * gamma_)
    lengths = Q * sr / (freqs + gamma_ / alpha)
    return lengths, f_cutoff


def wavelet(*, freqs, sr=22050, window="hann", filter_scale=1, pad_fft=True, norm=1,
            dtype=np.complex64, gamma=0, alpha=None, **kwargs):
    """librosa.filters.wavelet: time-domain constant-Q basis, rows padded to a power of two."""
    lengths, _ = wavelet_lengths(
        freqs=freqs, sr=sr, window=window, filter_scale=filter_scale, gamma=gamma, alpha=alpha
    )
    filters = []
    for ilen, freq in zip(lengths, freqs):
        sig = util.phasor(np.arange(-ilen // 2, ilen // 2, dtype=float) * 2 * np.pi * freq / sr)
        sig *= _float_window(window)(len(sig))
        sig = util.normalize(sig, norm=norm)
        filters.append(sig)
    max_len = max(lengths)
    if pad_fft:
        max_len = int(2.0 ** (np.ceil(np.log2(max_len))))
    else:
        max_len = int(np.ceil(max_len))
    filters = np.asarray(
        [util.pad_center(filt, size=max_len, **kwargs) for filt in filters], dtype=dtype
    )
    return filters, lengths
