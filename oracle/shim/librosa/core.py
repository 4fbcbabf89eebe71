"""Restatement of the librosa.core functions on the reference's fast-profile path (librosa 0.11.0).

TEST INFRASTRUCTURE (oracle).  Reference call sites: ser/_internal/utils/dsp.py:100
(stft), :101 (power_to_db), and -- through librosa.feature / librosa.effects --
piptrack / estimate_tuning / istft / cqt (dsp.py:113-118, 138-143);
ser/_internal/utils/audio_utils.py:104 (load).

dtype discipline follows upstream: float32 audio gives a float64 windowed frame, a
float64 FFT, and a complex64 result.
"""

from __future__ import annotations

import wave
import warnings

import numpy as np
import scipy.fft
import scipy.signal
import scipy.special

from . import filters, util
from .util import ParameterError


# ----------------------------------------------------------------------------
# audio I/O and resampling
# ----------------------------------------------------------------------------
def _decode_wav(path):
    """PCM WAV decode with soundfile's float32 convention (int16 / 32768)."""
    with wave.open(str(path), "rb") as handle:
        sr = handle.getframerate()
        channels = handle.getnchannels()
        width = handle.getsampwidth()
        raw = handle.readframes(handle.getnframes())
    if width == 2:
        data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / np.float32(32768.0)
    elif width == 1:
        data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / np.float32(128.0)
    elif width == 4:
        data = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        data = (v.astype(np.float64) / 8388608.0).astype(np.float32)
    else:
        raise ParameterError(f"Unsupported WAV sample width: {width}")
    if channels > 1:
        data = data.reshape(-1, channels)
    return data, int(sr)


def load(path, *, sr=22050, mono=True, offset=0.0, duration=None, dtype=np.float32, res_type="soxr_hq"):
    """librosa.load for PCM WAV at native rate (the reference always passes sr=None)."""
    data, sr_native = _decode_wav(path)
    start = int(offset * sr_native) if offset else 0
    if duration is not None:
        stop = start + int(duration * sr_native)
        data = data[start:stop]
    elif start:
        data = data[start:]
    if data.ndim > 1:
        data = data.T  # (channels, frames) like librosa
        if mono:
            data = np.mean(data, axis=0)
    if sr is not None and sr != sr_native:
        data = resample(data, orig_sr=sr_native, target_sr=sr, res_type=res_type)
    else:
        sr = sr_native
    return np.asarray(data, dtype=dtype), sr


_SOXR_FILTER_CACHE: dict[int, np.ndarray] = {}

# libsoxr's cubic fits of the Kaiser beta against attenuation (>= 60 dB), one row per octave of
# relative transition width 0.0005 * 2^row, interpolated linearly in log2(width).  Restated from
# memory of libsoxr's filter.c (the library is an un-vendored, un-installable dependency here:
# soxr 1.0.0, uv.lock:2175-2176); the digits cannot be verified offline and
# scripts/soxr_sensitivity_study.py bounds what an error in them would do (profiles/r02_soxr_sensitivity.txt).
_LSX_BETA_ROWS = (
    (-6.784957e-10, 1.02856e-05, 0.1087556, -0.8988365 + 0.001),
    (-6.897885e-10, 1.027433e-05, 0.10876, -0.8994658 + 0.002),
    (-1.000683e-09, 1.030092e-05, 0.1087677, -0.9007898 + 0.003),
    (-3.654474e-10, 1.040631e-05, 0.1087085, -0.8977766 + 0.006),
    (8.106988e-09, 6.983091e-06, 0.1091387, -0.9172048 + 0.015),
    (9.519571e-09, 7.272678e-06, 0.1090068, -0.9140768 + 0.025),
    (-5.626821e-09, 1.342186e-05, 0.1083999, -0.9065452 + 0.05),
    (-9.965946e-08, 5.073548e-05, 0.1040967, -0.7672778 + 0.085),
    (1.604808e-07, -5.856462e-05, 0.1185998, -1.34824 + 0.1),
    (-1.511964e-07, 6.363034e-05, 0.1064627, -0.9876665 + 0.18),
)


def _lsx_kaiser_beta(att, tr_bw):
    realm = np.log(tr_bw / 0.0005) / np.log(2.0)
    i0 = min(max(int(realm), 0), len(_LSX_BETA_ROWS) - 1)
    i1 = min(max(1 + int(realm), 0), len(_LSX_BETA_ROWS) - 1)
    b0, b1 = (((c[0] * att + c[1]) * att + c[2]) * att + c[3] for c in (_LSX_BETA_ROWS[i0], _LSX_BETA_ROWS[i1]))
    return float(b0 + (b1 - b0) * (realm - int(realm)))


def _soxr_hq_decimation_filter(factor):
    """Linear-phase low-pass restating libsoxr's "HQ" decimation filter (soxr 1.0.0, un-vendored).

    Quality recipe (soxr_quality_spec(SOXR_HQ)): 20-bit precision, i.e. rejection
    (20 + 1) * 20 log10(2) = 126.43 dB; pass-band end 1 - 0.05 / TO_3dB(120.41) = 0.913628 of the
    output Nyquist with TO_3dB(a) = (1.6e-6 a - 7.5e-4) a + 0.646; stop-band at the output Nyquist;
    linear phase.  Design procedure (lsx_design_lpf / lsx_kaiser_params / lsx_make_lpf): with
    frequencies normalised to the input Nyquist, tr_bw = (Fs - Fp) / 2, Fc = Fs - tr_bw, Kaiser beta
    from the library's cubic fit at relative width tr_bw / 2 / Fc, tap count
    ceil(A(beta) / tr_bw + 1) rounded up to 1 mod 4, h[i] = sin(Fc pi z) / (pi z) *
    I0(beta sqrt(1 - (z / (m/2 + 1/2))^2)) / I0(beta), z = i - m/2, no DC renormalisation.
    For the 2:1 stage of librosa.cqt that is 389 taps, beta = 13.04, Fc = 0.478407.
    The library applies it by FFT overlap-save in float32; here it is a float64 convolution.

    This is a restatement, not libsoxr: see scripts/soxr_sensitivity_study.py for what the
    uncertain digits (beta, hence the tap count) can move -- tonnetz <= 1e-4 scaled inside the
    family, no label change -- and for how far out-of-spec filters land (>= 1e-3).
    """
    if factor in _SOXR_FILTER_CACHE:
        return _SOXR_FILTER_CACHE[factor]
    bits = 20.0
    rej = bits * 20.0 * np.log10(2.0)
    att = (bits + 1.0) * 20.0 * np.log10(2.0)
    pass_end = 1.0 - 0.05 / ((1.6e-6 * rej - 7.5e-4) * rej + 0.646)
    fp, fs = pass_end / factor, 1.0 / factor           # fractions of the INPUT Nyquist
    tr_bw = min(0.5 * (fs - fp), 0.5 * fs)
    fc = fs - tr_bw
    beta = _lsx_kaiser_beta(att, tr_bw * 0.5 / fc)
    a = ((0.0007528358 - 1.577737e-05 * beta) * beta + 0.6248022) * beta + 0.06186902
    numtaps = (int(np.ceil(a / tr_bw + 1)) + 2) // 4 * 4 + 1
    m = numtaps - 1
    z = np.arange(numtaps, dtype=np.float64) - 0.5 * m
    x = z * np.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        h = np.where(x != 0, np.sin(fc * x) / x, fc)
    y = z / (0.5 * m + 0.5)
    taps = h * scipy.special.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - y * y))) / scipy.special.i0(beta)
    _SOXR_FILTER_CACHE[factor] = taps
    return taps


def _soxr_hq_decimate(y, factor):
    """Zero-latency FIR decimation: out[m] = sum_k h[k] y[factor*m + (K-1)/2 - k]."""
    taps = _soxr_hq_decimation_filter(factor)
    y64 = np.asarray(y, dtype=np.float64)
    n_out = int(np.ceil(y64.shape[-1] / factor))
    full = scipy.signal.fftconvolve(y64, taps, mode="full")
    half = (len(taps) - 1) // 2
    idx = half + factor * np.arange(n_out)
    return full[idx]


def resample(y, *, orig_sr, target_sr, res_type="soxr_hq", fix=True, scale=False, axis=-1, **kwargs):
    """librosa.resample for the integer decimations the CQT path performs (orig_sr/target_sr in {2, 4, ...})."""
    if orig_sr == target_sr:
        return y
    ratio = float(target_sr) / orig_sr
    n_samples = int(np.ceil(y.shape[axis] * ratio))
    factor = orig_sr / target_sr
    if y.ndim != 1 or abs(factor - round(factor)) > 0 or round(factor) < 2:
        raise ParameterError(
            f"oracle resample supports 1-D integer decimation only (orig_sr={orig_sr}, target_sr={target_sr})"
        )
    if not res_type.startswith("soxr"):
        raise ParameterError(f"oracle resample supports soxr_* only, got {res_type}")
    y_hat = _soxr_hq_decimate(y, int(round(factor)))
    if fix:
        y_hat = util.fix_length(y_hat, size=n_samples, **kwargs)
    if scale:
        y_hat = y_hat / np.sqrt(ratio)
    return np.asarray(y_hat, dtype=y.dtype)


# ----------------------------------------------------------------------------
# STFT / ISTFT
# ----------------------------------------------------------------------------
def stft(y, *, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
         dtype=None, pad_mode="constant", out=None):
    """librosa.stft: centred (zero-padded), periodic window, float64 FFT stored as complex64."""
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    elif not (isinstance(hop_length, (int, np.integer)) and hop_length > 0):
        raise ParameterError(f"hop_length={hop_length} must be a positive integer")
    util.valid_audio(y)
    fft_window = filters.get_window(window, win_length, fftbins=True)
    fft_window = util.pad_center(fft_window, size=n_fft)
    fft_window = fft_window.reshape((-1, 1))
    if center:
        if pad_mode in ("wrap", "maximum", "mean", "median", "minimum"):
            raise ParameterError(f"pad_mode='{pad_mode}' is not supported by librosa.stft")
        if n_fft > y.shape[-1]:
            warnings.warn(
                f"n_fft={n_fft} is too large for input signal of length={y.shape[-1]}",
                stacklevel=2,
            )
        y = np.pad(y, (n_fft // 2, n_fft // 2), mode=pad_mode)
    elif n_fft > y.shape[-1]:
        raise ParameterError(f"n_fft={n_fft} is too large for uncentered analysis of input signal of length={y.shape[-1]}")
    if dtype is None:
        dtype = util.dtype_r2c(y.dtype)
    y_frames = util.frame(y, frame_length=n_fft, hop_length=hop_length)
    n_frames = y_frames.shape[-1]
    stft_matrix = np.zeros((1 + n_fft // 2, n_frames), dtype=dtype, order="F")
    n_columns = int(util.MAX_MEM_BLOCK // (y_frames.shape[0] * y_frames.itemsize))
    n_columns = max(n_columns, 1)
    for bl_s in range(0, n_frames, n_columns):
        bl_t = min(bl_s + n_columns, n_frames)
        stft_matrix[:, bl_s:bl_t] = scipy.fft.rfft(fft_window * y_frames[:, bl_s:bl_t], axis=0)
    return stft_matrix


def _overlap_add(y, ytmp, hop_length):
    n_fft = ytmp.shape[-2]
    N = n_fft
    for frame in range(ytmp.shape[-1]):
        sample = frame * hop_length
        if N > y.shape[-1] - sample:
            N = y.shape[-1] - sample
        y[..., sample : (sample + N)] += ytmp[..., :N, frame]


def istft(stft_matrix, *, hop_length=None, win_length=None, n_fft=None, window="hann", center=True,
          dtype=None, length=None, out=None):
    """librosa.istft: windowed overlap-add divided by the window sum-of-squares."""
    if n_fft is None:
        n_fft = 2 * (stft_matrix.shape[-2] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    ifft_window = filters.get_window(window, win_length, fftbins=True)
    ifft_window = util.pad_center(ifft_window, size=n_fft).reshape((-1, 1))
    if length:
        padded_length = length + 2 * (n_fft // 2) if center else length
        n_frames = min(stft_matrix.shape[-1], int(np.ceil(padded_length / hop_length)))
    else:
        n_frames = stft_matrix.shape[-1]
    if dtype is None:
        dtype = util.dtype_c2r(stft_matrix.dtype)
    expected_signal_len = n_fft + hop_length * (n_frames - 1)
    if length:
        expected_signal_len = length
    elif center:
        expected_signal_len -= 2 * (n_fft // 2)
    y = np.zeros(expected_signal_len, dtype=dtype)

    if center:
        start_frame = int(np.ceil((n_fft // 2) / hop_length))
        ytmp = ifft_window * scipy.fft.irfft(stft_matrix[..., :start_frame], n=n_fft, axis=-2)
        head_len = n_fft + hop_length * (start_frame - 1)
        head_buffer = np.zeros(head_len, dtype=dtype)
        _overlap_add(head_buffer, ytmp, hop_length)
        if y.shape[-1] < head_len - n_fft // 2:
            y[..., :] = head_buffer[..., n_fft // 2 : y.shape[-1] + n_fft // 2]
        else:
            y[..., : head_len - n_fft // 2] = head_buffer[..., n_fft // 2 :]
        offset = start_frame * hop_length - n_fft // 2
    else:
        start_frame = 0
        offset = 0

    n_columns = int(util.MAX_MEM_BLOCK // (np.prod(stft_matrix.shape[:-1]) * stft_matrix.itemsize))
    n_columns = max(n_columns, 1)
    frame = 0
    for bl_s in range(start_frame, n_frames, n_columns):
        bl_t = min(bl_s + n_columns, n_frames)
        ytmp = ifft_window * scipy.fft.irfft(stft_matrix[..., bl_s:bl_t], n=n_fft, axis=-2)
        _overlap_add(y[..., frame * hop_length + offset :], ytmp, hop_length)
        frame += bl_t - bl_s

    ifft_window_sum = filters.window_sumsquare(
        window=window, n_frames=n_frames, win_length=win_length, n_fft=n_fft,
        hop_length=hop_length, dtype=dtype,
    )
    start = n_fft // 2 if center else 0
    ifft_window_sum = util.fix_length(ifft_window_sum[..., start:], size=y.shape[-1])
    approx_nonzero_indices = ifft_window_sum > util.tiny(ifft_window_sum)
    y[..., approx_nonzero_indices] /= ifft_window_sum[approx_nonzero_indices]
    return y


def magphase(D, *, power=1):
    """librosa.magphase."""
    mag = np.abs(D)
    zeros_to_ones = mag == 0
    mag_nonzero = mag + zeros_to_ones
    phase = np.empty_like(D, dtype=util.dtype_r2c(D.dtype))
    phase.real = D.real / mag_nonzero + zeros_to_ones
    phase.imag = D.imag / mag_nonzero
    mag **= power
    return mag, phase


def _spectrogram(*, y=None, S=None, n_fft=2048, hop_length=512, power=1, win_length=None,
                 window="hann", center=True, pad_mode="constant"):
    """librosa.core.spectrum._spectrogram."""
    if S is not None:
        if n_fft is None or n_fft // 2 + 1 != S.shape[-2]:
            n_fft = 2 * (S.shape[-2] - 1)
    else:
        if n_fft is None:
            raise ParameterError(f"Unable to compute spectrogram with n_fft={n_fft}")
        if y is None:
            raise ParameterError("Input signal must be provided to compute a spectrogram")
        S = (
            np.abs(
                stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                     center=center, window=window, pad_mode=pad_mode)
            )
            ** power
        )
    return S, n_fft


def power_to_db(S, *, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.power_to_db."""
    S = np.asarray(S)
    if amin <= 0:
        raise ParameterError("amin must be strictly positive")
    if np.issubdtype(S.dtype, np.complexfloating):
        warnings.warn("power_to_db was called on complex input", stacklevel=2)
        magnitude = np.abs(S)
    else:
        magnitude = S
    if callable(ref):
        ref_value = ref(magnitude)
    else:
        ref_value = np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        if top_db < 0:
            raise ParameterError("top_db must be non-negative")
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


# ----------------------------------------------------------------------------
# pitch: piptrack / estimate_tuning / pitch_tuning
# ----------------------------------------------------------------------------
def _parabolic_interpolation(x, *, axis=-2):
    """librosa.core.pitch._parabolic_interpolation.

    The upstream numba stencil computes, for float32 input,
    ``a = (x[1] + x[-1]) - 2 * x[0]`` and ``b = (x[1] - x[-1]) / 2`` where the
    parenthesised float32 sums round to float32 and the int-scaled terms promote to
    float64; the quotient is stored back as x's dtype.
    """
    xi = np.moveaxis(np.asarray(x), axis, -1)
    shifts = np.zeros(xi.shape, dtype=x.dtype)
    if xi.shape[-1] >= 3:
        s = (xi[..., 2:] + xi[..., :-2]).astype(np.float64)
        d = (xi[..., 2:] - xi[..., :-2]).astype(np.float64)
        a = s - 2.0 * xi[..., 1:-1].astype(np.float64)
        b = d / 2.0
        with np.errstate(divide="ignore", invalid="ignore"):
            val = np.where(np.abs(b) >= np.abs(a), 0.0, -b / a)
        shifts[..., 1:-1] = val.astype(x.dtype)
    return np.moveaxis(shifts, -1, axis)


def piptrack(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=None, fmin=150.0, fmax=4000.0,
             threshold=0.1, win_length=None, window="hann", center=True, pad_mode="constant", ref=None):
    """librosa.piptrack."""
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                            window=window, center=center, pad_mode=pad_mode)
    S = np.abs(S)
    fmin = np.maximum(fmin, 0)
    fmax = np.minimum(fmax, float(sr) / 2)
    fft_freqs = filters.fft_frequencies(sr=sr, n_fft=n_fft)
    avg = np.gradient(S, axis=-2)
    shift = _parabolic_interpolation(S, axis=-2)
    dskew = 0.5 * avg * shift
    pitches = np.zeros_like(S)
    mags = np.zeros_like(S)
    freq_mask = (fmin <= fft_freqs) & (fft_freqs < fmax)
    freq_mask = freq_mask.reshape((-1, 1))
    if ref is None:
        ref = np.max
    if callable(ref):
        ref_value = threshold * ref(S, axis=-2)
        ref_value = np.expand_dims(ref_value, -2)
    else:
        ref_value = np.abs(ref)
    idx = np.nonzero(freq_mask & util.localmax(S * (S > ref_value), axis=-2))
    pitches[idx] = (idx[-2] + shift[idx]) * float(sr) / n_fft
    mags[idx] = S[idx] + dskew[idx]
    return pitches, mags


def pitch_tuning(frequencies, *, resolution=0.01, bins_per_octave=12):
    """librosa.pitch_tuning: left edge of the fullest 0.01-wide residual bin."""
    frequencies = np.atleast_1d(frequencies)
    frequencies = frequencies[frequencies > 0]
    if not np.any(frequencies):
        warnings.warn("Trying to estimate tuning from empty frequency set.", stacklevel=2)
        return 0.0
    residual = np.mod(bins_per_octave * filters.hz_to_octs(frequencies), 1.0)
    residual[residual >= 0.5] -= 1.0
    bins = np.linspace(-0.5, 0.5, int(np.ceil(1.0 / resolution)) + 1)
    counts, tuning = np.histogram(residual, bins)
    tuning_est: float = tuning[np.argmax(counts)]
    return tuning_est


def estimate_tuning(*, y=None, sr=22050, S=None, n_fft=2048, resolution=0.01, bins_per_octave=12, **kwargs):
    """librosa.estimate_tuning."""
    pitch, mag = piptrack(y=y, sr=sr, S=S, n_fft=n_fft, **kwargs)
    pitch_mask = pitch > 0
    if pitch_mask.any():
        threshold = np.median(mag[pitch_mask])
    else:
        threshold = 0.0
    return pitch_tuning(
        pitch[(mag >= threshold) & pitch_mask], resolution=resolution, bins_per_octave=bins_per_octave
    )


# ----------------------------------------------------------------------------
# constant-Q
# ----------------------------------------------------------------------------
def _vqt_filter_fft(sr, freqs, filter_scale, norm, sparsity, hop_length=None, window="hann",
                    gamma=0.0, dtype=np.complex64, alpha=None):
    """librosa.core.constantq.__vqt_filter_fft."""
    basis, lengths = filters.wavelet(freqs=freqs, sr=sr, filter_scale=filter_scale, norm=norm,
                                     pad_fft=True, window=window, gamma=gamma, alpha=alpha)
    n_fft = basis.shape[1]
    if hop_length is not None and n_fft < 2.0 ** (1 + np.ceil(np.log2(hop_length))):
        n_fft = int(2.0 ** (1 + np.ceil(np.log2(hop_length))))
    basis *= lengths[:, np.newaxis] / float(n_fft)
    fft_basis = scipy.fft.fft(basis, n=n_fft, axis=1)[:, : (n_fft // 2) + 1]
    fft_basis = util.sparsify_rows(fft_basis, quantile=sparsity, dtype=dtype)
    return fft_basis, n_fft, lengths


def _cqt_response(y, n_fft, hop_length, fft_basis, mode, window="ones", phase=True, dtype=None):
    """librosa.core.constantq.__cqt_response."""
    D = stft(y, n_fft=n_fft, hop_length=hop_length, window=window, pad_mode=mode, dtype=dtype)
    if not phase:
        D = np.abs(D)
    return np.asarray(fft_basis.dot(D), dtype=D.dtype)


def _num_two_factors(x):
    if x <= 0:
        return 0
    num_twos = 0
    while x % 2 == 0:
        num_twos += 1
        x //= 2
    return num_twos


def _early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves):
    downsample_count1 = max(0, int(np.ceil(np.log2(nyquist / filter_cutoff)) - 1) - 1)
    num_twos = _num_two_factors(hop_length)
    downsample_count2 = max(0, num_twos - n_octaves + 1)
    return min(downsample_count1, downsample_count2)


def _early_downsample(y, sr, hop_length, res_type, n_octaves, nyquist, filter_cutoff, scale):
    downsample_count = _early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves)
    if downsample_count > 0:
        downsample_factor = 2 ** (downsample_count)
        hop_length //= downsample_factor
        if y.shape[-1] < downsample_factor:
            raise ParameterError(
                f"Input signal length={len(y):d} is too short for {n_octaves:d}-octave CQT"
            )
        new_sr = sr / float(downsample_factor)
        y = resample(y, orig_sr=downsample_factor, target_sr=1, res_type=res_type, scale=True)
        if not scale:
            y *= np.sqrt(downsample_factor)
        sr = new_sr
    return y, sr, hop_length


def _trim_stack(cqt_resp, n_bins, dtype):
    max_col = min(c_i.shape[-1] for c_i in cqt_resp)
    cqt_out = np.empty((n_bins, max_col), dtype=dtype, order="F")
    end = n_bins
    for c_i in cqt_resp:
        n_oct = c_i.shape[-2]
        if end < n_oct:
            cqt_out[:end, :] = c_i[-end:, :max_col]
        else:
            cqt_out[end - n_oct : end, :] = c_i[:, :max_col]
        end -= n_oct
    return cqt_out


def vqt(y, *, sr=22050, hop_length=512, fmin=None, n_bins=84, intervals="equal", gamma=None,
        bins_per_octave=12, tuning=0.0, filter_scale=1, norm=1, sparsity=0.01, window="hann",
        scale=True, pad_mode="constant", res_type="soxr_hq", dtype=None):
    """librosa.vqt (equal-temperament intervals only)."""
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    if fmin is None:
        fmin = filters.note_to_hz_C1()
    if tuning is None:
        tuning = estimate_tuning(y=y, sr=sr, bins_per_octave=bins_per_octave)
    if dtype is None:
        dtype = util.dtype_r2c(y.dtype)
    # hop_length must divide evenly through the octave recursion
    if _num_two_factors(hop_length) < n_octaves - 1:
        raise ParameterError(
            f"hop_length must be a positive integer multiple of 2^{n_octaves - 1:d} for {n_octaves:d}-octave CQT/VQT"
        )
    fmin = fmin * 2.0 ** (tuning / bins_per_octave)
    # interval_frequencies(intervals="equal", sort=True): one octave of ratios tiled upwards
    ratios = 2.0 ** (np.arange(0, bins_per_octave, dtype=float) / bins_per_octave)
    n_oct_tiles = np.ceil(n_bins / bins_per_octave)
    all_ratios = np.multiply.outer(2.0 ** np.arange(n_oct_tiles), ratios).flatten()[:n_bins]
    freqs = np.sort(all_ratios) * fmin
    freqs_top = freqs[-bins_per_octave:]
    fmax_t = np.max(freqs_top)
    if n_bins == 1:
        r = 2.0 ** (2.0 / bins_per_octave)
        alpha = np.atleast_1d((r - 1) / (r + 1))
    else:
        alpha = filters._relative_bandwidth(freqs=freqs)
    lengths, filter_cutoff = filters.wavelet_lengths(freqs=freqs, sr=sr, window=window,
                                                     filter_scale=filter_scale, gamma=gamma, alpha=alpha)
    nyquist = sr / 2.0
    if filter_cutoff > nyquist:
        raise ParameterError(
            f"Wavelet basis with max frequency={fmax_t} would exceed the Nyquist frequency={nyquist}. "
            "Try reducing the number of frequency bins."
        )
    y, sr, hop_length = _early_downsample(y, sr, hop_length, res_type, n_octaves, nyquist,
                                          filter_cutoff, scale)
    vqt_resp = []
    my_y, my_sr, my_hop = y, sr, hop_length
    for i in range(n_octaves):
        if i == 0:
            sl = slice(-n_filters, None)
        else:
            sl = slice(-n_filters * (i + 1), -n_filters * i)
        freqs_oct = freqs[sl]
        alpha_oct = alpha[sl]
        fft_basis, n_fft, _ = _vqt_filter_fft(my_sr, freqs_oct, filter_scale, norm, sparsity,
                                              window=window, gamma=gamma, dtype=dtype, alpha=alpha_oct)
        fft_basis = fft_basis * np.sqrt(sr / my_sr)
        fft_basis = fft_basis.astype(dtype)
        vqt_resp.append(_cqt_response(my_y, n_fft, my_hop, fft_basis, pad_mode, dtype=dtype))
        if my_hop % 2 == 0:
            my_hop //= 2
            my_sr /= 2.0
            my_y = resample(my_y, orig_sr=2, target_sr=1, res_type=res_type, scale=True)
    V = _trim_stack(vqt_resp, n_bins, dtype)
    if scale:
        lengths, _ = filters.wavelet_lengths(freqs=freqs, sr=sr, window=window,
                                             filter_scale=filter_scale, gamma=gamma, alpha=alpha)
        V /= np.sqrt(lengths).reshape((-1, 1))
    return V


def cqt(y, *, sr=22050, hop_length=512, fmin=None, n_bins=84, bins_per_octave=12, tuning=0.0,
        filter_scale=1, norm=1, sparsity=0.01, window="hann", scale=True, pad_mode="constant",
        res_type="soxr_hq", dtype=None):
    """librosa.cqt: the gamma=0 special case of vqt."""
    return vqt(y=y, sr=sr, hop_length=hop_length, fmin=fmin, n_bins=n_bins, intervals="equal",
               gamma=0, bins_per_octave=bins_per_octave, tuning=tuning, filter_scale=filter_scale,
               norm=norm, sparsity=sparsity, window=window, scale=scale, pad_mode=pad_mode,
               res_type=res_type, dtype=dtype)
