"""Restatement of librosa.effects.harmonic / librosa.decompose.hpss (librosa 0.11.0).

TEST INFRASTRUCTURE (oracle).  Reference call site: ser/_internal/utils/dsp.py:139.
"""

from __future__ import annotations

import numpy as np
from scipy.ndimage import median_filter

from . import core, util
from .util import ParameterError


def _median_filter_reflect(S, width, *, axis):
    """scipy.ndimage.median_filter(S, size=<width along axis>, mode="reflect").

    librosa calls scipy directly.  When the axis is much shorter than the kernel scipy 1.18.1's
    1-D rank filter is unreliable in this image: for 2-column (sometimes 3-column) spectrograms it
    returns run-to-run different values, NaN included (about 1 trial in 17 in
    tests/test_oracle_crosscheck.py::test_scipy_short_axis_median_instability_...; stable and
    equal to this restatement otherwise), so axes shorter than the kernel are restated explicitly
    with scipy's documented "reflect" extension (d c b a | a b c d | d c b a, period 2n).  For
    axes at least as long as the kernel the two are bit-identical (same test file) and scipy
    itself is used.
    """
    n = S.shape[axis]
    if n >= width:
        size = [1] * S.ndim
        size[axis] = width
        return median_filter(S, size=size, mode="reflect")
    half = width // 2
    pad = [(0, 0)] * S.ndim
    pad[axis] = (half, width - 1 - half)
    padded = np.pad(S, pad, mode="symmetric")
    windows = np.lib.stride_tricks.sliding_window_view(padded, width, axis=axis)
    return np.median(windows, axis=-1).astype(S.dtype)


def hpss(S, *, kernel_size=31, power=2.0, mask=False, margin=1.0):
    """librosa.decompose.hpss: median-filter harmonic/percussive separation with soft masks."""
    if np.iscomplexobj(S):
        S, phase = core.magphase(S)
    else:
        phase = 1
    if isinstance(kernel_size, (tuple, list)):
        win_harm, win_perc = kernel_size[0], kernel_size[1]
    else:
        win_harm = win_perc = kernel_size
    if isinstance(margin, (tuple, list)):
        margin_harm, margin_perc = margin[0], margin[1]
    else:
        margin_harm = margin_perc = margin
    if margin_harm < 1 or margin_perc < 1:
        raise ParameterError("Margins must be >= 1.0. A typical range is between 1 and 10.")
    harm_shape = [1] * S.ndim
    harm_shape[-1] = win_harm
    perc_shape = [1] * S.ndim
    perc_shape[-2] = win_perc
    harm = np.empty_like(S)
    harm[:] = _median_filter_reflect(S, win_harm, axis=-1)
    perc = np.empty_like(S)
    perc[:] = _median_filter_reflect(S, win_perc, axis=-2)
    split_zeros = margin_harm == 1 and margin_perc == 1
    mask_harm = util.softmask(harm, perc * margin_harm, power=power, split_zeros=split_zeros)
    mask_perc = util.softmask(perc, harm * margin_perc, power=power, split_zeros=split_zeros)
    if mask:
        return mask_harm, mask_perc
    return ((S * mask_harm) * phase, (S * mask_perc) * phase)


def harmonic(y, *, kernel_size=31, power=2.0, mask=False, margin=1.0, n_fft=2048, hop_length=None,
             win_length=None, window="hann", center=True, pad_mode="constant"):
    """librosa.effects.harmonic: istft(hpss(stft(y))[0])."""
    stft = core.stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                     center=center, pad_mode=pad_mode)
    stft_harm = hpss(stft, kernel_size=kernel_size, power=power, mask=mask, margin=margin)[0]
    y_harm = core.istft(stft_harm, dtype=y.dtype, length=y.shape[-1], hop_length=hop_length,
                        win_length=win_length, n_fft=n_fft, window=window, center=center)
    return y_harm
