"""Second routes to the oracle's restatement-only pieces (constant-Q recursion, HPSS as a whole,
piptrack): each quantity is recomputed from its published *definition* by a computation that shares
no code path with ``oracle/shim/librosa`` and is then compared with the oracle.

* constant-Q: the definition is a bank of L1-normalised Hann-windowed complex exponentials of
  length ``Q sr / f_k`` applied in the **time domain at the full sample rate**; librosa (and the
  oracle, and the CUDA path) evaluate it by FFT-domain products over a recursively decimated signal
  with 1 % sparsified rows.  The two agree to about a percent of the largest magnitude -- which pins
  the octave order and trimming, the hop / centring of every level, the decimator's gain and the
  ``sqrt(sr ratio)`` / ``length / n_fft`` / ``1 / sqrt(length)`` scalings of
  ``oracle/shim/librosa/core.py:vqt`` (ser/_internal/utils/dsp.py:140-143 is the call site).
* HPSS as a whole (dsp.py:139): ``torch.stft`` -> scipy's two median filters -> the soft-mask
  formula written out -> ``torch.istft``, all in float64.
* piptrack / estimate_tuning (dsp.py:113-118 via chroma_stft): explicit loops over columns and bins.

CPU only, seconds.
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle.shim.librosa import core, effects

torch = pytest.importorskip("torch")

C1_HZ = 440.0 * 2.0 ** ((24 - 69) / 12.0)      # MIDI note 24


# ---- constant-Q from its time-domain definition ----------------------------------------------------
def _direct_cq_magnitudes(y, sr, frames, *, bins_per_octave=36, n_bins=252, hop=512, tuning=0.0):
    """``|C[k, t]| = sqrt(L_k) |sum_n w_k[n] y[t hop + n]|`` with ``w_k`` the L1-normalised Hann-windowed
    phasor of ``L_k = Q sr / f_k`` samples centred on the frame (no FFT, no resampling, no sparsity)."""
    y = np.asarray(y, dtype=np.float64)
    r = 2.0 ** (2.0 / bins_per_octave)
    q = (r + 1.0) / (r - 1.0)                                   # filter_scale / alpha
    out = np.zeros((n_bins, len(frames)))
    for k in range(n_bins):
        f_k = C1_HZ * 2.0 ** ((k + tuning) / bins_per_octave)
        length = q * sr / f_k
        n = np.arange(np.floor(-length / 2.0), np.floor(length / 2.0))
        hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n.size) / n.size)       # periodic Hann
        w = hann * np.exp(2j * np.pi * f_k / sr * n)
        w /= np.sum(np.abs(w))
        for j, t in enumerate(frames):
            idx = t * hop + n.astype(np.int64)
            ok = (idx >= 0) & (idx < y.size)
            out[k, j] = np.sqrt(length) * np.abs(np.sum(np.conj(w[ok]) * y[idx[ok]]))
    return out


def _cq_test_signal(sr, seconds, seed):
    """Tones on and between bin centres over eight octaves, each with its own slow envelope, an onset
    in the middle and a little noise: every octave level and several frames differ from each other."""
    rng = np.random.default_rng(seed)
    t = np.arange(int(seconds * sr)) / sr
    y = np.zeros_like(t)
    for midi, detune in ((31, 0.0), (43, 0.21), (50, -0.4), (57, 0.0), (64, 0.33), (69, 0.0), (76, -0.15), (88, 0.5), (96, 0.0)):
        f = 440.0 * 2.0 ** ((midi - 69 + detune / 3.0) / 12.0)
        env = 0.6 + 0.4 * np.sin(2 * np.pi * rng.uniform(0.3, 1.1) * t + rng.uniform(0, 6))
        y += rng.uniform(0.3, 1.0) * env * np.sin(2 * np.pi * f * t + rng.uniform(0, 6))
    y[: y.size // 2] *= 0.35                                     # an onset half way
    y += 0.01 * rng.standard_normal(y.size)
    return (y / np.max(np.abs(y))).astype(np.float32)


@pytest.mark.parametrize("sr,tuning", [(16000, 0.0), (22050, 0.0), (48000, 0.0), (16000, -0.37)])
def test_cqt_recursion_agrees_with_the_time_domain_definition(sr, tuning, record_property):
    y = _cq_test_signal(sr, 3.0, seed=sr)
    C = np.abs(core.cqt(y, sr=sr, n_bins=252, bins_per_octave=36, tuning=tuning))
    n_cols = C.shape[1]
    assert n_cols == 1 + y.size // 512
    # interior frames (the longest wavelet, 34 periods of C1, spans +-1.3 s: compare the high six octaves
    # everywhere inside the clip and the low octave where its support fits or is cut by the same zero padding)
    frames = [n_cols // 2 - 9, n_cols // 2 - 1, n_cols // 2, n_cols // 2 + 2, n_cols // 2 + 11]
    D = _direct_cq_magnitudes(y, sr, frames, tuning=tuning)
    scale = D.max()
    err = np.abs(C[:, frames] - D) / scale
    record_property("max_err_of_peak", float(err.max()))
    # measured 0.3-1.2 % of the peak (the rows' 1 % sparsification and the decimators' transition bands)
    assert err.max() <= 0.02, f"constant-Q recursion vs definition: {err.max():.4f} of the peak"
    # the strongest bins agree far better than the tolerance (gain / scaling check proper)
    top = D >= 0.5 * scale
    assert np.max(np.abs(C[:, frames][top] / D[top] - 1.0)) <= 0.02
    # with the rows' sparsification switched off the recursion is closer still: 0.3-0.5 % of the peak
    # (what is left is the decimators' transition bands and the 0.01 noise floor they alias)
    C0 = np.abs(core.cqt(y, sr=sr, n_bins=252, bins_per_octave=36, tuning=tuning, sparsity=0.0))
    err0 = np.abs(C0[:, frames] - D) / scale
    record_property("max_err_of_peak_dense_rows", float(err0.max()))
    assert err0.max() <= 0.008


def test_cqt_tone_at_a_bin_centre_has_the_closed_form_magnitude():
    """``A sin(2 pi f_k t)`` gives ``|C[k]| = (A / 2) sqrt(L_k)`` at every octave level (L1-normalised
    wavelets, ``scale=True``): catches a missed ``sqrt(2)`` per decimation level or a wrong length."""
    sr = 22050
    t = np.arange(3 * sr) / sr
    for k in (40, 76, 112, 148, 184, 220, 247):                 # one bin per octave level
        f_k = C1_HZ * 2.0 ** (k / 36.0)
        y = (0.5 * np.sin(2 * np.pi * f_k * t)).astype(np.float32)
        C = np.abs(core.cqt(y, sr=sr, n_bins=252, bins_per_octave=36, tuning=0.0))
        r = 2.0 ** (2.0 / 36)
        length = (r + 1) / (r - 1) * sr / f_k
        got = C[k, C.shape[1] // 2]
        assert int(np.argmax(C[:, C.shape[1] // 2])) == k
        assert abs(got / (0.25 * np.sqrt(length)) - 1.0) <= 0.01, (k, got, 0.25 * np.sqrt(length))


# ---- HPSS end to end ----------------------------------------------------------------------------------
@pytest.mark.parametrize("n,seed", [(48000, 1), (20000, 2), (5000, 3)])
def test_harmonic_signal_matches_torch_scipy_route(n, seed):
    import scipy.ndimage

    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    y = 0.5 * np.sin(2 * np.pi * 196.0 * t) + 0.25 * np.sin(2 * np.pi * 587.3 * t + 1.0)
    clicks = rng.integers(0, n, size=12)
    y[clicks] += rng.uniform(-0.9, 0.9, size=12)                 # percussive content
    y += 0.02 * rng.standard_normal(n)
    y = (y / np.max(np.abs(y))).astype(np.float32)

    window = torch.hann_window(2048, periodic=True, dtype=torch.float64)
    X = torch.stft(torch.from_numpy(y.astype(np.float64)), n_fft=2048, hop_length=512, window=window, center=True,
                   pad_mode="constant", return_complex=True).numpy().astype(np.complex64)   # librosa keeps complex64
    S = np.abs(X)
    harm = scipy.ndimage.median_filter(S, size=(1, 31), mode="reflect")
    perc = scipy.ndimage.median_filter(S, size=(31, 1), mode="reflect")
    # softmask(harm, perc, power=2, split_zeros=True) written out
    Z = np.maximum(harm, perc).astype(np.float64)
    bad = Z < np.finfo(np.float32).tiny
    Zs = np.where(bad, 1.0, Z)
    m = (harm / Zs) ** 2
    r = (perc / Zs) ** 2
    mask = np.where(bad, 0.5, m / (m + r))
    Xh = (S * mask.astype(np.float32)) * (X / np.where(S == 0, 1.0, S))
    want = torch.istft(torch.from_numpy(Xh.astype(np.complex128)), n_fft=2048, hop_length=512, window=window,
                       center=True, length=n).numpy()
    got = effects.harmonic(y)
    assert got.shape == (n,) and got.dtype == np.float32
    # the median filter on an axis shorter than its kernel is where scipy itself is unstable
    # (test_oracle_crosscheck.py); all three sizes here have at least 10 columns and agree
    assert np.max(np.abs(got - want)) <= 2e-6 * np.max(np.abs(want))


# ---- piptrack / estimate_tuning by explicit loops ---------------------------------------------------------
def _piptrack_loops(S, sr, n_fft, fmin=150.0, fmax=4000.0, threshold=0.1):
    S = np.asarray(S)
    n_bins, n_cols = S.shape
    pitches = np.zeros((n_bins, n_cols))
    mags = np.zeros((n_bins, n_cols))
    fmax = min(fmax, sr / 2.0)
    for c in range(n_cols):
        col = S[:, c].astype(np.float64)
        ref = threshold * col.max()
        gated = np.where(col > ref, col, 0.0)
        for b in range(1, n_bins - 1):
            f_b = b * sr / n_fft
            if not (fmin <= f_b < fmax):
                continue
            if not (gated[b] > gated[b - 1] and gated[b] >= gated[b + 1]):       # util.localmax
                continue
            a = col[b + 1] + col[b - 1] - 2.0 * col[b]
            bb = (col[b + 1] - col[b - 1]) / 2.0
            shift = 0.0 if abs(bb) >= abs(a) else -bb / a
            pitches[b, c] = (b + shift) * sr / n_fft
            mags[b, c] = col[b] + 0.5 * ((col[b + 1] - col[b - 1]) / 2.0) * shift     # np.gradient interior
    return pitches, mags


@pytest.mark.parametrize("sr", [16000, 48000])
def test_piptrack_matches_explicit_loops(sr):
    rng = np.random.default_rng(sr)
    t = np.arange(int(1.5 * sr)) / sr
    y = sum(a * np.sin(2 * np.pi * f * t + p) for a, f, p in ((0.5, 233.1, 0.0), (0.3, 466.9, 1.0), (0.2, 1401.0, 2.0), (0.1, 3333.0, 0.5)))
    y = (y + 0.01 * rng.standard_normal(t.size)).astype(np.float32)
    S = np.abs(core.stft(y, n_fft=2048))
    pitches, mags = core.piptrack(S=S, sr=sr, n_fft=2048)
    want_p, want_m = _piptrack_loops(S, sr, 2048)
    assert np.array_equal(pitches > 0, want_p > 0)                 # the same peaks, column by column
    sel = want_p > 0
    assert sel.sum() > 100
    assert np.max(np.abs(pitches[sel] - want_p[sel]) / want_p[sel]) <= 2e-6     # float32 storage of the shift
    assert np.max(np.abs(mags[sel] - want_m[sel]) / want_m[sel]) <= 2e-6


@pytest.mark.parametrize("cents", [-31, 8, 44])
def test_estimate_tuning_matches_explicit_histogram(cents):
    sr = 22050
    t = np.arange(2 * sr) / sr
    y = sum(np.sin(2 * np.pi * 440.0 * 2.0 ** ((m - 69 + cents / 100.0) / 12.0) * t) / (1 + i)
            for i, m in enumerate((57, 61, 64, 69, 73))).astype(np.float32)
    S = np.abs(core.stft(y, n_fft=2048))
    p, m = _piptrack_loops(S, sr, 2048)
    sel = p > 0
    thr = np.median(m[sel])
    f = p[sel & (m >= thr)]
    resid = np.mod(12.0 * np.log2(f / (440.0 / 16.0)), 1.0)       # hz_to_octs: octaves above A440 / 16
    resid[resid >= 0.5] -= 1.0
    counts = np.zeros(100, dtype=np.int64)
    for v in resid:                                               # np.histogram over 100 bins of [-0.5, 0.5]
        counts[min(int(np.floor((v + 0.5) * 100.0)), 99)] += 1
    want = -0.5 + 0.01 * int(np.argmax(counts))
    got = core.estimate_tuning(S=S, sr=sr, n_fft=2048, bins_per_octave=12)
    assert abs(got - want) < 1e-9
    assert abs(got - cents / 100.0) <= 0.04                       # parabolic-interpolation bias of a 2048-point Hann STFT
