"""Host-side tables of the tonnetz chain against the oracle (no GPU needed).

The constant-Q plan, the sparsified FFT-domain wavelet bases and the decimation filter are
computed in C++ (ser_b200/csrc/cqt_tables.cpp) and exposed through serb_debug_* entry points;
the oracle computes the same tables the way librosa 0.11.0 does (oracle/shim/librosa/core.py
vqt / __vqt_filter_fft, filters.wavelet, util.sparsify_rows).
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle.shim.librosa import core, filters
from ser_b200 import _native


def _oracle_octave(sr0: int, tuning: float, octave: int):
    bpo, n_bins = 36, 252
    fmin = filters.note_to_hz_C1() * 2.0 ** (tuning / bpo)
    ratios = 2.0 ** (np.arange(0, bpo, dtype=float) / bpo)
    freqs = np.sort(np.multiply.outer(2.0 ** np.arange(7), ratios).flatten()[:n_bins]) * fmin
    alpha = filters._relative_bandwidth(freqs=freqs)
    _, cutoff = filters.wavelet_lengths(freqs=freqs, sr=sr0, window="hann", filter_scale=1, gamma=0, alpha=alpha)
    count = core._early_downsample_count(sr0 / 2.0, cutoff, 512, 7)
    sr = sr0 / 2 ** count
    my_sr = sr / 2 ** octave
    sl = slice(-36 * (octave + 1), -36 * octave if octave else None)
    fft_basis, n_fft, _ = core._vqt_filter_fft(my_sr, freqs[sl], 1, 1, 0.01, window="hann", gamma=0,
                                               dtype=np.complex64, alpha=alpha[sl])
    fft_basis = (fft_basis * np.sqrt(sr / my_sr)).astype(np.complex64)
    lengths, _ = filters.wavelet_lengths(freqs=freqs, sr=sr, window="hann", filter_scale=1, gamma=0, alpha=alpha)
    return np.asarray(fft_basis.todense()), n_fft, 2 ** count, 1.0 / np.sqrt(lengths[sl])


@pytest.mark.parametrize("factor", [2, 4, 8])
def test_decimation_taps_match_oracle(factor):
    np.testing.assert_allclose(_native.debug_decimation_taps(factor), core._soxr_hq_decimation_filter(factor),
                               rtol=0, atol=1e-14)


@pytest.mark.parametrize("sr", [16000, 48000, 22050])
def test_cqt_plan_and_basis_match_oracle(sr):
    plan = _native.debug_cqt_plan(sr)
    assert plan["status"] == 0
    for tuning_index, octave in ((50, 0), (0, 3), (99, 6)):
        tuning = float(np.linspace(-0.5, 0.5, 101)[tuning_index])
        ref, n_fft, early, ref_scale = _oracle_octave(sr, tuning, octave)
        assert plan["early_factor"] == early and plan["n_fft"][octave] == n_fft
        assert plan["hop0"] == 512 // early
        got, scale = _native.debug_cqt_basis(sr, tuning_index, octave)
        assert got.shape == ref.shape
        # identical sparsity pattern (util.sparsify_rows, quantile 0.01), values to float32 rounding
        assert np.array_equal(got != 0, ref != 0)
        assert np.max(np.abs(got - ref)) <= 4e-7 * np.max(np.abs(ref))
        np.testing.assert_allclose(scale, ref_scale, rtol=2e-7)
        spans = [np.flatnonzero(row) for row in got]
        assert max(int(idx[-1] - idx[0] + 1) for idx in spans) <= 32   # kCqtRowCap


def test_cqt_plan_rejects_rates_below_the_top_wavelet():
    assert _native.debug_cqt_plan(8000)["status"] == 1      # librosa: wavelet basis exceeds Nyquist
    assert _native.debug_cqt_plan(96000)["early_factor"] == 4


def test_tuning_dependent_plans_are_reported_as_unsupported_not_as_bad_input():
    """~5.5 % of sample rates have a constant-Q plan that depends on the tuning estimate
    (status 2).  They must surface as UnsupportedConfigurationError (a NotImplementedError), never as
    the ValueError the runtime boundary reads as "bad audio"."""
    assert _native.debug_cqt_plan(20600)["status"] == 2
    assert _native.debug_cqt_plan(41000)["status"] == 2
    for sr in (8000 * 2, 11025 * 2, 24000, 32000, 44100, 48000, 96000):
        assert _native.debug_cqt_plan(sr)["status"] == 0, sr
    assert issubclass(_native.UnsupportedConfigurationError, NotImplementedError)
    assert not issubclass(_native.UnsupportedConfigurationError, ValueError)


def test_column_mapped_row_sets_expand_to_the_same_basis():
    """cqtc_kernel reads the rows as 16 sets over the union of their bins (csrc/cqt_tables.cpp
    cqt_set_banks); expanded back they must be the sparsified basis itself, value for value."""
    from ser_b200 import _native

    for sr in (11025, 16000, 22050, 32000, 44100, 48000, 96000):
        for tuning in (0, 37, 99):
            for octave in (0, 3, 6):
                dense, scale = _native.debug_cqt_basis(sr, tuning, octave)
                expanded = _native.debug_cqt_set_basis(sr, tuning, octave)
                assert expanded is not None, (sr, tuning, octave)
                assert np.array_equal(expanded[0], dense) and np.array_equal(expanded[1], scale)


def test_median_networks_header_is_what_the_generator_emits():
    """csrc/median_net.cuh (the sorting / merging networks of the block medians) is generated; the generator
    checks every network on all sorted 0-1 inputs (the 0-1 principle) before printing it, so equality with
    the committed header means the kernels run verified networks."""
    import subprocess
    import sys
    from pathlib import Path

    repo = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(repo / "scripts" / "gen" / "median_net.py")], capture_output=True, text=True, check=True)
    assert out.stdout == (repo / "ser_b200" / "csrc" / "median_net.cuh").read_text()
