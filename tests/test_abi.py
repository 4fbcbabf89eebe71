"""The C-ABI library loads and exports every symbol include/ser_b200.h declares (no GPU needed)."""

from __future__ import annotations

import re
from pathlib import Path

import numpy as np
import pytest

from ser_b200 import _native

REPO = Path(__file__).resolve().parents[1]


def _declared_symbols() -> set[str]:
    header = (REPO / "include" / "ser_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return set(re.findall(r"\b(serb_[a-z0-9_]+)\s*\(", header))


def test_library_exports_every_declared_symbol():
    lib = _native.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ser_b200.h but not exported"
    assert declared == set(_native.SIGNATURES), "ctypes signatures and header disagree"


def test_version_and_feature_dim():
    lib = _native.load_library()
    assert b"sm_100a" in lib.serb_version()
    assert lib.serb_feature_dim(_native.FLAG_ALL) == 193
    assert lib.serb_feature_dim(_native.FLAG_ALL & ~_native.FLAG_TONNETZ) == 187
    assert lib.serb_feature_dim(_native.FLAG_MFCC) == 40
    assert lib.serb_feature_dim(0) == 0


def test_no_cpu_fallback_without_device():
    if _native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.Context(0)
    from ser_b200 import dsp

    with pytest.raises(RuntimeError):
        dsp.extract_feature_from_signal(np.zeros(4096, dtype=np.float32), 16000)


@pytest.mark.parametrize("sr", [16000, 22050, 44100, 48000])
def test_host_tables_match_oracle(sr):
    """The library's own mel / chroma / DCT / window tables equal the oracle's restatement of
    librosa.filters (these are computed on the host, so they can be checked without a GPU)."""
    import scipy.fftpack
    import scipy.signal

    from oracle.shim.librosa import filters

    assert np.array_equal(_native.debug_filterbank(0, sr, 2048), filters.mel(sr=sr, n_fft=2048))
    edges = np.linspace(-0.5, 0.5, 101)
    for idx in (0, 13, 50, 77, 99):
        assert np.array_equal(
            _native.debug_filterbank(1, sr, 2048, idx), filters.chroma(sr=sr, n_fft=2048, tuning=edges[idx])
        )
    dct = _native.debug_filterbank(2, sr, 2048)
    eye = np.eye(128)
    ref = scipy.fftpack.dct(eye, axis=0, type=2, norm="ortho")[:40]
    assert np.max(np.abs(dct - ref)) < 1e-7
    win = _native.debug_filterbank(3, sr, 2048)
    assert np.max(np.abs(win - scipy.signal.get_window("hann", 2048, fftbins=True))) < 1e-7
