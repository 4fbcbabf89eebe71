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


def test_every_context_entry_rejects_a_null_context():
    """An entry handed a NULL context reports SERB_ERR_INVALID_ARG before it reads any other
    argument (no device, no crash): what a binding sees when serb_ctx_create failed and its
    result was used anyway."""
    import ctypes

    lib = _native.load_library()
    checked = 0
    for name, (restype, argtypes) in _native.SIGNATURES.items():
        if not argtypes or argtypes[0] is not _native._P or name in ("serb_ctx_destroy", "serb_last_error"):
            continue
        args = [None] + [t() if not hasattr(t, "contents") and t is not _native._P else None for t in argtypes[1:]]
        result = getattr(lib, name)(*args)
        if restype is ctypes.c_int:
            expected = 0 if name == "serb_mlp_n_classes" else -1   # SERB_ERR_INVALID_ARG
            assert result == expected, f"{name}(NULL, ...) returned {result}"
        checked += 1
    assert checked >= 20
    lib.serb_ctx_destroy(None)                                   # a no-op, as free(NULL)
    assert isinstance(lib.serb_last_error(None), bytes)          # the creation error text


def test_ctx_create_without_out_pointer_is_an_argument_error():
    lib = _native.load_library()
    assert lib.serb_ctx_create(0, None) == -1
    assert b"out_ctx" in lib.serb_last_error(None)
