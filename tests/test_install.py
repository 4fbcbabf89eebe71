"""`ser_b200.install` patches an importable reference checkout at the attributes the reference
resolves at call time.  Needs /root/reference (build container only); the
reference's `import librosa` is satisfied by the oracle shim, no arithmetic runs here."""

from __future__ import annotations

import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
REFERENCE = Path("/root/reference")

pytestmark = pytest.mark.skipif(not (REFERENCE / "ser").exists(), reason="reference checkout not present")


@pytest.fixture()
def reference_on_path(monkeypatch):
    monkeypatch.syspath_prepend(str(REFERENCE))
    monkeypatch.syspath_prepend(str(REPO / "oracle" / "shim"))
    yield
    for name in [m for m in sys.modules if m == "ser" or m.startswith("ser.") or m in ("librosa", "soundfile", "colored")
                 or m.startswith("librosa.")]:
        sys.modules.pop(name, None)


def test_install_swaps_and_uninstall_restores(reference_on_path):
    import importlib

    from ser_b200 import install

    ref_dsp = importlib.import_module("ser._internal.utils.dsp")
    ref_loader = importlib.import_module("ser._internal.data.data_loader")
    ref_fast_path = importlib.import_module("ser._internal.models.fast_path")
    ref_handcrafted = importlib.import_module("ser._internal.repr.handcrafted")
    originals = (ref_dsp.extract_feature_from_signal, ref_loader.load_checked_fast_data,
                 ref_fast_path.predict_emotions_detailed_with_model, ref_handcrafted.HandcraftedBackend.encode_sequence)
    patched = install.install(device=0)
    try:
        assert any(name.endswith("dsp.extract_feature_from_signal") for name in patched)
        assert any(name.endswith("load_checked_fast_data") for name in patched)
        assert ref_dsp.extract_feature_from_signal is not originals[0]
        assert ref_loader.load_checked_fast_data is not originals[1]
        assert ref_fast_path.predict_emotions_detailed_with_model is not originals[2]
        assert ref_handcrafted.HandcraftedBackend.encode_sequence is not originals[3]
        # the runtime hook still resolves the reference's own run_fast_inference by module path
        hooks = importlib.import_module("ser._internal.runtime.fast_inference")
        assert callable(hooks.run_fast_inference)
        # validation errors keep the reference's texts without touching the GPU
        import numpy as np
        with pytest.raises(ValueError, match="Sample rate must be a positive integer."):
            ref_dsp.extract_feature_from_signal(np.zeros(10, dtype=np.float32), 0)
        assert ref_loader.load_checked_fast_data(utterances=[], settings=None) is None
    finally:
        install.uninstall()
    assert ref_dsp.extract_feature_from_signal is originals[0]
    assert ref_loader.load_checked_fast_data is originals[1]
    assert ref_handcrafted.HandcraftedBackend.encode_sequence is originals[3]
