"""Shared pytest plumbing: the ``gpu`` marker, repo on sys.path, golden fixtures, tolerance."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GROUPS = {"mfcc": (0, 40), "chroma": (40, 52), "mel": (52, 180), "contrast": (180, 187), "tonnetz": (187, 193)}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """tests/golden/fast_profile_golden.npz: outputs of the reference's host code over the oracle shim."""
    with np.load(REPO / "tests" / "golden" / "fast_profile_golden.npz", allow_pickle=False) as data:
        return {key: data[key] for key in data.files}


@pytest.fixture(scope="session")
def gpu_ctx():
    from ser_b200 import _native

    return _native.get_context(0)


def group_errors(actual: np.ndarray, expected: np.ndarray, groups=("mfcc", "chroma", "mel", "contrast")):
    """Per-group (scaled error, raw max relative error).

    Scaled error is |a - b| / max(|b|, 1e-3 * ||group||_inf): the north star's "max rel err
    <= 1e-4 on pooled features", with the floor SURVEY.md section 7 recommends so that
    near-zero coefficients (high-order MFCC means) are not held to a relative bound the
    reference's own float32 arithmetic does not meet.
    """
    report = {}
    for name in groups:
        lo, hi = GROUPS[name]
        a = np.asarray(actual[..., lo:hi], dtype=np.float64)
        b = np.asarray(expected[..., lo:hi], dtype=np.float64)
        scale = np.max(np.abs(b), axis=-1, keepdims=True)
        floor = np.maximum(np.abs(b), 1e-3 * scale)
        floor = np.where(floor == 0.0, 1.0, floor)
        scaled = float(np.max(np.abs(a - b) / floor)) if a.size else 0.0
        denom = np.where(np.abs(b) > 0, np.abs(b), 1.0)
        raw = float(np.max(np.abs(a - b) / denom)) if a.size else 0.0
        report[name] = (scaled, raw)
    return report


def binary_model_and_inputs():
    """A fitted two-class Pipeline(StandardScaler, MLPClassifier(300)): scikit-learn then uses ONE logistic
    output unit (training_support.py:87-106 builds the same pipeline whatever the label set is)."""
    from sklearn.neural_network import MLPClassifier
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    rng = np.random.default_rng(17)
    x = rng.standard_normal((300, 193)) * (1.0 + 5.0 * rng.random(193))
    y = np.where(x[:, 0] / np.std(x[:, 0]) + 0.5 * x[:, 7] / np.std(x[:, 7]) + 0.1 * rng.standard_normal(300) > 0, "happy", "sad")
    model = Pipeline([("scaler", StandardScaler()),
                      ("classifier", MLPClassifier(hidden_layer_sizes=(300,), max_iter=40, random_state=3))])
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # ConvergenceWarning: 40 iterations are plenty for a test
        model.fit(x, y)
    assert model.named_steps["classifier"].out_activation_ == "logistic"
    x_eval = rng.standard_normal((64, 193)) * (1.0 + 5.0 * rng.random(193))
    return model, x_eval
