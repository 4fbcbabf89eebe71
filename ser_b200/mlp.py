"""Classifier weights for the fused CUDA MLP.

The fast-profile model is ``Pipeline([("scaler", StandardScaler()), ("classifier",
MLPClassifier(hidden_layer_sizes=(300,), ...))])`` (ser/_internal/models/training_support.py:87-106),
persisted in the artifact envelope (ser/_internal/models/artifact_envelope.py:22-28).  This
module pulls the arrays the kernel needs out of a fitted model by duck typing -- scikit-learn
itself is never imported here -- and caches one upload per (model, device).
"""

from __future__ import annotations

import threading
import weakref
from dataclasses import dataclass

import numpy as np

from . import _native


@dataclass(frozen=True)
class MlpWeights:
    mean: np.ndarray
    scale: np.ndarray
    w1: np.ndarray
    b1: np.ndarray
    w2: np.ndarray
    b2: np.ndarray
    classes: tuple
    out_activation: int

    @property
    def n_in(self) -> int:
        return int(self.w1.shape[0])


def weights_from_model(model) -> MlpWeights:
    """Extracts scaler + single-hidden-layer relu MLP weights from a fitted sklearn model."""
    scaler = None
    classifier = model
    steps = getattr(model, "named_steps", None)
    if steps is not None:
        scaler = steps.get("scaler")
        classifier = steps.get("classifier", None)
        if classifier is None:
            classifier = list(steps.values())[-1]
    coefs = getattr(classifier, "coefs_", None)
    intercepts = getattr(classifier, "intercepts_", None)
    classes = getattr(classifier, "classes_", None)
    if coefs is None or intercepts is None or classes is None:
        raise TypeError(
            "ser_b200 needs a fitted sklearn MLPClassifier (optionally behind a StandardScaler "
            "pipeline) exposing coefs_/intercepts_/classes_; there is no CPU fallback."
        )
    if len(coefs) != 2:
        raise TypeError(f"the fused MLP kernel supports one hidden layer, model has {len(coefs) - 1}")
    activation = getattr(classifier, "activation", "relu")
    if activation != "relu":
        raise TypeError(f"the fused MLP kernel supports relu hidden units, model uses {activation!r}")
    out_name = str(getattr(classifier, "out_activation_", "softmax"))
    if out_name not in ("softmax", "logistic"):
        raise TypeError(f"unsupported output activation {out_name!r}")
    n_in = int(np.asarray(coefs[0]).shape[0])
    mean = np.zeros(n_in, dtype=np.float64)
    scale = np.ones(n_in, dtype=np.float64)
    if scaler is not None:
        if getattr(scaler, "with_mean", True) and getattr(scaler, "mean_", None) is not None:
            mean = np.asarray(scaler.mean_, dtype=np.float64)
        if getattr(scaler, "with_std", True) and getattr(scaler, "scale_", None) is not None:
            scale = np.asarray(scaler.scale_, dtype=np.float64)
    return MlpWeights(
        mean=mean,
        scale=scale,
        w1=np.ascontiguousarray(coefs[0], dtype=np.float64),
        b1=np.ascontiguousarray(intercepts[0], dtype=np.float64),
        w2=np.ascontiguousarray(coefs[1], dtype=np.float64),
        b2=np.ascontiguousarray(intercepts[1], dtype=np.float64),
        classes=tuple(np.asarray(classes).tolist()),
        out_activation=_native.OUT_SOFTMAX if out_name == "softmax" else _native.OUT_LOGISTIC,
    )


_loaded: dict[int, tuple] = {}   # device -> (weakref-or-id key, MlpWeights)
_loaded_lock = threading.Lock()


def ensure_loaded(model, device: int = 0) -> MlpWeights:
    """Uploads ``model``'s weights to ``device`` unless they are the ones already resident."""
    key = id(model)
    with _loaded_lock:
        current = _loaded.get(device)
        if current is not None and current[0] == key and current[2]() is model:
            return current[1]
        weights = model if isinstance(model, MlpWeights) else weights_from_model(model)
        ctx = _native.get_context(device)
        ctx.mlp_load(weights.mean, weights.scale, weights.w1, weights.b1, weights.w2, weights.b2,
                     weights.out_activation)
        try:
            ref = weakref.ref(model)
        except TypeError:
            ref = (lambda m=model: m)
        _loaded[device] = (key, weights, ref)
        return weights


def predict(model, feature_matrix: np.ndarray, device: int = 0) -> tuple[list, np.ndarray]:
    """(labels, probabilities): what ``model.predict`` and ``model.predict_proba`` return
    (ser/_internal/models/fast_path.py:48,181), from one fused forward pass on the GPU."""
    weights = ensure_loaded(model, device)
    x = np.asarray(feature_matrix, dtype=np.float64)
    if x.ndim != 2 or x.shape[1] != weights.n_in:
        raise ValueError(f"X has {x.shape[-1]} features, but the classifier expects {weights.n_in}.")
    proba, index = _native.get_context(device).mlp_predict_host(x)
    return [weights.classes[i] for i in index], proba
