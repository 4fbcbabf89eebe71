"""Classifier weights for the fused CUDA MLP.

The fast-profile model is ``Pipeline([("scaler", StandardScaler()), ("classifier",
MLPClassifier(hidden_layer_sizes=(300,), ...))])`` (ser/_internal/models/training_support.py:87-106),
persisted in the artifact envelope (ser/_internal/models/artifact_envelope.py:22-28).  This
module pulls the arrays the kernel needs out of a fitted model by duck typing -- scikit-learn
itself is never imported here -- and keeps one upload per device, re-done whenever the model's
weights change (``session``).
"""

from __future__ import annotations

import threading
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


_RAMPS: dict[int, np.ndarray] = {}      # position weights of the second checksum, by array length


def _fingerprint(model) -> tuple:
    """Cheap identity of a model's *current* weights: object id plus the data pointers, shapes and two
    checksums of every array the kernel reads.  A model re-fitted in place (new ``coefs_``
    arrays, or the same arrays overwritten) changes it, so stale device weights are never reused."""
    if isinstance(model, MlpWeights):
        arrays = (model.mean, model.scale, model.w1, model.b1, model.w2, model.b2)
    else:
        steps = getattr(model, "named_steps", None)
        classifier = model if steps is None else (steps.get("classifier") or list(steps.values())[-1])
        scaler = None if steps is None else steps.get("scaler")
        arrays = tuple(getattr(classifier, "coefs_", ())) + tuple(getattr(classifier, "intercepts_", ()))
        if scaler is not None:
            arrays += tuple(a for a in (getattr(scaler, "mean_", None), getattr(scaler, "scale_", None)) if a is not None)
    parts: list = [id(model)]
    for a in arrays:
        a = np.asarray(a)
        flat = a.reshape(-1)
        # two full-length sums (58 k doubles for the 193 x 300 layer: ~30 us), the second position-weighted,
        # so a single overwritten element or two swapped ones change the fingerprint too
        ramp = _RAMPS.get(flat.size)
        if ramp is None:
            ramp = _RAMPS.setdefault(flat.size, np.arange(1, flat.size + 1, dtype=np.float64))
        flat64 = flat if flat.dtype == np.float64 else flat.astype(np.float64)
        parts.append((a.__array_interface__["data"][0], a.shape, float(flat64.sum()), float(flat64 @ ramp)))
    return tuple(parts)


class _DeviceSlot:
    """The one weight slot of a device context plus the lock that makes load + forward atomic."""

    def __init__(self) -> None:
        self.lock = threading.RLock()
        self.fingerprint: tuple | None = None
        self.weights: MlpWeights | None = None


_slots: dict[int, _DeviceSlot] = {}
_slots_lock = threading.Lock()


def _slot(device: int) -> _DeviceSlot:
    with _slots_lock:
        slot = _slots.get(device)
        if slot is None:
            slot = _slots[device] = _DeviceSlot()
        return slot


class session:
    """``with mlp.session(model, device) as (ctx, weights):`` -- holds the device's weight-slot lock,
    makes sure ``model``'s current weights are the resident ones, and keeps them resident until the
    block exits.  Every forward pass of the package runs inside one, so two threads using
    different models on one device (or an abandoned timed-out call and its successor) can never
    read each other's weights; sklearn's ``predict`` is pure and so is this."""

    def __init__(self, model, device: int = 0) -> None:
        self._model = model
        self._device = int(device)
        self._slot = _slot(self._device)

    def __enter__(self):
        self._slot.lock.acquire()
        try:
            fingerprint = _fingerprint(self._model)
            ctx = _native.get_context(self._device)
            if self._slot.fingerprint != fingerprint:
                weights = self._model if isinstance(self._model, MlpWeights) else weights_from_model(self._model)
                self._slot.fingerprint = None          # a failed upload leaves no claim behind
                ctx.mlp_load(weights.mean, weights.scale, weights.w1, weights.b1, weights.w2, weights.b2,
                             weights.out_activation)
                self._slot.fingerprint, self._slot.weights = fingerprint, weights
            return ctx, self._slot.weights
        except BaseException:
            self._slot.lock.release()
            raise

    def __exit__(self, *exc) -> None:
        self._slot.lock.release()


def ensure_loaded(model, device: int = 0) -> MlpWeights:
    """Uploads ``model``'s weights to ``device`` unless they are the resident ones.  Single-threaded
    callers (bench.py) may use the context directly afterwards; concurrent callers use ``session``."""
    with session(model, device) as (_ctx, weights):
        return weights


def predict(model, feature_matrix: np.ndarray, device: int = 0) -> tuple[list, np.ndarray]:
    """(labels, probabilities): what ``model.predict`` and ``model.predict_proba`` return
    (ser/_internal/models/fast_path.py:48,181), from one fused forward pass on the GPU."""
    x = np.asarray(feature_matrix, dtype=np.float64)
    with session(model, device) as (ctx, weights):
        if x.ndim != 2 or x.shape[1] != weights.n_in:
            raise ValueError(f"X has {x.shape[-1]} features, but the classifier expects {weights.n_in}.")
        proba, index = ctx.mlp_predict_host(x)
    return [weights.classes[i] for i in index], proba
