"""Host side of the fused classifier (ser_b200/mlp.py): what is pulled out of a fitted scikit-learn model
(ser/_internal/models/training_support.py:87-106 builds Pipeline(StandardScaler, MLPClassifier(300))),
which models are refused, and the fingerprint that decides when device weights are stale.  No GPU:
the arrays are checked by recomputing sklearn's own predict_proba from them in numpy."""

from __future__ import annotations

import warnings

import numpy as np
import pytest

from ser_b200 import _native, mlp

sklearn = pytest.importorskip("sklearn")


def _fit(n_classes=4, hidden=(300,), activation="relu", scaler=True, seed=0, n_in=193):
    from sklearn.neural_network import MLPClassifier
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    rng = np.random.default_rng(seed)
    x = rng.standard_normal((40 * n_classes, n_in)) * (1.0 + 3.0 * rng.random(n_in))
    y = np.asarray([f"c{i % n_classes}" for i in range(x.shape[0])])
    x += np.asarray([np.sin(np.arange(n_in) * (1 + int(label[1:]))) for label in y])
    clf = MLPClassifier(hidden_layer_sizes=hidden, activation=activation, max_iter=15, random_state=seed)
    model = Pipeline([("scaler", StandardScaler()), ("classifier", clf)]) if scaler else clf
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(x, y)
    return model, x


def _forward(w: mlp.MlpWeights, x):
    z = (x - w.mean) / w.scale
    h = np.maximum(z @ w.w1 + w.b1, 0.0)
    o = h @ w.w2 + w.b2
    if w.out_activation == _native.OUT_LOGISTIC:
        p1 = 1.0 / (1.0 + np.exp(-o[:, 0]))
        return np.stack([1.0 - p1, p1], axis=1)
    e = np.exp(o - o.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


@pytest.mark.parametrize("n_classes,scaler", [(8, True), (3, False), (2, True)])
def test_extracted_arrays_reproduce_sklearn(n_classes, scaler):
    model, x = _fit(n_classes=n_classes, scaler=scaler, seed=n_classes)
    w = mlp.weights_from_model(model)
    assert w.n_in == 193 and w.w1.shape == (193, 300) and w.b1.shape == (300,)
    assert w.w2.shape == (300, 1 if n_classes == 2 else n_classes)
    assert w.out_activation == (_native.OUT_LOGISTIC if n_classes == 2 else _native.OUT_SOFTMAX)
    assert list(w.classes) == sorted(w.classes) and len(w.classes) == n_classes
    if not scaler:
        assert np.all(w.mean == 0.0) and np.all(w.scale == 1.0)
    np.testing.assert_allclose(_forward(w, x[:32]), model.predict_proba(x[:32]), rtol=0, atol=1e-12)
    assert [w.classes[i] for i in np.argmax(_forward(w, x[:32]), axis=1)] == model.predict(x[:32]).tolist()


def test_models_the_kernel_cannot_run_are_refused_by_name():
    with pytest.raises(TypeError, match="one hidden layer, model has 2"):
        mlp.weights_from_model(_fit(hidden=(32, 16))[0])
    with pytest.raises(TypeError, match="relu hidden units, model uses 'tanh'"):
        mlp.weights_from_model(_fit(activation="tanh")[0])
    from sklearn.neural_network import MLPClassifier

    with pytest.raises(TypeError, match="no CPU fallback"):
        mlp.weights_from_model(MLPClassifier())            # not fitted: no coefs_
    with pytest.raises(TypeError, match="no CPU fallback"):
        mlp.weights_from_model(object())


def test_fingerprint_tracks_the_current_weights():
    model, x = _fit(n_classes=4, seed=5)
    first = mlp._fingerprint(model)
    assert mlp._fingerprint(model) == first                                    # stable while nothing changes
    clf = model.named_steps["classifier"]
    clf.coefs_[0][101, 77] += 1e-3                                             # one element overwritten in place
    one = mlp._fingerprint(model)
    assert one != first
    a, b = clf.coefs_[0][5, 5], clf.coefs_[0][6, 6]
    clf.coefs_[0][5, 5], clf.coefs_[0][6, 6] = b, a                            # two elements swapped: same plain sum
    changed = mlp._fingerprint(model)
    assert changed not in (first, one)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(x, np.asarray([f"c{(i * 7) % 4}" for i in range(x.shape[0])]))   # re-fitted in place: new arrays
    assert mlp._fingerprint(model) not in (first, changed)
    other, _ = _fit(n_classes=4, seed=5)                                       # equal weights, another object
    assert mlp._fingerprint(other) != mlp._fingerprint(model)
    weights = mlp.weights_from_model(other)
    assert mlp._fingerprint(weights) == mlp._fingerprint(weights)
