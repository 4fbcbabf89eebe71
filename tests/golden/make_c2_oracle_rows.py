"""Oracle rows, labels and a FITTED classifier for BASELINE config c2 (and c3's whole-clip rows).

TEST INFRASTRUCTURE.  Run in any container with the repo (CPU only, ~3 minutes on 8 cores):

    python tests/golden/make_c2_oracle_rows.py [--clips 256] [--workers 8]

Writes tests/golden/c2_oracle_rows.npz:

* ``clip_index``  (256,)      which clips of the 1 440-clip c2 grid (``synth.ravdess_specs``) were sampled
* ``window_rows`` (1024, 193) float32  oracle ``encode_sequence`` rows of those clips, 3 s / 1 s windows
                              @ 48 kHz (ser/_internal/repr/handcrafted.py:65-107 over oracle/shim/librosa)
* ``clip_rows``   (64, 193)   float64  oracle whole-clip ``extract_vector`` rows of the first 64 sampled
                              clips (config c3's unit of work, handcrafted.py:124-137)
* ``labels``      (1024,)     the emotion of each window's clip (the synthetic generator's label)
* ``window_margins`` (1024, 2), ``clip_margins`` (64, 2)  ``ser_oracle.tuning_margins``: count of the fullest
                              tuning-histogram bin minus the runner-up, for chroma_stft's and chroma_cqt's
                              tuning estimate; <= 2 marks a near-tie whose arg-max one rounding can move
* ``model/*``                 Pipeline(StandardScaler, MLPClassifier(300)) with the reference's
                              hyper-parameters (ser/_internal/models/training_support.py:87-106) FITTED by
                              scikit-learn on ``window_rows`` -> ``labels``: mean, scale, w1, b1, w2, b2, classes
* ``sk_labels`` / ``sk_proba`` scikit-learn's own ``predict`` / ``predict_proba`` on ``window_rows``

The synthetic clips are regenerated bit-for-bit on the GPU box from ``synth.clip_pcm16`` (numpy
Generator, seeded), so the GPU tests and ``bench.py`` need only this file.
"""

from __future__ import annotations

import argparse
import os
import sys
import warnings
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
OUT = Path(__file__).resolve().parent / "c2_oracle_rows.npz"
SR, N_SAMPLES, N_GRID = 48000, 168000, 1440


def sampled_clip_indices(n_clips: int) -> np.ndarray:
    """Evenly spread over the grid so that every actor / emotion / intensity appears."""
    return (np.arange(n_clips, dtype=np.int64) * N_GRID) // n_clips


def _job(index: int):
    from oracle import ser_oracle
    from ser_b200 import synth

    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=1)
    except Exception:
        pass
    spec = synth.ravdess_specs(N_GRID)[index]
    audio = synth.clip_audio(spec, SR, N_SAMPLES)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rows, starts, ends = ser_oracle.encode_sequence(audio, SR)
        margins = [ser_oracle.tuning_margins(audio[a:b], SR) for a, b in ser_oracle.frame_bounds(audio.size, SR)]
    return index, rows.astype(np.float32), spec.label, np.asarray(margins, dtype=np.int64)


def _clip_job(index: int):
    from oracle import ser_oracle
    from ser_b200 import synth

    spec = synth.ravdess_specs(N_GRID)[index]
    audio = synth.clip_audio(spec, SR, N_SAMPLES)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ser_oracle.extract_feature_from_signal(audio, SR), ser_oracle.tuning_margins(audio, SR)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=256)
    ap.add_argument("--workers", type=int, default=min(8, os.cpu_count() or 1))
    args = ap.parse_args()
    import multiprocessing as mp

    from sklearn.neural_network import MLPClassifier
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    index = sampled_clip_indices(args.clips)
    with mp.get_context("fork").Pool(args.workers) as pool:
        results = pool.map(_job, index.tolist(), chunksize=2)
        clip_results = pool.map(_clip_job, index[:64].tolist(), chunksize=2)
    clip_rows = [r[0] for r in clip_results]
    clip_margins = np.asarray([r[1] for r in clip_results], dtype=np.int64)
    window_rows = np.concatenate([r[1] for r in results], axis=0)
    window_margins = np.concatenate([r[3] for r in results], axis=0)
    per_clip = results[0][1].shape[0]
    labels = np.repeat(np.asarray([r[2] for r in results]), per_clip)
    assert window_rows.shape == (args.clips * per_clip, 193)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = Pipeline([("scaler", StandardScaler()),
                          ("classifier", MLPClassifier(alpha=0.01, batch_size=256, epsilon=1e-8, hidden_layer_sizes=(300,),
                                                       learning_rate="adaptive", max_iter=500, random_state=42))])
        x = window_rows.astype(np.float64)           # fast_path.py:166-180 widens the float32 rows
        model.fit(x, labels)
    classifier, scaler = model.named_steps["classifier"], model.named_steps["scaler"]
    sk_labels, sk_proba = model.predict(x), model.predict_proba(x)
    top2 = np.sort(sk_proba, axis=1)[:, -2:]
    print(f"near-tie windows (margin <= 2): chroma {int(np.sum(window_margins[:, 0] <= 2))}, "
          f"tonnetz {int(np.sum(window_margins[:, 1] <= 2))} of {window_margins.shape[0]}")
    print(f"{window_rows.shape[0]} window rows, train accuracy {np.mean(sk_labels == labels):.4f}, "
          f"smallest top-1/top-2 margin {np.min(top2[:, 1] - top2[:, 0]):.3e}, classes {classifier.classes_.tolist()}")
    np.savez_compressed(
        OUT, clip_index=index, window_rows=window_rows, clip_rows=np.stack(clip_rows), labels=labels,
        sk_labels=sk_labels, sk_proba=sk_proba, window_margins=window_margins, clip_margins=clip_margins,
        **{"model/mean": scaler.mean_, "model/scale": scaler.scale_, "model/w1": classifier.coefs_[0],
           "model/b1": classifier.intercepts_[0], "model/w2": classifier.coefs_[1], "model/b2": classifier.intercepts_[1],
           "model/classes": np.asarray(classifier.classes_), "model/out_activation": np.asarray(classifier.out_activation_)})
    print(f"wrote {OUT} ({OUT.stat().st_size / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
