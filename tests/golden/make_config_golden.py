"""Oracle rows for BASELINE configs c1, c4 and c5 at their real sizes (c2 / c3: make_c2_oracle_rows.py).

TEST INFRASTRUCTURE.  Run in the build container (CPU only, a few minutes on 8 cores):

    python tests/golden/make_config_golden.py [--workers 8]

Writes tests/golden/config_golden.npz and copies the reference's bundled recording to
tests/golden/sample.wav (a data file, 140 KB; config c1 names it).

* c1  ``ser.api.infer`` on sample.wav: the five window rows (float32, what encode_sequence returns)
      the oracle computes from the file exactly as the reference reads it (PCM16 / 32768, peak
      normalise), window bounds, and -- with the fitted c2 classifier -- labels and segments.
* c4  1-hour 16 kHz recording (``synth.long_recording(16000, 57_600_000)``): oracle rows of 64 sampled
      windows, and ALL 300 windows of the first five minutes with labels and merged segments.
Every row comes with ``ser_oracle.tuning_margins`` (``*/margins``, ``*_margins``): the lead of the
fullest tuning-histogram bin over the runner-up for the two tuning estimates; <= 2 marks a near-tie.

* c5  clip length {1, 2, 3.5, 5, 10, 30, 60} s @ 48 kHz: whole-clip oracle rows of three base clips
      (positions 0, 7, 63 of the 64-clip base set the sweep tiles into every batch size).
"""

from __future__ import annotations

import argparse
import os
import shutil
import sys
import warnings
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
HERE = Path(__file__).resolve().parent
C4_SR, C4_SAMPLES = 16000, 57_600_000
C5_SR = 48000
C5_SECONDS = (1, 2, 3.5, 5, 10, 30, 60)
C5_BASE_CLIPS = (0, 7, 63)


def _limit_threads():
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=1)
    except Exception:
        pass


def _window_row(args):
    audio, sr = args
    from oracle import ser_oracle

    _limit_threads()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ser_oracle.extract_feature_from_signal(audio, sr).astype(np.float32), ser_oracle.tuning_margins(audio, sr)


def _clip_row(args):
    audio, sr = args
    from oracle import ser_oracle

    _limit_threads()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ser_oracle.extract_feature_from_signal(audio, sr), ser_oracle.tuning_margins(audio, sr)


def c5_base_clip(position: int, n_samples: int) -> np.ndarray:
    """Base clip ``position`` of the c5 sweep: the first n samples of a 60 s member of the family."""
    from ser_b200 import synth

    spec = synth.ravdess_specs(64)[position]
    return synth.clip_audio(spec, C5_SR, 60 * C5_SR)[:n_samples]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=min(8, os.cpu_count() or 1))
    args = ap.parse_args()
    import multiprocessing as mp

    from oracle import ser_oracle
    from oracle.shim import librosa
    from ser_b200 import synth

    payload: dict[str, np.ndarray] = {}
    with np.load(HERE / "c2_oracle_rows.npz", allow_pickle=False) as data:
        model = {k.split("/", 1)[1]: data[k] for k in data.files if k.startswith("model/")}
    weights = ser_oracle.MlpWeights(mean=model["mean"], scale=model["scale"], coefs=(model["w1"], model["w2"]),
                                    intercepts=(model["b1"], model["b2"]), classes=tuple(model["classes"].tolist()),
                                    out_activation=str(model["out_activation"]))

    def predict(rows, starts, ends, prefix):
        frames, segments = ser_oracle.predict_frames(weights, rows, starts, ends)
        payload[f"{prefix}/labels"] = np.asarray([f.emotion for f in frames])
        payload[f"{prefix}/confidence"] = np.asarray([f.confidence for f in frames])
        payload[f"{prefix}/seg_labels"] = np.asarray([s.emotion for s in segments])
        payload[f"{prefix}/seg_starts"] = np.asarray([s.start_seconds for s in segments])
        payload[f"{prefix}/seg_ends"] = np.asarray([s.end_seconds for s in segments])
        payload[f"{prefix}/seg_confidence"] = np.asarray([s.confidence for s in segments])

    def split(results):
        return np.stack([r[0] for r in results]), np.asarray([r[1] for r in results], dtype=np.int64)

    with mp.get_context("fork").Pool(args.workers) as pool:
        # ---- c1: the bundled recording
        sample = Path("/root/reference/sample.wav")
        shutil.copyfile(sample, HERE / "sample.wav")
        raw, sr = librosa.load(str(sample), sr=None)
        audio = ser_oracle.prepare_audio_buffer(raw)
        bounds = ser_oracle.frame_bounds(audio.size, sr)
        rows, payload["c1/margins"] = split(pool.map(_window_row, [(audio[a:b], sr) for a, b in bounds]))
        starts = np.asarray([a for a, _ in bounds], dtype=np.float64) / float(sr)
        ends = np.asarray([b for _, b in bounds], dtype=np.float64) / float(sr)
        payload.update({"c1/sr": np.asarray(sr), "c1/n_samples": np.asarray(audio.size), "c1/rows": rows,
                        "c1/starts": starts, "c1/ends": ends})
        predict(rows, starts, ends, "c1")
        print("c1", rows.shape, payload["c1/labels"].tolist(), payload["c1/seg_labels"].tolist())

        # ---- c4: one hour at 16 kHz
        recording = synth.long_recording(C4_SR, C4_SAMPLES)
        bounds = ser_oracle.frame_bounds(recording.size, C4_SR)
        assert len(bounds) == 3600
        sampled = np.unique(np.concatenate([np.linspace(0, 3599, 60).astype(np.int64), [3597, 3598, 3599, 1]]))[:64]
        rows, margins = split(pool.map(_window_row, [(recording[bounds[i][0]:bounds[i][1]], C4_SR) for i in sampled]))
        payload.update({"c4/sampled_windows": sampled, "c4/sampled_rows": rows, "c4/sampled_margins": margins})
        first = 300
        rows5, payload["c4/first5min_margins"] = split(pool.map(_window_row, [(recording[a:b], C4_SR) for a, b in bounds[:first]], chunksize=4))
        starts = np.asarray([a for a, _ in bounds[:first]], dtype=np.float64) / float(C4_SR)
        ends = np.asarray([b for _, b in bounds[:first]], dtype=np.float64) / float(C4_SR)
        payload.update({"c4/first5min_rows": rows5, "c4/first5min_starts": starts, "c4/first5min_ends": ends})
        predict(rows5, starts, ends, "c4/first5min")
        print("c4", rows.shape, rows5.shape, len(payload["c4/first5min/seg_labels"]), "segments in the first five minutes")
        del recording

        # ---- c5: whole-clip rows over the length axis
        jobs, keys = [], []
        for seconds in C5_SECONDS:
            n = int(seconds * C5_SR)
            for position in C5_BASE_CLIPS:
                jobs.append((c5_base_clip(position, n), C5_SR))
                keys.append((seconds, position))
        rows, payload["c5/margins"] = split(pool.map(_clip_row, jobs))
        payload.update({"c5/seconds": np.asarray([k[0] for k in keys], dtype=np.float64),
                        "c5/position": np.asarray([k[1] for k in keys], dtype=np.int64), "c5/rows": rows})
        print("c5", rows.shape)
    np.savez_compressed(HERE / "config_golden.npz", **payload)
    print("wrote", HERE / "config_golden.npz")


if __name__ == "__main__":
    main()
