"""Fast-profile prediction (mirror of ser/_internal/models/fast_path.py:19-226).

Labels and probabilities come from ONE fused CUDA forward pass (the reference runs the
sklearn forward pass twice: ``predict`` then ``predict_proba``).  Confidence, the
run-length segment merge and the ``fmean`` aggregation are host logic and kept exactly as
the reference computes them so segment boundaries and labels are bit-identical.
"""

from __future__ import annotations

import logging
from collections.abc import Callable, Sequence
from statistics import fmean

import numpy as np

from . import mlp
from .feature_extractor import FeatureFrame
from .schema import FramePrediction, InferenceResult, SegmentPrediction


def aggregate_probabilities(probabilities: list[dict[str, float] | None]) -> dict[str, float] | None:
    """Per-label mean over a segment's frames (fast_path.py:78-96)."""
    if not probabilities or any(item is None for item in probabilities):
        return None
    labels = list(probabilities[0].keys())
    if any(set(item.keys()) != set(labels) for item in probabilities[1:]):
        return None
    return {label: float(fmean([item[label] for item in probabilities])) for label in labels}


def segment_predictions(frame_predictions: list[FramePrediction]) -> list[SegmentPrediction]:
    """Merges adjacent equal frame labels; start of the first frame, end of the last
    (fast_path.py:99-144; no smoothing on the fast profile)."""
    segments: list[SegmentPrediction] = []
    run: list[FramePrediction] = []

    def close_run() -> None:
        segments.append(
            SegmentPrediction(
                emotion=run[0].emotion,
                start_seconds=run[0].start_seconds,
                end_seconds=run[-1].end_seconds,
                confidence=float(fmean([f.confidence for f in run])),
                probabilities=aggregate_probabilities([f.probabilities for f in run]),
            )
        )

    for frame in frame_predictions:
        if run and frame.emotion != run[0].emotion:
            close_run()
            run = []
        run.append(frame)
    if run:
        close_run()
    return segments


def predict_frames(model, feature_matrix: np.ndarray, starts: Sequence[float], ends: Sequence[float],
                   *, device: int = 0) -> list[FramePrediction]:
    """GPU forward pass over a (frames, dim) matrix -> FramePrediction list."""
    x = np.asarray(feature_matrix, dtype=np.float64)
    with mlp.session(model, device) as (ctx, weights):     # load + forward under one lock
        if x.ndim != 2 or x.shape[1] != weights.n_in:
            raise ValueError(f"X has {x.shape[-1]} features, but the classifier expects {weights.n_in}.")
        proba, index = ctx.mlp_predict_host(x)
    labels = [weights.classes[i] for i in index]
    class_labels = [str(item) for item in weights.classes]
    return [
        FramePrediction(
            start_seconds=float(starts[i]),
            end_seconds=float(ends[i]),
            emotion=str(labels[i]),
            confidence=float(np.max(proba[i])),
            probabilities={class_labels[j]: float(proba[i, j]) for j in range(len(class_labels))},
        )
        for i in range(len(labels))
    ]


def predict_emotions_detailed_with_model(
    file: str,
    *,
    model,
    expected_feature_size: int | None,
    output_schema_version: str,
    extract_feature_frames_fn: Callable[[str], Sequence[FeatureFrame]],
    logger: logging.Logger,
    device: int = 0,
) -> InferenceResult:
    """Same contract as fast_path.py:147-226."""
    feature_frames = list(extract_feature_frames_fn(file))
    if not feature_frames:
        logger.warning("No features extracted for file %s.", file)
        return InferenceResult(schema_version=output_schema_version, segments=[], frames=[])
    vectors = [frame.features for frame in feature_frames]
    if expected_feature_size is not None:
        wrong = {v.shape[0] for v in vectors if v.shape[0] != expected_feature_size}
        if wrong:
            raise ValueError(
                "Feature vector size mismatch for loaded model. "
                f"Expected {expected_feature_size}, got {sorted(wrong)}."
            )
    feature_matrix = np.asarray(vectors, dtype=np.float64)
    frames = predict_frames(
        model, feature_matrix,
        [f.start_seconds for f in feature_frames], [f.end_seconds for f in feature_frames],
        device=device,
    )
    if len(frames) != len(feature_frames):
        raise RuntimeError(
            "Frame/prediction length mismatch. "
            f"Got {len(feature_frames)} frames and {len(frames)} predictions."
        )
    logger.debug("Emotion model prediction completed for %d frames.", len(frames))
    segments = segment_predictions(frames)
    logger.debug("Timestamp extraction completed for %d segments.", len(segments))
    return InferenceResult(schema_version=output_schema_version, segments=segments, frames=frames)
