"""Host half of the fast-profile prediction (ser_b200/fast_path.py): run-length segment merge, fmean
confidence and per-label probability aggregation, checked on the CPU against what the reference's real
``fast_path.predict_emotions_detailed_with_model`` produced (tests/golden/fast_profile_golden.npz, made
by tests/golden/make_golden.py from ser/_internal/models/fast_path.py:99-226) and against hand-computed
cases for the branches that golden does not reach (fast_path.py:78-96)."""

from __future__ import annotations

from statistics import fmean

import numpy as np

from ser_b200 import fast_path
from ser_b200.schema import FramePrediction


def _golden_frames(golden):
    classes = [str(c) for c in golden["mlp/classes"].tolist()]
    proba = golden["mlp/proba"][golden["fast/frame_rows"]]
    return classes, [
        FramePrediction(start_seconds=float(golden["fast/frame_starts"][i]), end_seconds=float(golden["fast/frame_ends"][i]),
                        emotion=str(golden["fast/frame_labels"][i]), confidence=float(golden["fast/frame_conf"][i]),
                        probabilities={c: float(proba[i, j]) for j, c in enumerate(classes)})
        for i in range(len(golden["fast/frame_labels"]))
    ]


def test_segment_merge_reproduces_the_reference_run(golden):
    classes, frames = _golden_frames(golden)
    # the golden's confidences are the row maxima of sklearn's predict_proba (fast_path.py:68)
    assert np.array_equal([f.confidence for f in frames], np.max(golden["mlp/proba"][golden["fast/frame_rows"]], axis=1))
    segments = fast_path.segment_predictions(frames)
    assert [s.emotion for s in segments] == golden["fast/seg_labels"].tolist()
    assert np.array_equal([s.start_seconds for s in segments], golden["fast/seg_starts"])
    assert np.array_equal([s.end_seconds for s in segments], golden["fast/seg_ends"])
    assert np.array_equal([s.confidence for s in segments], golden["fast/seg_conf"])          # fmean: bit-identical
    # the per-frame maps here come from the golden's 64-row predict_proba call, the reference's run made its own
    # 12-row call: sklearn's BLAS blocks the two differently, one ulp on the small probabilities
    np.testing.assert_allclose([[s.probabilities[c] for c in classes] for s in segments], golden["fast/seg_proba"],
                               rtol=4e-15, atol=0)


def _frame(label, start, end, conf, probs=None):
    return FramePrediction(start_seconds=start, end_seconds=end, emotion=label, confidence=conf, probabilities=probs)


def test_segment_merge_edge_cases():
    assert fast_path.segment_predictions([]) == []
    one = fast_path.segment_predictions([_frame("calm", 0.0, 2.5, 0.75, {"calm": 0.75, "sad": 0.25})])
    assert [(s.emotion, s.start_seconds, s.end_seconds, s.confidence) for s in one] == [("calm", 0.0, 2.5, 0.75)]
    assert one[0].probabilities == {"calm": 0.75, "sad": 0.25}
    # a label that returns later opens a new segment; overlapping windows keep first start / last end
    frames = [_frame("a", 0.0, 3.0, 0.9), _frame("a", 1.0, 4.0, 0.7), _frame("b", 2.0, 5.0, 0.6),
              _frame("a", 3.0, 6.0, 0.8), _frame("a", 4.0, 6.5, 0.4), _frame("a", 5.0, 6.5, 0.3)]
    merged = fast_path.segment_predictions(frames)
    assert [(s.emotion, s.start_seconds, s.end_seconds) for s in merged] == [("a", 0.0, 4.0), ("b", 2.0, 5.0), ("a", 3.0, 6.5)]
    assert [s.confidence for s in merged] == [fmean([0.9, 0.7]), 0.6, fmean([0.8, 0.4, 0.3])]
    assert all(s.probabilities is None for s in merged)             # no per-frame maps -> none per segment


def test_probability_aggregation_rules():
    agg = fast_path.aggregate_probabilities
    assert agg([]) is None
    assert agg([{"x": 0.2, "y": 0.8}, None]) is None                            # one frame without a map
    assert agg([{"x": 0.2, "y": 0.8}, {"x": 0.5, "z": 0.5}]) is None            # label sets differ
    out = agg([{"x": 0.2, "y": 0.8}, {"y": 0.6, "x": 0.4}, {"x": 0.9, "y": 0.1}])
    assert list(out) == ["x", "y"]                                             # the first frame's label order
    assert out == {"x": fmean([0.2, 0.4, 0.9]), "y": fmean([0.8, 0.6, 0.1])}
    mixed = fast_path.segment_predictions([_frame("a", 0.0, 1.0, 0.5, {"a": 0.5, "b": 0.5}), _frame("a", 1.0, 2.0, 0.7, None)])
    assert mixed[0].probabilities is None and mixed[0].confidence == fmean([0.5, 0.7])
