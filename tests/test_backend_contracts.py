"""The wire DTOs of the representation layer reject what the reference's reject, with the same text and
in the same order (ser/_internal/repr/backend.py:19-111).  Expected texts are written out here; when the
reference package is importable (this container, or the copy ``build()`` stages under baseline/_ref)
the two implementations are also run side by side on every case."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

from ser_b200 import backend as mine

REPO = Path(__file__).resolve().parents[1]
NAN = float("nan")


def _t(*values):
    return np.array(values, dtype=np.float64)


def _rows(n, dim=3):
    return np.zeros((n, dim), dtype=np.float32)


WINDOW_CASES = [
    ((0.0, 1.0), None),
    ((0.5, 0.6), None),
    ((NAN, 1.0), "PoolingWindow bounds must be finite numbers."),
    ((0.0, float("inf")), "PoolingWindow bounds must be finite numbers."),
    ((-1.0, NAN), "PoolingWindow bounds must be finite numbers."),
    ((-1.0, 1.0), "PoolingWindow start_seconds must be non-negative."),
    ((-2.0, -3.0), "PoolingWindow start_seconds must be non-negative."),
    ((1.0, 1.0), "PoolingWindow end_seconds must be greater than start_seconds."),
    ((2.0, 1.0), "PoolingWindow end_seconds must be greater than start_seconds."),
]

_OK = dict(embeddings=_rows(2), frame_start_seconds=_t(0, 1), frame_end_seconds=_t(1, 2), backend_id="x")
SEQUENCE_CASES = [
    ({}, None),
    (dict(embeddings=_rows(1), frame_start_seconds=_t(0), frame_end_seconds=_t(1)), None),
    (dict(backend_id=""), "EncodedSequence backend_id must be a non-empty string."),
    (dict(embeddings=np.zeros(3, np.float32)), "EncodedSequence embeddings must be 2D (frames, features)."),
    (dict(frame_start_seconds=_t(0, 1).reshape(1, 2)), "Frame timestamp arrays must be 1D."),
    (dict(embeddings=_rows(0), frame_start_seconds=_t(), frame_end_seconds=_t()),
     "EncodedSequence must contain at least one frame."),
    (dict(frame_start_seconds=_t(0, 1, 2)), "frame_start_seconds length must match embeddings frame count."),
    (dict(frame_end_seconds=_t(1, 2, 3)), "frame_end_seconds length must match embeddings frame count."),
    (dict(embeddings=_rows(2) + NAN), "EncodedSequence embeddings contain non-finite values."),
    (dict(frame_start_seconds=_t(0, NAN)), "EncodedSequence frame_start_seconds contain non-finite values."),
    (dict(frame_end_seconds=_t(1, np.inf)), "EncodedSequence frame_end_seconds contain non-finite values."),
    (dict(frame_start_seconds=_t(1, 0), frame_end_seconds=_t(2, 3)), "frame_start_seconds must be non-decreasing."),
    (dict(frame_end_seconds=_t(3, 2)), "frame_end_seconds must be non-decreasing."),
    (dict(frame_end_seconds=_t(1, 1)), "Each frame must satisfy end_seconds > start_seconds."),
    # several rules broken at once: the first in the reference's order wins
    (dict(embeddings=_rows(2) + NAN, frame_start_seconds=_t(1, 0), frame_end_seconds=_t(NAN, 1), backend_id=""),
     "EncodedSequence backend_id must be a non-empty string."),
    (dict(embeddings=_rows(2) + NAN, frame_start_seconds=_t(1, 0)), "EncodedSequence embeddings contain non-finite values."),
    (dict(frame_start_seconds=_t(1, 0), frame_end_seconds=_t(1, 0.5)), "frame_start_seconds must be non-decreasing."),
]


def _outcome(make):
    try:
        make()
    except ValueError as exc:
        return str(exc)
    return None


_THEIRS: list = []


def _reference_backend():
    """The reference's module loaded from its file alone (it imports numpy and typing only), so that
    neither its package ``__init__`` chain nor any ``sys.path`` change leaks into the session."""
    if not _THEIRS:
        module = None
        for root in (Path("/root/reference"), REPO / "baseline" / "_ref"):
            path = root / "ser" / "_internal" / "repr" / "backend.py"
            if path.is_file():
                import importlib.util

                spec = importlib.util.spec_from_file_location("_reference_repr_backend", path)
                module = importlib.util.module_from_spec(spec)
                sys.modules[spec.name] = module           # dataclasses resolves annotations through sys.modules
                spec.loader.exec_module(module)
                break
        _THEIRS.append(module)
    return _THEIRS[0]


@pytest.mark.parametrize("bounds,expected", WINDOW_CASES)
def test_pooling_window_validation(bounds, expected):
    assert _outcome(lambda: mine.PoolingWindow(*bounds)) == expected
    theirs = _reference_backend()
    if theirs is not None:
        assert _outcome(lambda: theirs.PoolingWindow(*bounds)) == expected


@pytest.mark.parametrize("change,expected", SEQUENCE_CASES)
def test_encoded_sequence_validation(change, expected):
    fields = {**_OK, **change}
    assert _outcome(lambda: mine.EncodedSequence(**fields)) == expected
    theirs = _reference_backend()
    if theirs is not None:
        assert _outcome(lambda: theirs.EncodedSequence(**fields)) == expected


def test_overlap_frame_mask_texts_and_mask():
    encoded = mine.EncodedSequence(embeddings=_rows(3), frame_start_seconds=_t(0, 1, 2), frame_end_seconds=_t(1, 2, 3),
                                   backend_id="x")
    assert mine.overlap_frame_mask(encoded, mine.PoolingWindow(0.5, 1.5)).tolist() == [True, True, False]
    assert mine.overlap_frame_mask(encoded, mine.PoolingWindow(1.0, 2.0)).tolist() == [False, True, False]
    with pytest.raises(ValueError, match=r"Pooling window is outside encoded sequence range: \[2.5, 3.5\] vs \[0.0, 3.0\]"):
        mine.overlap_frame_mask(encoded, mine.PoolingWindow(2.5, 3.5))
    gap = mine.EncodedSequence(embeddings=_rows(2), frame_start_seconds=_t(0, 2), frame_end_seconds=_t(1, 3), backend_id="x")
    with pytest.raises(ValueError, match=r"Pooling window does not overlap any encoded frames: \[1.25, 1.75\]"):
        mine.overlap_frame_mask(gap, mine.PoolingWindow(1.25, 1.75))
    theirs = _reference_backend()
    if theirs is not None:
        ref_encoded = theirs.EncodedSequence(embeddings=_rows(3), frame_start_seconds=_t(0, 1, 2),
                                             frame_end_seconds=_t(1, 2, 3), backend_id="x")
        for lo, hi in ((0.5, 1.5), (1.0, 2.0), (0.0, 3.0), (2.999, 3.0)):
            assert np.array_equal(theirs.overlap_frame_mask(ref_encoded, theirs.PoolingWindow(lo, hi)),
                                  mine.overlap_frame_mask(encoded, mine.PoolingWindow(lo, hi)))
