"""Wire types of the fast-profile runtime boundary.

When the reference package is importable its own classes are re-exported, so results
produced here are instances of ser.runtime.schema.* / ser.domain.* / ser.runtime.contracts.*
and flow through the reference's pipeline unchanged.  Otherwise structurally identical
stand-ins are defined (ser/runtime/schema.py:14-53, ser/domain.py:23, ser/runtime/contracts.py:17).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple

OUTPUT_SCHEMA_VERSION = "v1"

try:  # pragma: no cover - exercised only where the reference is installed
    from ser.domain import EmotionSegment
    from ser.runtime.contracts import InferenceRequest
    from ser.runtime.schema import (
        FramePrediction,
        InferenceResult,
        SegmentPrediction,
        to_legacy_emotion_segments,
    )

    USING_REFERENCE_TYPES = True
except Exception:  # ImportError, or the reference failing to import its own dependencies
    USING_REFERENCE_TYPES = False

    class EmotionSegment(NamedTuple):
        """Emotion label over a time interval."""

        emotion: str
        start_seconds: float
        end_seconds: float

    @dataclass(frozen=True)
    class InferenceRequest:
        """Input of one inference execution."""

        file_path: str
        language: str = "en"
        save_transcript: bool = False
        include_transcript: bool = True
        subtitle_output_path: str | None = None
        subtitle_format: str | None = None

    @dataclass(frozen=True)
    class FramePrediction:
        """Prediction for one analysis window."""

        start_seconds: float
        end_seconds: float
        emotion: str
        confidence: float
        probabilities: dict[str, float] | None

    @dataclass(frozen=True)
    class SegmentPrediction:
        """Run of equal adjacent window labels."""

        emotion: str
        start_seconds: float
        end_seconds: float
        confidence: float
        probabilities: dict[str, float] | None = None

    @dataclass(frozen=True)
    class InferenceResult:
        """Frames plus merged segments."""

        schema_version: str
        segments: list[SegmentPrediction]
        frames: list[FramePrediction]

    def to_legacy_emotion_segments(result: InferenceResult) -> list[EmotionSegment]:
        return [EmotionSegment(s.emotion, s.start_seconds, s.end_seconds) for s in result.segments]
