"""Batched training-set feature extraction.

Mirror of ``load_checked_fast_data`` / ``_extract_partition``
(ser/_internal/data/data_loader.py:467-535): same inputs (readiness-checked utterances, a
settings snapshot, an optional ``handle_sample_failure`` quarantine callback), same outputs
(float64 feature matrix + label list per split), same failure semantics per sample, same
progress callback order -- but the feature vectors of a whole partition come from a few ragged
GPU calls (one per sample rate and per ~2^28-sample block) instead of one librosa pass per file.
Nothing forks: the reference's legacy Pool path (data_loader.py:368-379) is not needed.
"""

from __future__ import annotations

from collections.abc import Callable, Sequence
from typing import Any

import numpy as np
from numpy.typing import NDArray

from . import dsp
from .audio import read_audio_file
from .config import FeatureFlags

MAX_SAMPLES_PER_CALL = 1 << 28     # 1 GiB of float32 per native call


def _validated(audio: NDArray, sample_rate: int) -> NDArray[np.float32]:
    """Per-sample checks of dsp.extract_feature_from_signal, so that a bad clip fails alone."""
    audio = np.asarray(audio)
    if sample_rate <= 0:
        raise ValueError("Sample rate must be a positive integer.")
    if audio.ndim != 1:
        raise ValueError("Audio must be mono (1D array).")
    if audio.size == 0:
        raise ValueError("Audio contains no samples.")
    prepared = np.ascontiguousarray(audio, dtype=np.float32)
    if not bool(np.all(np.isfinite(prepared))):
        raise ValueError("Audio buffer is not finite everywhere.")
    return prepared


def extract_partition(
    partition: Sequence[Any],
    *,
    feature_flags: FeatureFlags | None = None,
    handle_sample_failure: Callable[[Any, Exception], bool] | None = None,
    record_progress: Callable[..., None] | None = None,
    read_audio: Callable[..., tuple[NDArray[np.float32], int]] = read_audio_file,
    extract_batch: Callable[..., NDArray[np.float64]] | None = None,
    device: int = 0,
) -> tuple[NDArray[np.float64], list[str]]:
    """Feature matrix and labels of one split partition (data_loader.py:485-529)."""
    flags = feature_flags if feature_flags is not None else FeatureFlags()
    if extract_batch is None:
        def extract_batch(clips, sample_rate):
            return dsp.extract_features_batch(clips, sample_rate, feature_flags=flags, device=device)

    total = len(partition)
    outcome: list[Any] = [None] * total          # feature row, or the exception of that sample
    by_rate: dict[int, list[tuple[int, NDArray[np.float32]]]] = {}
    for index, utterance in enumerate(partition):
        try:
            audio, sample_rate = read_audio(
                str(utterance.audio_path),
                start_seconds=getattr(utterance, "start_seconds", None),
                duration_seconds=getattr(utterance, "duration_seconds", None),
            )
            by_rate.setdefault(int(sample_rate), []).append((index, _validated(audio, int(sample_rate))))
        except Exception as error:  # noqa: BLE001 - routed to the caller's quarantine policy below
            outcome[index] = error
    for sample_rate, items in by_rate.items():
        block: list[tuple[int, NDArray[np.float32]]] = []
        size = 0

        def flush() -> None:
            nonlocal block, size
            if not block:
                return
            try:
                rows = extract_batch([clip for _, clip in block], sample_rate)
                for (index, _), row in zip(block, rows):
                    outcome[index] = np.asarray(row, dtype=np.float64)
            except Exception as error:  # noqa: BLE001 - e.g. a librosa ParameterError for this sample rate
                for index, _ in block:
                    outcome[index] = error
            block, size = [], 0

        for index, clip in items:
            if block and size + clip.size > MAX_SAMPLES_PER_CALL:
                flush()
            block.append((index, clip))
            size += clip.size
        flush()

    rows: list[NDArray[np.float64]] = []
    labels: list[str] = []
    for processed, (utterance, result) in enumerate(zip(partition, outcome), start=1):
        if isinstance(result, Exception):
            if handle_sample_failure is not None and handle_sample_failure(utterance, result):
                if record_progress is not None:
                    record_progress(processed=processed, total=total, sample_id=utterance.sample_id)
                continue
            raise result
        feature = result
        if feature.ndim != 1 or feature.size <= 0 or not np.all(np.isfinite(feature)):
            raise ValueError(f"Fast feature contract failed for sample {utterance.sample_id!r}.")
        rows.append(feature)
        labels.append(utterance.require_label())
        if record_progress is not None:
            record_progress(processed=processed, total=total, sample_id=utterance.sample_id)
    if not rows:
        raise RuntimeError("Fast checked preparation produced an empty split partition.")
    return np.vstack(rows).astype(np.float64, copy=False), labels


def load_checked_fast_data(
    *,
    utterances: Sequence[Any],
    settings: Any,
    handle_sample_failure: Callable[[Any, Exception], bool] | None = None,
    split_utterances: Callable[..., tuple[Sequence[Any], Sequence[Any], Any]],
    record_progress: Callable[..., None] | None = None,
    logger: Any = None,
    device: int = 0,
):
    """(x_train, x_test, y_train, y_test) from exactly the readiness-checked utterance view.

    ``split_utterances`` is the reference's ser/_internal/models/dataset_splitting.py function
    (injected, since dataset management is outside this repo's scope)."""
    if not utterances:
        return None
    train_utterances, test_utterances, _ = split_utterances(samples=list(utterances), settings=settings, logger=logger)
    flags = getattr(settings, "feature_flags", None)
    common = dict(feature_flags=flags, handle_sample_failure=handle_sample_failure,
                  record_progress=record_progress, device=device)
    x_train, y_train = extract_partition(train_utterances, **common)
    x_test, y_test = extract_partition(test_utterances, **common)
    if len(set(y_train)) < 2:
        raise RuntimeError("Fast checked preparation left fewer than two training classes.")
    return x_train, x_test, y_train, y_test
