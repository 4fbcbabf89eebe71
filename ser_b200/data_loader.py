"""Batched training-set feature extraction.

Mirror of ``load_checked_fast_data`` / ``_extract_partition``
(ser/_internal/data/data_loader.py:467-535): same inputs (readiness-checked utterances, a
settings snapshot, an optional ``handle_sample_failure`` quarantine callback), same outputs
(float64 feature matrix + label list per split), same failure semantics per sample, same
progress callback order -- but the feature vectors come from ragged GPU calls over blocks of
files instead of one librosa pass per file.  The partition is STREAMED: files are read until a
block (per sample rate) holds ``MAX_SAMPLES_PER_CALL`` samples, the block is extracted, its audio
is dropped, and results are handed on in partition order, so host memory is O(block) like the
reference's O(file), not O(dataset).  16-bit PCM WAV files are shipped as int16 and prepared on
the device (row N1); anything else goes through ``read_audio``.
Nothing forks: the reference's legacy Pool path (data_loader.py:368-379) is not needed.
"""

from __future__ import annotations

from collections.abc import Callable, Sequence
from typing import Any

import numpy as np
from numpy.typing import NDArray

from . import dsp
from .audio import read_audio_file, read_pcm16_file
from .config import FeatureFlags

MAX_SAMPLES_PER_CALL = 1 << 28     # samples per native call (1 GiB as float32, 512 MiB as PCM16)
_PENDING = object()


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


class _Block:
    """Clips of one (sample rate, kind) waiting for their native call."""

    def __init__(self) -> None:
        self.items: list[tuple[int, Any]] = []     # (partition index, float32 clip | (int16 pcm, channels))
        self.samples = 0


def iter_file_features(
    files: Sequence[tuple[str, float | None, float | None]],
    *,
    feature_flags: FeatureFlags | None = None,
    read_audio: Callable[..., tuple[NDArray[np.float32], int]] = read_audio_file,
    read_pcm16: Callable[..., tuple[NDArray[np.int16], int, int] | None] | str | None = "auto",
    extract_batch: Callable[..., NDArray[np.float64]] | None = None,
    device: int = 0,
    max_samples_per_call: int | None = None,
):
    """Yields ``(position, outcome)`` for every ``(path, start_seconds, duration_seconds)`` entry IN
    ORDER, where ``outcome`` is the float64 feature row of that file or the exception that file raised.

    Files are read until a block (per sample rate) holds the per-call sample budget, the block goes to
    the GPU in one ragged call, its audio is dropped, and the finished prefix is yielded: host memory
    is O(block), and the consumer sees results while later files are still unread.  Failure routing
    follows the reference, where every sample runs alone inside one ``try`` (data_loader.py:495-512):
    an error that belongs to a sample -- decode / validation errors, and the deterministic argument
    errors of a block (``ValueError`` / ``ParameterError``, e.g. librosa's Nyquist check, which every
    sample of that block would raise on its own) -- is that sample's outcome.  A batch-level
    ``RuntimeError`` (CUDA failure, out of memory) is nobody's sample and propagates at once."""
    flags = feature_flags if feature_flags is not None else FeatureFlags()
    if read_pcm16 == "auto":          # the int16 fast path pairs with this package's own file reader only
        read_pcm16 = read_pcm16_file if (read_audio is read_audio_file and extract_batch is None) else None
    budget = MAX_SAMPLES_PER_CALL if max_samples_per_call is None else int(max_samples_per_call)
    if extract_batch is None:
        def extract_batch(clips, sample_rate):
            return dsp.extract_features_batch(clips, sample_rate, feature_flags=flags, device=device)

    def extract_pcm_batch(items, sample_rate):
        frames = np.asarray([pcm.size // ch for pcm, ch in items], dtype=np.int64)
        n = len(items)
        return dsp.extract_features_pcm16([pcm for pcm, _ in items], [ch for _, ch in items], np.arange(n, dtype=np.int64),
                                          np.zeros(n, dtype=np.int64), frames, sample_rate, feature_flags=flags,
                                          device=device).astype(np.float64)

    total = len(files)
    outcome: list[Any] = [_PENDING] * total
    blocks: dict[tuple[int, bool], _Block] = {}
    cursor = 0

    def drain():
        nonlocal cursor
        while cursor < total and outcome[cursor] is not _PENDING:
            result, outcome[cursor] = outcome[cursor], None
            cursor += 1
            yield cursor - 1, result

    def flush(key: tuple[int, bool]) -> None:
        block = blocks.pop(key, None)
        if block is None or not block.items:
            return
        sample_rate, is_pcm = key
        try:
            payload = [item for _, item in block.items]
            result = extract_pcm_batch(payload, sample_rate) if is_pcm else extract_batch(payload, sample_rate)
            for (index, _), row in zip(block.items, result):
                outcome[index] = np.asarray(row, dtype=np.float64)
        except (ValueError, dsp.ParameterError) as error:
            for index, _ in block.items:           # each sample would have raised this on its own
                outcome[index] = error
        block.items.clear()                        # the audio is dropped here

    for index, (path, start_seconds, duration_seconds) in enumerate(files):
        try:
            segment = dict(start_seconds=start_seconds, duration_seconds=duration_seconds)
            raw = read_pcm16(str(path), **segment) if read_pcm16 is not None else None
            if raw is not None:
                pcm, channels, sample_rate = raw
                if sample_rate <= 0:
                    raise ValueError("Sample rate must be a positive integer.")
                if pcm.size < channels:
                    raise OSError("Audio file contains no samples.")
                key, item, size = (int(sample_rate), True), (pcm, int(channels)), pcm.size // channels
            else:
                audio, sample_rate = read_audio(str(path), **segment)
                clip = _validated(audio, int(sample_rate))
                key, item, size = (int(sample_rate), False), clip, clip.size
        except Exception as error:  # noqa: BLE001 - a sample-local failure: the caller's policy decides
            outcome[index] = error
            yield from drain()
            continue
        block = blocks.get(key)
        if block is not None and block.items and block.samples + size > budget:
            flush(key)
            yield from drain()
            block = None
        if block is None:
            block = blocks[key] = _Block()
        block.items.append((index, item))
        block.samples += size
    for key in list(blocks):
        flush(key)
        yield from drain()
    yield from drain()


def extract_partition(
    partition: Sequence[Any],
    *,
    feature_flags: FeatureFlags | None = None,
    handle_sample_failure: Callable[[Any, Exception], bool] | None = None,
    record_progress: Callable[..., None] | None = None,
    read_audio: Callable[..., tuple[NDArray[np.float32], int]] = read_audio_file,
    read_pcm16: Callable[..., tuple[NDArray[np.int16], int, int] | None] | str | None = "auto",
    extract_batch: Callable[..., NDArray[np.float64]] | None = None,
    device: int = 0,
    max_samples_per_call: int | None = None,
) -> tuple[NDArray[np.float64], list[str]]:
    """Feature matrix and labels of one split partition (data_loader.py:485-529): quarantine
    callback, finite / shape contract and progress records in partition order, over
    ``iter_file_features``."""
    total = len(partition)
    files = [(str(u.audio_path), getattr(u, "start_seconds", None), getattr(u, "duration_seconds", None)) for u in partition]
    rows: list[NDArray[np.float64]] = []
    labels: list[str] = []
    for position, result in iter_file_features(files, feature_flags=feature_flags, read_audio=read_audio,
                                               read_pcm16=read_pcm16, extract_batch=extract_batch, device=device,
                                               max_samples_per_call=max_samples_per_call):
        utterance = partition[position]
        if isinstance(result, Exception):
            if handle_sample_failure is not None and handle_sample_failure(utterance, result):
                if record_progress is not None:
                    record_progress(processed=position + 1, total=total, sample_id=utterance.sample_id)
                continue
            raise result
        if result.ndim != 1 or result.size <= 0 or not np.all(np.isfinite(result)):
            raise ValueError(f"Fast feature contract failed for sample {utterance.sample_id!r}.")
        rows.append(result)
        labels.append(utterance.require_label())
        if record_progress is not None:
            record_progress(processed=position + 1, total=total, sample_id=utterance.sample_id)
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
