"""Batched training extraction: host logic on CPU with a stand-in extractor,
the real GPU call under the gpu marker.  Semantics follow ser/_internal/data/data_loader.py:467-535."""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np
import pytest

from ser_b200 import data_loader


@dataclass
class FakeUtterance:
    sample_id: str
    audio_path: Path
    label: str | None
    start_seconds: float | None = None
    duration_seconds: float | None = None

    def require_label(self) -> str:
        if self.label is None:
            raise ValueError(f"Utterance {self.sample_id!r} has no primary emotion target.")
        return self.label


def _reader(table):
    def read(path, *, start_seconds=None, duration_seconds=None):
        entry = table[Path(path).name]
        if isinstance(entry, Exception):
            raise entry
        return entry
    return read


def _fake_batch(clips, sample_rate):
    return np.stack([np.asarray([clip.size, float(clip.sum()), sample_rate], dtype=np.float64) for clip in clips])


def test_partition_groups_by_rate_keeps_order_and_quarantines_bad_samples():
    rng = np.random.default_rng(0)
    table = {
        "a.wav": (rng.standard_normal(4000).astype(np.float32), 16000),
        "b.wav": (rng.standard_normal(9000).astype(np.float32), 48000),
        "c.wav": FileNotFoundError("Audio file not found: c.wav"),
        "d.wav": (np.asarray([0.0, np.nan], dtype=np.float32), 16000),
        "e.wav": (rng.standard_normal(5000).astype(np.float32), 16000),
    }
    utterances = [FakeUtterance(n[0], Path(n), lab) for n, lab in zip(table, ["happy", "sad", "happy", "sad", "calm"])]
    seen, progress = [], []
    def handle(utterance, error):
        seen.append((utterance.sample_id, type(error).__name__, str(error)))
        return True
    x, y = data_loader.extract_partition(
        utterances, handle_sample_failure=handle, read_audio=_reader(table), extract_batch=_fake_batch,
        record_progress=lambda **kw: progress.append((kw["processed"], kw["total"], kw["sample_id"])))
    assert y == ["happy", "sad", "calm"]
    np.testing.assert_array_equal(x[:, 0], [4000, 9000, 5000])
    np.testing.assert_array_equal(x[:, 2], [16000, 48000, 16000])
    assert seen == [("c", "FileNotFoundError", "Audio file not found: c.wav"),
                    ("d", "ValueError", "Audio buffer is not finite everywhere.")]
    assert progress == [(i, 5, s) for i, s in zip(range(1, 6), "abcde")]
    # without a quarantine policy the first failing sample raises, as the reference does
    with pytest.raises(FileNotFoundError):
        data_loader.extract_partition(utterances, read_audio=_reader(table), extract_batch=_fake_batch)
    with pytest.raises(RuntimeError, match="empty split partition"):
        data_loader.extract_partition(utterances[2:4], handle_sample_failure=handle, read_audio=_reader(table),
                                      extract_batch=_fake_batch)


def test_feature_contract_and_split_checks():
    table = {"a.wav": (np.ones(100, dtype=np.float32), 16000), "b.wav": (np.ones(200, dtype=np.float32), 16000)}
    utterances = [FakeUtterance("a", Path("a.wav"), "happy"), FakeUtterance("b", Path("b.wav"), "happy")]
    bad = lambda clips, sr: np.full((len(clips), 3), np.inf)
    with pytest.raises(ValueError, match="Fast feature contract failed for sample 'a'"):
        data_loader.extract_partition(utterances, read_audio=_reader(table), extract_batch=bad)
    assert data_loader.load_checked_fast_data(utterances=[], settings=None, split_utterances=None) is None


def test_blocks_respect_the_per_call_sample_budget(monkeypatch):
    monkeypatch.setattr(data_loader, "MAX_SAMPLES_PER_CALL", 1000)
    table = {f"{i}.wav": (np.full(400, i, dtype=np.float32), 16000) for i in range(5)}
    utterances = [FakeUtterance(str(i), Path(f"{i}.wav"), "x") for i in range(5)]
    calls = []
    def batch(clips, sr):
        calls.append(len(clips))
        return _fake_batch(clips, sr)
    x, _ = data_loader.extract_partition(utterances, read_audio=_reader(table), extract_batch=batch)
    assert calls == [2, 2, 1] and x.shape == (5, 3)


def test_partition_is_streamed_block_by_block():
    """Host memory is O(block): a block's audio is extracted and dropped before later files are read,
    and finished samples are handed on (progress callbacks) while later files are still unread."""
    reads, events = [], []
    def read(path, *, start_seconds=None, duration_seconds=None):
        reads.append(Path(path).name)
        events.append(("read", Path(path).stem))
        return np.full(400, 1.0, dtype=np.float32), 16000
    def batch(clips, sr):
        events.append(("extract", len(clips)))
        return _fake_batch(clips, sr)
    utterances = [FakeUtterance(str(i), Path(f"{i}.wav"), "x") for i in range(6)]
    data_loader.extract_partition(utterances, read_audio=read, extract_batch=batch, max_samples_per_call=800,
                                  record_progress=lambda **kw: events.append(("progress", kw["sample_id"])))
    assert events == [("read", "0"), ("read", "1"), ("read", "2"), ("extract", 2), ("progress", "0"), ("progress", "1"),
                      ("read", "3"), ("read", "4"), ("extract", 2), ("progress", "2"), ("progress", "3"),
                      ("read", "5"), ("extract", 2), ("progress", "4"), ("progress", "5")]


def test_progress_keeps_partition_order_across_sample_rates():
    table = {"a.wav": (np.ones(300, dtype=np.float32), 16000), "b.wav": (np.ones(300, dtype=np.float32), 48000),
             "c.wav": (np.ones(300, dtype=np.float32), 16000), "d.wav": (np.ones(300, dtype=np.float32), 48000)}
    utterances = [FakeUtterance(n[0], Path(n), "x") for n in table]
    progress = []
    x, _ = data_loader.extract_partition(utterances, read_audio=_reader(table), extract_batch=_fake_batch,
                                         record_progress=lambda **kw: progress.append(kw["sample_id"]))
    assert progress == ["a", "b", "c", "d"]
    np.testing.assert_array_equal(x[:, 2], [16000, 48000, 16000, 48000])


def test_batch_level_errors_are_not_reported_as_sample_failures():
    from ser_b200 import dsp

    table = {f"{i}.wav": (np.ones(300, dtype=np.float32), 16000) for i in range(3)}
    utterances = [FakeUtterance(str(i), Path(f"{i}.wav"), "x") for i in range(3)]
    handled = []
    def handle(utterance, error):
        handled.append((utterance.sample_id, type(error).__name__))
        return True
    def cuda_oom(clips, sr):
        raise RuntimeError("cudaMalloc: out of memory")
    with pytest.raises(RuntimeError, match="out of memory"):
        data_loader.extract_partition(utterances, read_audio=_reader(table), extract_batch=cuda_oom, handle_sample_failure=handle)
    assert handled == []                      # a device failure is nobody's sample
    def nyquist(clips, sr):                   # deterministic argument error: every sample of the block fails alone upstream
        raise dsp.ParameterError("Frequency band exceeds Nyquist. Reduce either fmin or n_bands.")
    with pytest.raises(RuntimeError, match="empty split partition"):
        data_loader.extract_partition(utterances, read_audio=_reader(table), extract_batch=nyquist, handle_sample_failure=handle)
    assert handled == [("0", "ParameterError"), ("1", "ParameterError"), ("2", "ParameterError")]


@pytest.mark.gpu
def test_partition_on_the_gpu_matches_single_clip_calls(tmp_path, golden):
    import wave

    from ser_b200 import dsp, synth

    names = ["c16k_3s", "c48k_3p5s", "c16k_short_300", "c22k_2s"]
    utterances = []
    for i, name in enumerate(names):
        path = tmp_path / f"{name}.wav"
        with wave.open(str(path), "wb") as handle:
            handle.setnchannels(1)
            handle.setsampwidth(2)
            handle.setframerate(int(golden[f"{name}/sr"]))
            handle.writeframes(golden[f"{name}/pcm"].astype("<i2").tobytes())
        utterances.append(FakeUtterance(name, path, ["happy", "sad"][i % 2]))
    utterances.insert(2, FakeUtterance("missing", tmp_path / "missing.wav", "sad"))
    skipped = []
    x, y = data_loader.extract_partition(utterances, handle_sample_failure=lambda u, e: skipped.append(u.sample_id) or True)
    assert skipped == ["missing"] and y == ["happy", "sad", "happy", "sad"] and x.shape == (4, 193)
    for row, name in zip(x, names):
        single = dsp.extract_feature_from_signal(synth.decode_pcm16(golden[f"{name}/pcm"]), int(golden[f"{name}/sr"]))
        np.testing.assert_array_equal(row, single)
