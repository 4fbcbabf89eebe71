"""Fast-profile feature backend on the GPU (boundary B2).

Mirror of ser/_internal/repr/handcrafted.py:22-137: same constructor, ``backend_id``,
``feature_dim``, ``encode_sequence``, ``pool`` and ``extract_vector``.  ``encode_sequence``
turns the reference's per-window Python loop (handcrafted.py:85-97) into one ragged-batch
native call over the un-duplicated waveform: windows are (start, length) views of one buffer.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np
from numpy.typing import NDArray

from . import dsp
from .backend import EncodedSequence, PoolingWindow, overlap_frame_mask
from .config import FeatureFlags, feature_dim


def frame_bounds(n_samples: int, sample_rate: int, frame_size_seconds: float,
                 frame_stride_seconds: float) -> tuple[NDArray[np.int64], NDArray[np.int64]]:
    """(start, end) sample indices of the sliding windows (handcrafted.py:78-87)."""
    frame_length = max(1, int(round(frame_size_seconds * sample_rate)))
    frame_step = max(1, int(round(frame_stride_seconds * sample_rate)))
    starts = np.arange(0, n_samples, frame_step, dtype=np.int64)
    ends = np.minimum(starts + frame_length, n_samples)
    keep = ends > starts
    return starts[keep], ends[keep]


class HandcraftedBackend:
    """``FeatureBackend`` implementation of the handcrafted (librosa-style) feature set."""

    def __init__(
        self,
        *,
        frame_size_seconds: int = 3,
        frame_stride_seconds: int = 1,
        feature_flags: FeatureFlags | None = None,
        device: int = 0,
    ) -> None:
        if frame_size_seconds <= 0:
            raise ValueError("frame_size_seconds must be greater than zero.")
        if frame_stride_seconds <= 0:
            raise ValueError("frame_stride_seconds must be greater than zero.")
        self._frame_size_seconds = frame_size_seconds
        self._frame_stride_seconds = frame_stride_seconds
        self._feature_flags = feature_flags if feature_flags is not None else FeatureFlags()
        self._device = device

    @property
    def backend_id(self) -> str:
        return "handcrafted"

    @property
    def feature_dim(self) -> int:
        return feature_dim(self._feature_flags)

    def prepare_runtime(self) -> None:
        """No-op warm-up hook kept for contract parity (handcrafted.py:61-63)."""
        return None

    def encode_sequence(self, audio: NDArray[np.float32], sample_rate: int) -> EncodedSequence:
        """All sliding windows of ``audio`` in one GPU call."""
        if sample_rate <= 0:
            raise ValueError("sample_rate must be a positive integer.")
        audio = np.asarray(audio)
        if audio.ndim != 1:
            raise ValueError("audio must be mono (1D array).")
        if audio.size == 0:
            raise ValueError("audio must contain at least one sample.")
        wave = np.ascontiguousarray(audio, dtype=np.float32)
        if not bool(np.all(np.isfinite(wave))):
            raise ValueError("Audio buffer is not finite everywhere.")
        starts, ends = frame_bounds(wave.size, sample_rate, self._frame_size_seconds,
                                    self._frame_stride_seconds)
        if starts.size == 0:
            raise ValueError("Could not extract handcrafted features from provided audio.")
        embeddings = dsp.extract_features_ragged(
            wave, starts, ends - starts, sample_rate,
            feature_flags=self._feature_flags, device=self._device,
        )
        return EncodedSequence(
            embeddings=embeddings.astype(np.float32, copy=False),
            # exact float64 quotients index / sr, as handcrafted.py:96-97
            frame_start_seconds=starts.astype(np.float64) / float(sample_rate),
            frame_end_seconds=ends.astype(np.float64) / float(sample_rate),
            backend_id=self.backend_id,
        )

    def encode_sequence_pcm16(self, pcm: NDArray[np.int16], channels: int, sample_rate: int) -> EncodedSequence:
        """``encode_sequence`` of a raw 16-bit PCM file: decode scaling, mono mix, peak normalisation
        (audio_utils.py:28-60) and every sliding window in one GPU call, 2 bytes per sample over PCIe."""
        if sample_rate <= 0:
            raise ValueError("sample_rate must be a positive integer.")
        pcm = np.asarray(pcm)
        if pcm.ndim != 1 or pcm.dtype != np.int16:
            raise ValueError("pcm must be a 1-D int16 array of interleaved samples.")
        frames = pcm.size // max(int(channels), 1)
        if frames == 0:
            raise ValueError("audio must contain at least one sample.")
        starts, ends = frame_bounds(frames, sample_rate, self._frame_size_seconds, self._frame_stride_seconds)
        if starts.size == 0:
            raise ValueError("Could not extract handcrafted features from provided audio.")
        embeddings = dsp.extract_features_pcm16(
            [pcm], [int(channels)], np.zeros(starts.size, dtype=np.int64), starts, ends - starts, sample_rate,
            feature_flags=self._feature_flags, device=self._device)
        return EncodedSequence(
            embeddings=embeddings.astype(np.float32, copy=False),
            frame_start_seconds=starts.astype(np.float64) / float(sample_rate),
            frame_end_seconds=ends.astype(np.float64) / float(sample_rate),
            backend_id=self.backend_id,
        )

    def extract_vector_pcm16(self, pcm: NDArray[np.int16], channels: int, sample_rate: int) -> NDArray[np.float64]:
        """``extract_vector`` of a raw 16-bit PCM file (one whole-file row)."""
        pcm = np.asarray(pcm)
        frames = pcm.size // max(int(channels), 1)
        if frames == 0:
            raise ValueError("Audio contains no samples.")
        rows = dsp.extract_features_pcm16([pcm], [int(channels)], np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64),
                                          np.asarray([frames], dtype=np.int64), sample_rate,
                                          feature_flags=self._feature_flags, device=self._device)
        return rows[0].astype(np.float64)

    def pool(self, encoded: EncodedSequence, windows: Sequence[PoolingWindow]) -> NDArray[np.float64]:
        """Mean of the frames overlapping each window (handcrafted.py:109-122)."""
        from .pooling import mean_pool

        return mean_pool(encoded, windows, device=self._device)

    def extract_vector(self, audio: NDArray[np.float32], sample_rate: int) -> NDArray[np.float64]:
        """One whole-clip vector (training path, handcrafted.py:124-137)."""
        return np.asarray(
            dsp.extract_feature_from_signal(audio, sample_rate, feature_flags=self._feature_flags,
                                            device=self._device),
            dtype=np.float64,
        )

    def extract_vectors(self, clips: Sequence[NDArray[np.float32]], sample_rate: int) -> NDArray[np.float64]:
        """Batched ``extract_vector``: one native call for a list of clips."""
        return dsp.extract_features_batch(clips, sample_rate, feature_flags=self._feature_flags,
                                          device=self._device)
