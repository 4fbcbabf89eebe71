"""Drop-in installation behind the reference's fast-profile seams.

``install()`` swaps the CUDA implementations into an importable jsugg/ser checkout at the
exact attributes the reference resolves at call time:

=====================================================================  =========================================
reference attribute                                                    replaced by
=====================================================================  =========================================
ser._internal.utils.dsp.extract_feature_from_signal                    ser_b200.dsp.extract_feature_from_signal
ser._internal.features.feature_extractor._extract_feature_from_signal  (same; it is a from-import alias)
ser._internal.repr.handcrafted.HandcraftedBackend.encode_sequence      one ragged-batch GPU call per recording
ser._internal.models.fast_path.predict_emotions_detailed_with_model    ser_b200.fast_path (fused CUDA MLP)
ser._internal.models.emotion_model._fast_predict_emotions_detailed_with_model  (same; from-import alias)
ser._internal.data.data_loader.load_checked_fast_data                  ser_b200.data_loader (ragged GPU batches)
ser._internal.features.feature_extractor._extract_feature_frames_for_settings  16-bit PCM WAV: int16 to the device (N1),
ser._internal.features.feature_extractor._extract_feature_for_settings         anything else: the reference's reader
ser._internal.data.data_loader._extract_feature_for_settings                   (same; from-import alias)
ser._internal.data.data_loader.mp                                              ``Pool`` without a fork (below)
ser._internal.runtime.fast_public_boundary._fast_worker_entry                 re-installs inside the spawned worker
=====================================================================  =========================================

Process boundaries.  The reference's spawn-isolated fast worker (``SER_FAST_PROCESS_ISOLATION=1``,
fast_public_boundary.py:251-277) starts a fresh interpreter, where no patch exists; its target is
therefore swapped for ``spawned_fast_worker_entry`` (module-level, picklable by name), which calls
``install()`` in the child -- the CUDA context is created lazily there -- and then runs the
reference's own entry.  The reference's legacy training loader forks a ``multiprocessing.Pool``
(data_loader.py:374-379): a CUDA context does not survive a fork, so that module's ``mp`` is
replaced by an object whose ``Pool`` runs ``process_file`` over the files IN THIS PROCESS, feeding
them to the GPU in ragged blocks (same results, same ``imap_unordered`` interface, no fork).

``ser.api.infer``, ``ser --file``, ``ser --train`` and ``run_fast_inference`` keep their
signatures; the registry hook ``"handcrafted"`` still resolves
``ser._internal.runtime.fast_inference.run_fast_inference`` by module path
(ser/_internal/runtime/backend_hooks.py:56-59) and reaches the patched functions through it.
``uninstall()`` restores the originals.
"""

from __future__ import annotations

import importlib
import os
from typing import Any

from . import data_loader as _data_loader
from . import dsp as _dsp
from . import fast_path as _fast_path
from .audio import read_pcm16_file
from .handcrafted import HandcraftedBackend as _GpuBackend
from .handcrafted import frame_bounds

_originals: list[tuple[Any, str, Any]] = []
_DEVICE_ENV = "SERB_INSTALL_DEVICE"


def spawned_fast_worker_entry(payload, connection) -> None:
    """Target of the reference's spawn-isolated fast worker after ``install()``: patches the fresh
    interpreter, then hands over to the reference's own ``_fast_worker_entry``."""
    install(device=int(os.environ.get(_DEVICE_ENV, "0")))
    boundary = importlib.import_module("ser._internal.runtime.fast_public_boundary")
    original = next(value for owner, name, value in _originals if owner is boundary and name == "_fast_worker_entry")
    original(payload, connection)


class _InProcessPool:
    """``multiprocessing.Pool`` stand-in for ser/_internal/data/data_loader.py:374-379: no fork (a CUDA
    context does not survive one).  ``imap_unordered(partial(process_file, ...), files)`` is recognised
    and served in ragged GPU blocks; any other callable is mapped serially."""

    def __init__(self, ref_loader, device: int, _processes=None) -> None:
        self._ref_loader = ref_loader
        self._device = device

    def __enter__(self):
        return self

    def __exit__(self, *exc) -> bool:
        return False

    def imap_unordered(self, fn, files, chunksize: int = 1):
        keywords = getattr(fn, "keywords", None) or {}
        target = getattr(fn, "func", None)
        wanted = {"observed_emotions", "emotion_map", "feature_flags", "audio_read_config"}
        if target is not self._ref_loader.process_file or not wanted <= set(keywords) or getattr(fn, "args", ()):
            return (fn(file) for file in files)
        return self._process_files(list(files), **{k: keywords[k] for k in wanted})

    map = imap = imap_unordered

    def _process_files(self, files, *, observed_emotions, emotion_map, feature_flags, audio_read_config):
        # process_file (data_loader.py:241-292) with the feature call batched: label filter first,
        # then one streamed GPU pass over the accepted files
        loader = self._ref_loader
        result_type = loader.ProcessFileResult
        results: list[Any] = [None] * len(files)
        accepted: list[int] = []
        emotions: dict[int, str] = {}
        for i, file in enumerate(files):
            name = os.path.basename(file)
            code = loader._extract_emotion_code(name)
            if code is None:
                results[i] = result_type(sample=None, error=f"Skipping file with unexpected name format (missing emotion code): {name}")
                continue
            emotion = emotion_map.get(code)
            if not emotion or emotion not in observed_emotions:
                results[i] = result_type(sample=None, error=None)
                continue
            accepted.append(i)
            emotions[i] = emotion

        def read_audio(path, *, start_seconds=None, duration_seconds=None):
            return loader.read_audio_file(path, start_seconds=start_seconds, duration_seconds=duration_seconds,
                                          audio_read_config=audio_read_config)

        entries = [(files[i], None, None) for i in accepted]
        for position, outcome in _data_loader.iter_file_features(entries, feature_flags=feature_flags, read_audio=read_audio,
                                                                 read_pcm16=read_pcm16_file, device=self._device):
            i = accepted[position]
            if isinstance(outcome, Exception):
                results[i] = result_type(sample=None, error=f"Failed to process file {files[i]}: {outcome}")
            else:
                results[i] = result_type(sample=(outcome, emotions[i]), error=None)
        return iter(results)


class _ForkFreeMultiprocessing:
    """What ``data_loader.mp`` becomes: everything of ``multiprocessing`` except that ``Pool`` does not fork."""

    def __init__(self, real, ref_loader, device: int) -> None:
        self._real = real
        self._ref_loader = ref_loader
        self._device = device

    def Pool(self, processes=None, *args, **kwargs):  # noqa: N802 - multiprocessing's name
        return _InProcessPool(self._ref_loader, self._device, processes)

    def __getattr__(self, name):
        return getattr(self._real, name)


def _swap(owner: Any, name: str, value: Any) -> None:
    _originals.append((owner, name, getattr(owner, name)))
    setattr(owner, name, value)


def _make_encode_sequence(encoded_sequence_type, device: int):
    def encode_sequence(self, audio, sample_rate):
        import numpy as np

        if sample_rate <= 0:
            raise ValueError("sample_rate must be a positive integer.")
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
        rows = _dsp.extract_features_ragged(wave, starts, ends - starts, sample_rate,
                                            feature_flags=self._feature_flags, device=device)
        return encoded_sequence_type(
            embeddings=rows.astype(np.float32, copy=False),
            frame_start_seconds=starts.astype(np.float64) / float(sample_rate),
            frame_end_seconds=ends.astype(np.float64) / float(sample_rate),
            backend_id=self.backend_id,
        )

    return encode_sequence


def install(device: int = 0) -> list[str]:
    """Patches the importable reference package in place; returns the patched attribute names."""
    if _originals:
        return [f"{getattr(o, '__name__', o)}.{n}" for o, n, _ in _originals]
    ref_dsp = importlib.import_module("ser._internal.utils.dsp")
    ref_features = importlib.import_module("ser._internal.features.feature_extractor")
    ref_handcrafted = importlib.import_module("ser._internal.repr.handcrafted")
    ref_backend = importlib.import_module("ser._internal.repr.backend")
    ref_fast_path = importlib.import_module("ser._internal.models.fast_path")
    ref_emotion_model = importlib.import_module("ser._internal.models.emotion_model")

    def extract_feature_from_signal(audio, sample_rate, *, feature_flags=None):
        return _dsp.extract_feature_from_signal(audio, sample_rate, feature_flags=feature_flags, device=device)

    def predict_emotions_detailed_with_model(file, *, model, expected_feature_size, output_schema_version,
                                             extract_feature_frames_fn, logger):
        return _fast_path.predict_emotions_detailed_with_model(
            file, model=model, expected_feature_size=expected_feature_size,
            output_schema_version=output_schema_version,
            extract_feature_frames_fn=extract_feature_frames_fn, logger=logger, device=device)

    ref_data_loader = importlib.import_module("ser._internal.data.data_loader")

    def load_checked_fast_data(*, utterances, settings, handle_sample_failure=None):
        # same signature as ser/_internal/data/data_loader.py:467-472; the reference's own splitter
        # and progress recorder are used unchanged
        from ser._internal.models.dataset_splitting import split_utterances
        from ser._internal.models.training_orchestration import record_preparation_progress

        def read_audio(path, *, start_seconds=None, duration_seconds=None):
            return ref_data_loader.read_audio_file(path, start_seconds=start_seconds,
                                                   duration_seconds=duration_seconds,
                                                   audio_read_config=settings.audio_read)

        if not utterances:
            return None
        train, test, _ = split_utterances(samples=list(utterances), settings=settings, logger=ref_data_loader.logger)
        common = dict(feature_flags=settings.feature_flags, handle_sample_failure=handle_sample_failure,
                      record_progress=record_preparation_progress, read_audio=read_audio,
                      read_pcm16=read_pcm16_file, device=device)
        x_train, y_train = _data_loader.extract_partition(train, **common)
        x_test, y_test = _data_loader.extract_partition(test, **common)
        if len(set(y_train)) < 2:
            raise RuntimeError("Fast checked preparation left fewer than two training classes.")
        return x_train, x_test, y_train, y_test

    original_frames = ref_features._extract_feature_frames_for_settings
    original_vector = ref_features._extract_feature_for_settings

    def _pcm16_or_none(file):
        try:
            return read_pcm16_file(file)
        except Exception:           # noqa: BLE001 - the reference's reader reports path / decode problems its own way
            return None

    def extract_feature_frames_for_settings(audiofile, *, frame_size, frame_stride, feature_flags, audio_read_config):
        # feature_extractor.py:70-103; 16-bit PCM WAV skips the host float32 pass (row N1)
        if frame_size <= 0:
            raise ValueError("frame_size must be greater than zero.")
        if frame_stride <= 0:
            raise ValueError("frame_stride must be greater than zero.")
        raw = _pcm16_or_none(audiofile)
        if raw is None:
            return original_frames(audiofile, frame_size=frame_size, frame_stride=frame_stride,
                                   feature_flags=feature_flags, audio_read_config=audio_read_config)
        import numpy as np

        encoded = _GpuBackend(frame_size_seconds=frame_size, frame_stride_seconds=frame_stride,
                              feature_flags=feature_flags, device=device).encode_sequence_pcm16(*raw)
        return [ref_features.FeatureFrame(start_seconds=float(encoded.frame_start_seconds[i]),
                                          end_seconds=float(encoded.frame_end_seconds[i]),
                                          features=np.asarray(encoded.embeddings[i], dtype=np.float64))
                for i in range(encoded.embeddings.shape[0])]

    def extract_feature_for_settings(file, *, feature_flags, audio_read_config):
        raw = _pcm16_or_none(file)
        if raw is None:
            return original_vector(file, feature_flags=feature_flags, audio_read_config=audio_read_config)
        return _GpuBackend(feature_flags=feature_flags, device=device).extract_vector_pcm16(*raw)

    _swap(ref_features, "_extract_feature_frames_for_settings", extract_feature_frames_for_settings)
    _swap(ref_features, "_extract_feature_for_settings", extract_feature_for_settings)
    if hasattr(ref_data_loader, "_extract_feature_for_settings"):
        _swap(ref_data_loader, "_extract_feature_for_settings", extract_feature_for_settings)
    if hasattr(ref_data_loader, "mp"):
        _swap(ref_data_loader, "mp", _ForkFreeMultiprocessing(ref_data_loader.mp, ref_data_loader, device))
    ref_boundary = importlib.import_module("ser._internal.runtime.fast_public_boundary")
    os.environ[_DEVICE_ENV] = str(int(device))
    _swap(ref_boundary, "_fast_worker_entry", spawned_fast_worker_entry)
    _swap(ref_data_loader, "load_checked_fast_data", load_checked_fast_data)
    _swap(ref_dsp, "extract_feature_from_signal", extract_feature_from_signal)
    _swap(ref_features, "_extract_feature_from_signal", extract_feature_from_signal)
    _swap(ref_handcrafted.HandcraftedBackend, "encode_sequence",
          _make_encode_sequence(ref_backend.EncodedSequence, device))
    _swap(ref_fast_path, "predict_emotions_detailed_with_model", predict_emotions_detailed_with_model)
    _swap(ref_emotion_model, "_fast_predict_emotions_detailed_with_model", predict_emotions_detailed_with_model)
    return [f"{getattr(o, '__name__', o)}.{n}" for o, n, _ in _originals]


def uninstall() -> None:
    """Restores every attribute ``install`` replaced."""
    while _originals:
        owner, name, value = _originals.pop()
        setattr(owner, name, value)
