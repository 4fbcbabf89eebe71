"""Fast-profile inference boundary on the GPU (boundary B4).

``run_fast_inference(request, settings, *, loaded_model=None, enforce_timeout=True,
allow_retries=True)`` keeps the signature and the error taxonomy of
ser/_internal/runtime/fast_inference.py:35-57 / fast_public_boundary.py:139-411:

* ``FileNotFoundError`` while resolving the model  -> ``FastModelUnavailableError``
* ``ValueError`` while loading the model           -> ``FastModelLoadError``
* ``ValueError`` from the compute path             passes through unchanged
* ``RuntimeError`` (CUDA failures included)        -> ``FastInferenceExecutionError``

The reference's control plane around that call -- retry budget (policy.py:16-73), soft timeout
and spawn isolation (worker_lifecycle.py:98-208), the single-flight registry
(fast_public_boundary.py:319) -- is out of scope here ("used unchanged"):
``ser_b200.install.install()`` leaves the reference's own ``run_fast_inference`` in place and
swaps only the arithmetic underneath it, so those policies are the reference's code.  This
module is the same boundary for callers that do not have the reference package importable
(``enforce_timeout`` / ``allow_retries`` are accepted for signature parity and have nothing to
act on: there is no timeout or retry layer here).  Concurrent calls are safe: load + forward of
the classifier hold one per-device lock (``ser_b200.mlp.session``).
"""

from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Any

from . import fast_path
from .feature_extractor import extract_feature_frames
from .schema import OUTPUT_SCHEMA_VERSION, InferenceRequest, InferenceResult

logger = logging.getLogger(__name__)


class FastModelUnavailableError(FileNotFoundError):
    """No compatible fast-profile model artifact is available."""


class FastModelLoadError(RuntimeError):
    """The fast model artifact could not be loaded."""


class FastInferenceTimeoutError(TimeoutError):
    """Fast inference exceeded its timeout budget."""


class FastInferenceExecutionError(RuntimeError):
    """Fast inference failed after exhausting retries."""


class FastTransientBackendError(RuntimeError):
    """Retryable backend failure."""


@dataclass(frozen=True)
class LoadedModel:
    """Mirror of ser/_internal/models/artifact_envelope.py:33-38."""

    model: Any
    expected_feature_size: int | None = None
    artifact_metadata: dict | None = None


def _ensure_fast_compatible(loaded_model) -> None:
    metadata = getattr(loaded_model, "artifact_metadata", None)
    if not isinstance(metadata, dict):
        return
    backend_id = metadata.get("backend_id")
    profile = metadata.get("profile")
    if backend_id not in (None, "handcrafted") or profile not in (None, "fast"):
        raise FastModelUnavailableError(
            "No compatible fast-profile model artifact is available for backend_id='handcrafted'. "
            f"Found backend_id={backend_id!r}, profile={profile!r}."
        )


def _load_model(settings):
    """Uses the reference's artifact loader when the reference package is installed."""
    try:
        from ser._internal.models.emotion_model import load_model  # type: ignore
    except Exception as err:
        raise FastModelUnavailableError(
            "loaded_model was not given and the reference artifact loader "
            "(ser._internal.models.emotion_model.load_model) is not importable"
        ) from err
    try:
        return load_model(settings=settings, expected_backend_id="handcrafted", expected_profile="fast")
    except FileNotFoundError as err:
        raise FastModelUnavailableError(str(err)) from err
    except ValueError as err:
        raise FastModelLoadError("Failed to load fast-profile model artifact from configured paths.") from err


def predict_emotions_detailed(file: str, *, loaded_model, settings=None, device: int = 0) -> InferenceResult:
    """ser/_internal/models/emotion_model.py:140-163 with the GPU feature + MLP path."""
    return fast_path.predict_emotions_detailed_with_model(
        file,
        model=loaded_model.model,
        expected_feature_size=getattr(loaded_model, "expected_feature_size", None),
        output_schema_version=OUTPUT_SCHEMA_VERSION,
        extract_feature_frames_fn=lambda path: extract_feature_frames(path, settings=settings, device=device),
        logger=logger,
        device=device,
    )


def run_fast_inference(
    request: InferenceRequest,
    settings=None,
    *,
    loaded_model=None,
    enforce_timeout: bool = True,
    allow_retries: bool = True,
    device: int = 0,
) -> InferenceResult:
    """Runs fast-profile inference for ``request.file_path`` and returns frames + segments."""
    active_model = loaded_model if loaded_model is not None else _load_model(settings)
    _ensure_fast_compatible(active_model)
    del enforce_timeout, allow_retries      # the reference's control plane owns these (module docstring)
    try:
        # settings are deliberately NOT forwarded: the reference's hook reloads defaults, so all
        # five feature groups are always extracted
        return predict_emotions_detailed(request.file_path, loaded_model=active_model, device=device)
    except (ValueError, FastInferenceExecutionError):
        raise
    except RuntimeError as err:
        raise FastInferenceExecutionError(str(err)) from err
