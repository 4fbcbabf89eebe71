"""Runs the REFERENCE's own public surfaces from the staged checkout (baseline/_ref/ser), either on
its stock arithmetic path (the oracle shim standing in for librosa: ``--arm oracle``) or with
``ser_b200.install.install()`` swapped in underneath (``--arm b200``).  Used by
tests/test_gpu_reference_surfaces.py; both arms run the same unmodified reference code above the
feature / classifier seams.

    python scripts/run_reference_surface.py --arm b200 infer FILE            -> JSON on stdout
    python scripts/run_reference_surface.py --arm b200 cli -- --train --profile fast --preflight off

``oracle/shim`` is on sys.path only so that the reference's ``import librosa`` / ``soundfile`` /
``colored`` succeed (and so that librosa.load can decode PCM WAV); with ``--arm b200`` none of the
shim's feature arithmetic runs -- the test asserts that by poisoning it (``--poison-oracle``).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
STAGED = REPO / "baseline" / "_ref"


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["oracle", "b200"], required=True)
    ap.add_argument("--poison-oracle", action="store_true",
                    help="make every feature entry point of the shim raise: proves the b200 arm never calls it")
    ap.add_argument("--touch-cuda-first", action="store_true", help="create the CUDA context before anything else")
    ap.add_argument("mode", choices=["infer", "cli"])
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    args = ap.parse_args()

    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / "oracle" / "shim"))
    sys.path.insert(0, str(STAGED))
    import librosa  # the shim

    if args.poison_oracle:
        def poisoned(*_a, **_k):
            raise AssertionError("the oracle's arithmetic was called on the b200 arm")
        for owner, names in ((librosa, ("stft", "power_to_db")),
                             (librosa.feature, ("mfcc", "chroma_stft", "melspectrogram", "spectral_contrast", "tonnetz")),
                             (librosa.effects, ("harmonic",))):
            for name in names:
                setattr(owner, name, poisoned)
    if args.arm == "b200":
        from ser_b200 import _native, install

        if args.touch_cuda_first:
            _native.get_context(0)
        install.install(device=0)

    rest = [a for a in args.rest if a != "--"]
    if args.mode == "cli":
        from ser.__main__ import main as ser_main

        sys.argv = ["ser", *rest]
        ser_main()
        return

    import ser.api as api

    execution = api.infer(rest[0], profile="fast", include_transcript=False)
    detailed = execution.detailed_result
    out = {
        "profile": execution.profile, "backend_id": execution.backend_id, "used_backend_path": execution.used_backend_path,
        "output_schema_version": execution.output_schema_version,
        "phase_timings_seconds": dict(execution.phase_timings_seconds or {}),
        "emotions": [[e.emotion, e.start_seconds, e.end_seconds] for e in execution.emotions],
        "frames": [[f.emotion, f.start_seconds, f.end_seconds, f.confidence] for f in detailed.frames],
        "segments": [[s.emotion, s.start_seconds, s.end_seconds, s.confidence] for s in detailed.segments],
        "probabilities": [f.probabilities for f in detailed.frames],
        "native_loaded": any("libser_b200" in line for line in open("/proc/self/maps")),
    }
    if args.arm == "b200":
        from ser_b200 import _native

        out["gpu_kernel_launches"] = _native.get_context(0).launch_count
    os.write(3 if os.environ.get("SERB_RESULT_FD") == "3" else 1, (json.dumps(out) + "\n").encode())


if __name__ == "__main__":
    main()
