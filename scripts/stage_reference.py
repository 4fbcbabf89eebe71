"""Stages the reference's pure-Python package under baseline/_ref/ so that it travels to the GPU box.

    python scripts/stage_reference.py            # run in the build container (needs /root/reference)

``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference`` cannot run here: the reference's build backend (hatchling) and its arithmetic
dependencies (librosa, soundfile, soxr) have no wheel offline.  The package is pure Python, so an
"install" is the package tree itself: ``ser/`` plus the bundled sample.wav and the synthetic data-set
generator.  baseline/_ref/ is git-ignored (never committed) and not gpurun-ignored (it ships with the
snapshot), which is what tests/test_gpu_reference_surfaces.py needs to run the reference's OWN
``ser.api.infer`` / ``ser --train`` through ``ser_b200.install`` on a GPU.
"""

from __future__ import annotations

import shutil
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
REFERENCE = Path("/root/reference")
TARGET = REPO / "baseline" / "_ref"


def stage() -> Path | None:
    if not (REFERENCE / "ser").is_dir():
        return None
    TARGET.mkdir(parents=True, exist_ok=True)
    if (TARGET / "ser").exists():
        shutil.rmtree(TARGET / "ser")
    shutil.copytree(REFERENCE / "ser", TARGET / "ser", ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copyfile(REFERENCE / "sample.wav", TARGET / "sample.wav")
    (TARGET / "scripts").mkdir(exist_ok=True)
    shutil.copyfile(REFERENCE / "scripts" / "build_synthetic_ravdess_dataset.py",
                    TARGET / "scripts" / "build_synthetic_ravdess_dataset.py")
    return TARGET


if __name__ == "__main__":
    out = stage()
    print(out if out else "reference checkout not present; nothing staged")
    sys.exit(0)
