"""Seams B4 / B5 for real: the REFERENCE's own ``ser --train``, ``ser.api.infer``, spawn-isolated fast
worker and legacy Pool loader, run unmodified from the staged checkout (baseline/_ref/ser, put there
by scripts/stage_reference.py -- git-ignored, shipped with the snapshot) with
``ser_b200.install.install()`` swapped in underneath, on a GPU.

Each surface runs in a subprocess through scripts/run_reference_surface.py.  On the ``b200`` arm the
oracle shim's feature entry points are POISONED (they raise), so a pass proves the arithmetic came from
libser_b200; the ``oracle`` arm runs the same reference code on its stock call path (the shim standing
in for librosa) and provides the expected labels / segments.

Reference call sites: ser/api.py:165-202; ser/_internal/runtime/backend_hooks.py:56-59, 95-115;
ser/_internal/runtime/fast_public_boundary.py:139-200, 251-277;
ser/_internal/data/data_loader.py:374-379, 467-535; ser/_internal/models/fast_training.py:166-197.
"""

from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
STAGED = REPO / "baseline" / "_ref"
DRIVER = REPO / "scripts" / "run_reference_surface.py"

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (STAGED / "ser").is_dir(), reason="reference checkout not staged (scripts/stage_reference.py)")]


def _run(env, arm, mode, *rest, flags=(), check=True):
    cmd = [sys.executable, str(DRIVER), "--arm", arm, *flags, mode, "--", *rest]
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=REPO)
    if check and proc.returncode != 0:
        raise AssertionError(f"{' '.join(cmd)} failed ({proc.returncode}):\n{proc.stdout[-1500:]}\n{proc.stderr[-3000:]}")
    return proc


def _infer(env, arm, path, flags=()):
    proc = _run(env, arm, "infer", str(path), flags=flags)
    return json.loads([line for line in proc.stdout.splitlines() if line.startswith("{")][-1])


@pytest.fixture(scope="module")
def workspace(tmp_path_factory):
    """A 64-file synthetic RAVDESS tree (the reference's own generator), dataset consents, and a model
    trained by the reference's ``ser --train --profile fast`` THROUGH install() on the GPU."""
    root = tmp_path_factory.mktemp("ser_workspace")
    env = dict(os.environ)
    for name in ("home", "data", "cache", "dataset"):
        (root / name).mkdir()
    env.update(HOME=str(root / "home"), SER_DATA_DIR=str(root / "data"), SER_CACHE_DIR=str(root / "cache"),
               DATASET_FOLDER=str(root / "dataset"), PYTHONWARNINGS="ignore")
    env.pop("SER_FAST_PROCESS_ISOLATION", None)
    subprocess.run([sys.executable, str(STAGED / "scripts" / "build_synthetic_ravdess_dataset.py"), "--output-root",
                    str(root / "dataset"), "--actors", *[str(a) for a in range(1, 9)]], check=True, capture_output=True)
    assert len(list((root / "dataset").glob("Actor_*/*.wav"))) == 64
    _run(env, "oracle", "cli", "configure", "--accept-dataset-policy", "noncommercial",
         "--accept-dataset-license", "cc-by-nc-sa-4.0", "--persist")
    train = _run(env, "b200", "cli", "--train", "--profile", "fast", "--preflight", "off", flags=("--poison-oracle",))
    assert (root / "data" / "models" / "ser_model.pkl").exists(), train.stderr[-2000:]
    return {"env": env, "root": root, "train_log": train.stderr}


def test_ser_train_ran_on_the_gpu_path(workspace):
    log = workspace["train_log"]
    assert "Model saved to" in log and "Training completed" in log
    assert "PREPARE" in log or "processed=" in log          # the reference's own progress records


def test_api_infer_matches_the_reference_on_its_stock_path(workspace):
    env = workspace["env"]
    sample = STAGED / "sample.wav"
    gpu = _infer(env, "b200", sample, flags=("--poison-oracle",))
    ref = _infer(env, "oracle", sample)
    assert gpu["native_loaded"] and gpu["gpu_kernel_launches"] > 0 and not ref["native_loaded"]
    assert gpu["profile"] == "fast" and gpu["backend_id"] == "handcrafted" and gpu["used_backend_path"] is True
    assert gpu["output_schema_version"] == ref["output_schema_version"]
    assert gpu["phase_timings_seconds"] and all(v >= 0 for v in gpu["phase_timings_seconds"].values())
    # bit-identical labels and timestamps, frame by frame and segment by segment
    assert [f[:3] for f in gpu["frames"]] == [f[:3] for f in ref["frames"]]
    assert [s[:3] for s in gpu["segments"]] == [s[:3] for s in ref["segments"]]
    assert gpu["emotions"] == ref["emotions"]
    assert [f[1:3] for f in gpu["frames"]] == [[0.0, 3.0], [1.0, 4.0], [2.0, 4.3710625], [3.0, 4.3710625], [4.0, 4.3710625]]
    np.testing.assert_allclose([f[3] for f in gpu["frames"]], [f[3] for f in ref["frames"]], rtol=0, atol=1e-4)
    for a, b in zip(gpu["probabilities"], ref["probabilities"]):
        assert list(a) == list(b)
        np.testing.assert_allclose(list(a.values()), list(b.values()), rtol=0, atol=1e-4)


def test_spawn_isolated_fast_worker_runs_on_the_gpu(workspace):
    """SER_FAST_PROCESS_ISOLATION=1: the reference spawns a fresh interpreter per attempt
    (process_timeout.py:46); install()'s worker entry patches it and creates the CUDA context lazily."""
    env = dict(workspace["env"], SER_FAST_PROCESS_ISOLATION="1", SER_FAST_TIMEOUT_SECONDS="120")
    sample = STAGED / "sample.wav"
    isolated = _infer(env, "b200", sample, flags=("--poison-oracle",))
    inline = _infer(workspace["env"], "b200", sample, flags=("--poison-oracle",))
    assert [f[:3] for f in isolated["frames"]] == [f[:3] for f in inline["frames"]]
    assert isolated["segments"] == inline["segments"]
    # the parent process never ran a kernel: the work happened in the spawned child
    assert isolated["gpu_kernel_launches"] == 0 and inline["gpu_kernel_launches"] > 0


def test_prepare_only_payload_is_float64(workspace):
    env = workspace["env"]
    _run(env, "b200", "cli", "--train", "--profile", "fast", "--preflight", "off", "--prepare-only", flags=("--poison-oracle",))
    payloads = list(Path(env["SER_CACHE_DIR"]).rglob("*.npz")) + list(Path(env["SER_DATA_DIR"]).rglob("*.npz"))
    assert payloads, "no prepared-feature payload was published"
    with np.load(payloads[0], allow_pickle=False) as data:
        assert data["x_train"].dtype == np.float64 and data["x_train"].shape[1] == 193
        assert np.all(np.isfinite(data["x_train"])) and data["x_test"].dtype == np.float64


def test_legacy_pool_loader_does_not_fork_with_cuda(workspace):
    """ser/_internal/data/data_loader.py:335-444 (mp.Pool over process_file) after the CUDA context
    already exists in the parent: rows equal the per-file calls, nothing forks."""
    env = workspace["env"]
    code = """
import sys, json, numpy as np
sys.path[:0] = [%r, %r, %r]
from ser_b200 import _native, install
_native.get_context(0)                     # CUDA is live BEFORE the loader would fork
install.install(device=0)
import multiprocessing
def no_fork(*a, **k):
    raise AssertionError("the legacy loader forked a Pool with a live CUDA context")
multiprocessing.Pool = no_fork
from ser.config import reload_settings
from ser._internal.data import data_loader
settings = reload_settings()
split = data_loader.load_data(test_size=0.25, settings=settings)
x_train, x_test, y_train, y_test = split
from ser._internal.features.feature_extractor import extract_feature
import glob
files = sorted(glob.glob(settings.dataset.glob_pattern))
rows = {tuple(np.round(extract_feature(f, settings=settings), 12)) for f in files}
batch = {tuple(np.round(r, 12)) for r in np.vstack([x_train, x_test])}
print(json.dumps({"n": int(len(x_train) + len(x_test)), "files": len(files), "dtype": str(x_train.dtype),
                  "dim": int(x_train.shape[1]), "same_rows": batch <= rows, "labels": sorted(set(y_train) | set(y_test))}))
""" % (str(STAGED), str(REPO / "oracle" / "shim"), str(REPO))
    proc = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900, cwd=REPO)
    assert proc.returncode == 0, proc.stderr[-3000:]
    out = json.loads([line for line in proc.stdout.splitlines() if line.startswith("{")][-1])
    assert out["n"] == out["files"] == 64 and out["dtype"] == "float64" and out["dim"] == 193
    assert out["same_rows"] and len(out["labels"]) >= 2
