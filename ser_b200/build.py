"""In-tree build of libser_b200.so (nvcc, sm_100a only).

    python -m ser_b200.build [--force] [--verbose]

The shared library lands next to this file (``ser_b200/libser_b200.so``) so that it travels
with the repository snapshot to the GPU box; nothing is installed into site-packages.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
SRC_DIR = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libser_b200.so"

SOURCES = [
    "filterbanks.cpp",
    "cqt_tables.cpp",
    "stft_kernel.cu",
    "proj_kernels.cu",
    "short_kernel.cu",
    "hpss_kernels.cu",
    "cqt_kernels.cu",
    "decimate_mma.cu",
    "mlp_kernel.cu",
    "pcm_kernels.cu",
    "api.cu",
]
HEADERS = ["common.cuh", "fft.cuh", "kernels.h", "filterbanks.h", "cqt_tables.h", "median_net.cuh", "../../include/ser_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unknown-pragmas",
    "-Xptxas", "-v",
    "-Wno-deprecated-gpu-targets",
]


def _nvcc() -> str:
    found = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(found).exists():
        raise RuntimeError("nvcc not found: ser_b200 needs the CUDA toolkit to build its only backend")
    return found


def _fingerprint() -> str:
    digest = hashlib.sha256()
    for name in SOURCES + HEADERS:
        digest.update(name.encode())
        digest.update((SRC_DIR / name).read_bytes())
    digest.update(" ".join(NVCC_FLAGS).encode())
    return digest.hexdigest()


def _compile_one(nvcc: str, name: str, verbose: bool) -> tuple[str, str]:
    obj = BUILD_DIR / (Path(name).stem + ".o")
    cmd = [nvcc, *NVCC_FLAGS, "-I", str(SRC_DIR), "-x", "cu", "-c", str(SRC_DIR / name), "-o", str(obj)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed on {name}:\n{log}")
    if verbose:
        print(f"--- {name}\n{log}")
    return str(obj), log


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compiles every source for sm_100a and links the shared library; returns its path."""
    BUILD_DIR.mkdir(exist_ok=True)
    stamp = BUILD_DIR / "fingerprint"
    fingerprint = _fingerprint()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == fingerprint:
        return LIB_PATH
    nvcc = _nvcc()
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda name: _compile_one(nvcc, name, verbose), SOURCES))
    objects = [obj for obj, _ in results]
    (BUILD_DIR / "ptxas.log").write_text("\n".join(f"--- {n}\n{log}" for n, (_, log) in zip(SOURCES, results)))
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *objects, "-Wno-deprecated-gpu-targets"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed:\n{proc.stdout}{proc.stderr}")
    stamp.write_text(fingerprint)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
