"""Importing the package must not touch CUDA or even load the native library: the reference forks a
process pool for its legacy training loader (ser/_internal/data/data_loader.py:374-379) and imports
fresh in spawn workers (ser/_internal/runtime/process_timeout.py:46), so device state may only come
into being at the first computing call.  Checked in a clean interpreter through /proc/self/maps."""

from __future__ import annotations

import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]

PROBE = r"""
import sys
sys.path.insert(0, {repo!r})
import ser_b200
from ser_b200 import (_native, audio, backend, config, data_loader, dsp, fast_inference, fast_path,
                      feature_extractor, handcrafted, install, mlp, multi_gpu, pooling, schema, sharding, synth)
maps = open('/proc/self/maps').read()
assert _native._lib is None, 'libser_b200.so was loaded at import time'
assert not _native._contexts, 'a device context exists at import time'
for needle in ('libser_b200', 'libcuda.so', 'libcudart'):
    assert needle not in maps, needle + ' is mapped after a bare import'
assert 'torch' not in sys.modules, 'the host mirror imported torch'
# argument validation and host-only helpers still leave the library alone
try:
    dsp.extract_feature_from_signal([[0.0]], 16000)
except ValueError:
    pass
handcrafted.frame_bounds(1000, 16000, 3, 1)
assert _native._lib is None
import os
pid = os.fork()                 # the legacy loader's fork: the child can import and validate too
if pid == 0:
    try:
        dsp.extract_feature_from_signal([], 16000)
    except ValueError:
        os._exit(0)
    os._exit(3)
_, status = os.waitpid(pid, 0)
assert os.waitstatus_to_exitcode(status) == 0
print('lazy')
"""


def test_import_loads_no_native_code_and_survives_a_fork():
    proc = subprocess.run([sys.executable, "-c", PROBE.format(repo=str(REPO))], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-3000:]
    assert proc.stdout.strip().endswith("lazy")
