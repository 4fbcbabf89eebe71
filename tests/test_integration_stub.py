"""The ctypes stub INTEGRATION.md section 1 shows a reference maintainer is real code: it is cut out of the
document, pointed at the in-tree library and executed.  Without a device the call must end in the
library's own "no device" error (no CPU fallback); symbol names and argument counts are checked
against ser_b200/_native.py's signature table, which tests/test_abi.py ties to include/ser_b200.h."""

from __future__ import annotations

import re
from pathlib import Path

import numpy as np
import pytest

from ser_b200 import _native, build

REPO = Path(__file__).resolve().parents[1]


def _stub_source() -> str:
    text = (REPO / "INTEGRATION.md").read_text()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "ctypes.CDLL" in b)
    assert '"libser_b200.so"' in stub
    return stub.replace('"libser_b200.so"', repr(str(build.LIB_PATH)))


def test_stub_binds_existing_entries_with_the_declared_arity():
    source = _stub_source()
    for name, count in ((m.group(1), m.group(2).count(",") + 1) for m in
                        re.finditer(r"_lib\.(serb_\w+)\.argtypes = \[(.*?)\]", source, flags=re.S)):
        assert name in _native.SIGNATURES, name
        assert count == len(_native.SIGNATURES[name][1]), f"{name}: the stub passes {count} arguments"
    for name in re.findall(r"_lib\.(serb_\w+)\(", source):
        assert name in _native.SIGNATURES, name


def test_stub_runs_and_reports_the_missing_device():
    _native.load_library()                      # builds nothing: the library must already be in the tree
    namespace: dict = {}
    exec(compile(_stub_source(), "INTEGRATION.md#stub", "exec"), namespace)     # noqa: S102 - our own document
    audio = np.zeros(4096, dtype=np.float32)
    if _native.device_count() > 0:
        pytest.skip("a CUDA device is present; the GPU suite covers the computing path")
    with pytest.raises(RuntimeError) as err:
        namespace["features"](audio, [0], [4096], 16000, 0x1F)
    assert "device" in str(err.value).lower()
