"""Tonnetz parity of very short clips (1 .. 64 samples) against the CPU oracle, optionally with an
alternative build of the library (SERB_LIB=path) to compare two kernels on the same inputs.

    python scripts/gpu_tiny_lengths.py            # prints length, scaled tonnetz error
"""
import os
import sys
import warnings
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ser_b200 import _native  # noqa: E402

if os.environ.get("SERB_LIB"):
    _native.LIB_PATH = Path(os.environ["SERB_LIB"]).resolve()
from oracle import ser_oracle  # noqa: E402
from ser_b200 import dsp, synth  # noqa: E402
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("serb_tests_conftest", Path(__file__).resolve().parents[1] / "tests" / "conftest.py")
_conftest = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_conftest)
group_errors = _conftest.group_errors

sr = 16000
worst = 0.0
for length in [1] + list(range(3, 65)):
    audio = synth.clip_audio(synth.ClipSpec(50 + length % 7, 2 + length % 20, 1 + length % 8), sr, max(length, 8))[:length]
    if not np.any(audio):
        audio = audio + np.float32(0.25)
    got = dsp.extract_feature_from_signal(audio, sr)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ser_oracle.extract_feature_from_signal(audio, sr)
    report = group_errors(got, ref, groups=("mfcc", "chroma", "mel", "contrast", "tonnetz"))
    worst = max(worst, report["tonnetz"][0])
    print(length, f"tonnetz {report['tonnetz'][0]:.2e}", f"chroma {report['chroma'][0]:.2e}", flush=True)
print("worst tonnetz", f"{worst:.2e}")
