"""Parity fuzz against the oracle over signal families the synthetic generator does not cover."""
import sys, warnings, numpy as np
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
from ser_b200 import dsp
from oracle import ser_oracle
sys.path.insert(0, "tests")
from conftest import group_errors
rng = np.random.default_rng(2027)
def norm(x):
    x = np.asarray(x, dtype=np.float32)
    m = np.max(np.abs(x)); return x / m if m > 0 else x
worst = {}
for sr in (16000, 22050, 44100):
    n = int(1.7 * sr)
    t = np.arange(n) / sr
    signals = {
        "white": norm(rng.standard_normal(n)),
        "chirp": norm(np.sin(2 * np.pi * (100 * t + 0.5 * 2500 * t * t))),
        "dc+tone": norm(0.5 + 0.3 * np.sin(2 * np.pi * 440 * t)),
        "impulses": norm((np.arange(n) % 997 == 0).astype(np.float32)),
        "square": norm(np.sign(np.sin(2 * np.pi * 233.08 * t))),
        "am_noise": norm(rng.standard_normal(n) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t)) ** 2),
        "two_tones": norm(np.sin(2 * np.pi * 261.63 * t) + 0.7 * np.sin(2 * np.pi * 392.0 * t)),
        "quiet": (norm(rng.standard_normal(n)) * 1e-4).astype(np.float32),
    }
    for name, x in signals.items():
        got = dsp.extract_feature_from_signal(x, sr)
        ref = ser_oracle.extract_feature_from_signal(x, sr)
        rep = group_errors(got, ref, groups=("mfcc", "chroma", "mel", "contrast", "tonnetz"))
        bad = {k: f"{v[0]:.1e}" for k, v in rep.items() if v[0] > 1e-4}
        ton_abs = float(np.max(np.abs(got[187:] - ref[187:])))
        print(sr, name, {k: f"{v[0]:.1e}" for k, v in rep.items()}, "tonnetz abs err %.1e, max |tonnetz| %.1e" % (ton_abs, float(np.max(np.abs(ref[187:])))), "FAIL" if bad else "")
