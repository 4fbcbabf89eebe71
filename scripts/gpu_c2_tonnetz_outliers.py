"""Diagnostic (GPU box): which c2 windows differ from the committed oracle rows in tonnetz, and why.
For every outlier prints the GPU's and the oracle's tuning bin of the harmonic signal (36 bins per
octave) and the histogram counts around the arg-max, to tell tuning flips from arithmetic errors."""
import sys
import warnings

import numpy as np

sys.path.insert(0, ".")
from oracle.shim import librosa  # noqa: E402
from oracle.shim.librosa import core  # noqa: E402
from ser_b200 import _native, synth  # noqa: E402
from ser_b200.handcrafted import HandcraftedBackend, frame_bounds  # noqa: E402

g = np.load("tests/golden/c2_oracle_rows.npz")
sr, n = 48000, 168000
specs = synth.ravdess_specs(1440)
backend = HandcraftedBackend()
ctx = _native.get_context(0)
bad = 0
for k, index in enumerate(g["clip_index"]):
    audio = synth.clip_audio(specs[int(index)], sr, n)
    rows = backend.encode_sequence(audio, sr).embeddings
    ref = g["window_rows"][4 * k: 4 * k + 4]
    err = np.abs(rows[:, 187:] - ref[:, 187:]).max(axis=1)
    starts, ends = frame_bounds(n, sr, 3, 1)
    for w in np.flatnonzero(err > 2e-5):
        bad += 1
        clip = audio[starts[w]:ends[w]]
        stages = ctx.debug_tonnetz_stages(clip, sr)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            yh = librosa.effects.harmonic(clip)
            pitch, mag = core.piptrack(y=yh, sr=sr)
            mask = pitch > 0
            thr = np.median(mag[mask])
            freqs = pitch[(mag >= thr) & mask]
            resid = np.mod(36 * np.log2(freqs / (440.0 / 16)), 1.0)
            resid[resid >= 0.5] -= 1.0
            counts, edges = np.histogram(resid, np.linspace(-0.5, 0.5, 101))
        o_bin = int(np.argmax(counts))
        top = np.argsort(counts)[::-1][:3]
        print(f"clip {int(index)} window {w}: tonnetz abs err {err[w]:.2e}  gpu bin {stages['tuning_index']} oracle bin {o_bin} "
              f"top counts {[(int(b), int(counts[b])) for b in top]}  yharm max diff {np.abs(stages['yharm'][:clip.size] - yh).max():.2e}")
print("windows above 2e-5:", bad, "of", 4 * len(g["clip_index"]))
