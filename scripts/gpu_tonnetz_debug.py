"""Stage-by-stage comparison of the tonnetz chain with the oracle (run on the GPU box)."""
import sys
import warnings

import numpy as np

sys.path.insert(0, ".")
warnings.simplefilter("ignore")
from ser_b200 import _native, synth  # noqa: E402
from oracle.shim import librosa  # noqa: E402

golden = np.load("tests/golden/fast_profile_golden.npz")
ctx = _native.get_context(0)
names = sys.argv[1:] or ["c16k_3s", "c48k_3p5s", "c22k_2s", "c44k_1s", "c16k_tail_5937", "c16k_2048", "sine16k_1p5s",
                         "silence16k", "c16k_short_1500", "c16k_short_1001", "c16k_short_300", "c48k_short_512"]
for name in names:
    audio = synth.decode_pcm16(golden[f"{name}/pcm"])
    sr = int(golden[f"{name}/sr"])
    if audio.size < 512:
        audio_p = np.pad(audio, (0, 512 - audio.size))
    else:
        audio_p = audio
    got = ctx.debug_tonnetz_stages(audio, sr)
    yh = librosa.effects.harmonic(audio_p)
    e_h = np.abs(got["yharm"] - yh).max() / max(np.abs(yh).max(), 1e-30)
    tun = librosa.estimate_tuning(y=yh, sr=sr, bins_per_octave=36)
    tun_idx = int(round((tun + 0.5) * 100))
    C = np.abs(librosa.cqt(yh, sr=sr, hop_length=512, n_bins=252, bins_per_octave=36, tuning=tun)).T
    cq = got["cqmag"]
    if cq.shape == C.shape:
        e_c = np.abs(cq - C).max() / max(C.max(), 1e-30)
    else:
        e_c = float("nan")
    ton = np.mean(librosa.feature.tonnetz(y=yh, sr=sr), axis=1)
    ref = golden[f"{name}/features"][187:193]
    e_t = np.abs(got["tonnetz"] - ton).max()
    print(f"{name:>18s} harm {e_h:.2e} tuning {got['tuning_index']} vs {tun_idx}  cq {cq.shape} vs {C.shape} {e_c:.2e}  "
          f"tonnetz abs err {e_t:.2e} (golden {np.abs(got['tonnetz'] - ref).max():.2e}) max|ton| {np.abs(ref).max():.3f}")
