"""Prints SHA-256 prefixes of feature rows and STFT magnitudes for three sample rates: run it with two builds of
libser_b200.so on the same box to show a change is bit-identical (how the halved-window STFT was checked)."""
import hashlib, sys
import numpy as np
sys.path.insert(0, ".")
from ser_b200 import _native, synth
ctx = _native.get_context(0)
out = []
for sr, n in ((48000, 168000), (16000, 70000), (22050, 50001)):
    rng = np.random.default_rng(sr)
    clips = [synth.clip_audio(synth.ClipSpec(1 + i % 24, 1 + i % 2, 1 + i % 8), sr, n - 997 * i) for i in range(6)]
    clips.append((0.3 * rng.standard_normal(n // 3)).astype(np.float32))
    rows = ctx.features_host_clips(clips, sr, 0x1F)
    st = ctx.debug_stft_host(clips[0][:40000])
    out.append(hashlib.sha256(rows.tobytes()).hexdigest()[:16] + " " + hashlib.sha256(np.ascontiguousarray(st).tobytes()).hexdigest()[:16])
print(" | ".join(out))
