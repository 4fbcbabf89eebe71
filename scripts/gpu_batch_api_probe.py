import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from ser_b200 import dsp, synth
sr, n, k = 48000, 168000, 1440
wave = synth.batch_audio_torch(k, sr, n, device="cuda").cpu().numpy()
clips = [np.ascontiguousarray(wave[i]) for i in range(k)]
dsp.extract_features_batch(clips[:8], sr)
for _ in range(3):
    t0 = time.perf_counter(); rows = dsp.extract_features_batch(clips, sr); dt = time.perf_counter() - t0
    print("extract_features_batch(1440 clips): %.1f ms  -> %.0f audio-s/s, %.0f clips/s" % (dt * 1e3, k * n / sr / dt, k / dt))
print(rows.shape, np.isfinite(rows).all())
