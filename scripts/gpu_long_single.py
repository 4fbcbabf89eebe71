import sys, time, numpy as np
sys.path.insert(0, ".")
from ser_b200 import dsp, synth
sr = 16000
audio = synth.long_recording(sr, 20 * 60 * sr, seed=3)
t0 = time.perf_counter(); a = dsp.extract_feature_from_signal(audio, sr); t1 = time.perf_counter()
b = dsp.extract_feature_from_signal(audio, sr); t2 = time.perf_counter()
print("20 min clip:", a.shape, np.isfinite(a).all(), np.array_equal(a, b), "first %.2fs second %.3fs" % (t1 - t0, t2 - t1))
print("tonnetz", a[187:])
try:
    big = np.zeros(90_000_000, dtype=np.float32); big[::1000] = 0.1
    dsp.extract_feature_from_signal(big, sr)
    print("90M samples ok")
except Exception as e:
    print("90M samples:", type(e).__name__, e)
