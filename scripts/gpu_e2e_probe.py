import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from ser_b200 import _native, mlp, synth
from ser_b200.config import FeatureFlags, flag_bits, feature_dim
import bench
ctx = _native.get_context(0)
flags = FeatureFlags(); bits = flag_bits(flags); dim = 193
n_clips, n_samples, sr = 1440, 168000, 48000
wave = synth.batch_audio_torch(n_clips, sr, n_samples, device="cuda").reshape(-1).contiguous()
starts, lengths, _ = bench.window_plan(n_clips, n_samples, sr)
rng = np.random.default_rng(0)
w = mlp.MlpWeights(mean=rng.standard_normal(dim), scale=1.0 + rng.random(dim), w1=rng.standard_normal((dim, 300)) * 0.1, b1=rng.standard_normal(300) * 0.1, w2=rng.standard_normal((300, 8)) * 0.1, b2=rng.standard_normal(8) * 0.1, classes=tuple(sorted(synth.RAVDESS_EMOTIONS.values())), out_activation=_native.OUT_SOFTMAX)
mlp.ensure_loaded(w, 0)
host = torch.empty(wave.numel(), dtype=torch.float32, pin_memory=True); host.copy_(wave); torch.cuda.synchronize()
hw = host.numpy()
# pure H2D bandwidth
dst = torch.empty_like(wave)
for _ in range(2):
    t0 = time.perf_counter(); dst.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D 968MB: %.1f ms (%.1f GB/s)" % (dt * 1e3, hw.nbytes / dt / 1e9))
for _ in range(2): ctx.infer_host(hw, starts, lengths, sr, bits, want_features=False)
for i in range(4):
    t0 = time.perf_counter(); ctx.infer_host(hw, starts, lengths, sr, bits, want_features=False); dt = time.perf_counter() - t0
    print("infer_host wall %.1f ms, device chain %.1f ms" % (dt * 1e3, ctx.last_compute_ms()))
