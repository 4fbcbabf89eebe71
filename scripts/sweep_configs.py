"""BASELINE configs c3 and c5 on one B200 (run on the GPU box):

    python scripts/sweep_configs.py > gpurun_out/sweep_configs.json

c5: clip length {1, 2, 3.5, 5, 10, 30, 60} s x batch {1, 8, 64, 512, 4096} @ 48 kHz, whole-clip
193-d vectors (the training call), inputs resident in HBM, CUDA events, best of 3 after 2 warm-ups.
Combinations above 2^31 samples are skipped (stated in the output).
c3: 20 000 clips x 168 000 samples @ 48 kHz, whole-clip vectors, 1 GPU.
"""

import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ser_b200 import _native, synth  # noqa: E402
from ser_b200.config import FeatureFlags, flag_bits  # noqa: E402

SR = 48000
ctx = _native.get_context(0)
bits = flag_bits(FeatureFlags())


def timed(wave, starts, lengths, reps=3, warm=2):
    out = torch.empty((starts.size, 193), dtype=torch.float32, device="cuda")
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for _ in range(warm):
            ctx.features_device(wave.data_ptr(), wave.numel(), starts, lengths, SR, bits, out.data_ptr(), side.cuda_stream)
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            ctx.features_device(wave.data_ptr(), wave.numel(), starts, lengths, SR, bits, out.data_ptr(), side.cuda_stream)
            e1.record(side)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    assert bool(torch.isfinite(out).all())
    return best


results = {"sample_rate": SR, "feature_dim": 193, "c5": [], "skipped": []}
base = synth.batch_audio_torch(64, SR, 60 * SR, device="cuda")          # 64 distinct 60 s clips
for seconds in (1, 2, 3.5, 5, 10, 30, 60):
    n = int(seconds * SR)
    for batch in (1, 8, 64, 512, 4096):
        if batch * n > 2**31:
            results["skipped"].append({"clip_seconds": seconds, "batch": batch, "why": "more than 2^31 samples"})
            continue
        # clips are slices of the 64 base clips at varying offsets: distinct content, no extra memory
        reps = (batch + 63) // 64
        wave = base[:, :n].repeat(reps, 1)[:batch].contiguous().reshape(-1) if batch > 64 else base[:batch, :n].contiguous().reshape(-1)
        starts = np.arange(batch, dtype=np.int64) * n
        lengths = np.full(batch, n, dtype=np.int64)
        ms = timed(wave, starts, lengths)
        results["c5"].append({"clip_seconds": seconds, "batch": batch, "ms": ms,
                              "audio_s_per_s": batch * seconds / (ms / 1e3)})
        del wave
        torch.cuda.empty_cache()
del base
torch.cuda.empty_cache()

n_clips, n = 20000, 168000
wave = torch.empty((n_clips, n), dtype=torch.float32, device="cuda")
for lo in range(0, n_clips, 2000):
    wave[lo:lo + 2000] = synth.batch_audio_torch(2000, SR, n, device="cuda", first_index=lo)
wave = wave.reshape(-1)
starts = np.arange(n_clips, dtype=np.int64) * n
lengths = np.full(n_clips, n, dtype=np.int64)
ms = timed(wave, starts, lengths, reps=2, warm=1)
results["c3"] = {"clips": n_clips, "clip_samples": n, "ms": ms, "audio_s_per_s": n_clips * n / SR / (ms / 1e3),
                 "clips_per_s": n_clips / (ms / 1e3), "input_bytes": int(wave.numel() * 4)}
print(json.dumps(results))
