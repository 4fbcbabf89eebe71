"""e2e of the host entries from PAGEABLE numpy memory (what the Python mirror's callers pass) against
the staging-ring knobs: SERB_STAGE_THREADS (0 = pageable pointers handed to cudaMemcpyAsync as they
are) x SERB_RAMP_FACTOR_PAGEABLE_X10.  Config c2: 1 440 files x 168 000 int16 samples, 5 760 windows.

    python scripts/gpu_pageable_sweep.py > gpurun_out/pageable_sweep.txt
"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import bench
from ser_b200 import _native, synth

sr, n_samples, n_clips = 48000, 168000, int(os.environ.get("SWEEP_CLIPS", "1440"))
wave, pcm = synth.batch_audio_torch(n_clips, sr, n_samples, device="cuda", first_index=0, return_pcm=True)
pcm_host = pcm.cpu().numpy()
files = [np.array(pcm_host[i]) for i in range(n_clips)]                  # one pageable array per file
wave_host = wave.cpu().numpy()
clips32 = [np.array(wave_host[i]) for i in range(n_clips)]
flat32 = np.ascontiguousarray(wave_host.reshape(-1))
clip_of, w_starts, lengths, _ = bench.window_plan(n_clips, n_samples, sr)
starts = clip_of * n_samples + w_starts
bits = 31
audio_s = n_clips * n_samples / sr


def timed(fn, steps=3):
    out = fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    return (time.perf_counter() - t0) * 1e3 / steps, out


ref = {}
configs = [(0, 20)] + [(t, f) for t in (2, 4, 8) for f in (15, 20, 30, 40)]
if len(sys.argv) > 1:
    configs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for threads, factor in configs:
    os.environ["SERB_STAGE_THREADS"] = str(threads)
    os.environ["SERB_RAMP_FACTOR_PAGEABLE_X10"] = str(factor)
    ctx = _native.Context(0)
    line = [f"threads {threads} ramp x{factor / 10:.1f}:"]
    for name, fn in (
        ("pcm16 files", lambda: ctx.features_host_pcm16(files, 1, clip_of, w_starts, lengths, sr, bits)),
        ("float32 buffer", lambda: ctx.features_host(flat32, starts, lengths, sr, bits)),
    ):
        ms, rows = timed(fn)
        same = name not in ref or bool(np.array_equal(ref[name], rows))
        ref.setdefault(name, rows)
        line.append(f"{name} {ms:7.2f} ms ({audio_s / ms * 1e3 / 1e3:6.1f} k audio-s/s, chain {ctx.last_compute_ms():6.2f} ms, rows identical {same})")
    print("  ".join(line), flush=True)
    ctx.close()
# whole clips as a list of float32 arrays (the training loader's shape), c3-like: 1 440 rows
ctx = _native.Context(0)
os.environ["SERB_STAGE_THREADS"] = "4"
ms, rows = timed(lambda: ctx.features_host_clips(clips32, sr, bits))
print(f"float32 clip list (whole clips), default knobs: {ms:7.2f} ms", flush=True)
