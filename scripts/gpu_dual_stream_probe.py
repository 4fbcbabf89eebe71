"""Probe: does running two half-batches concurrently (two contexts, two streams, two host
threads) beat one full-batch launch chain?  Prints ms per c2 step for both."""
import sys, threading, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from ser_b200 import _native, synth
import bench

sr, n_samples, n_clips = 48000, 168000, 1440
wave = synth.batch_audio_torch(n_clips, sr, n_samples, device="cuda", first_index=0).reshape(-1).contiguous()
clip_of, w_starts, lengths, _ = bench.window_plan(n_clips, n_samples, sr)
starts = clip_of * n_samples + w_starts
bits = 31
feats = torch.empty((starts.size, 193), dtype=torch.float32, device="cuda")
ctxs = [_native.Context(0) for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2)]
streams = [torch.cuda.Stream() for _ in ctxs]
steps = 5

def run(parts):
    # parts: list of (ctx, stream, row slice)
    def work(ctx, stream, sl):
        for _ in range(steps):
            ctx.features_device(wave.data_ptr(), wave.numel(), starts[sl], lengths[sl], sr, bits,
                                feats[sl.start:sl.stop].data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=p) for p in parts]
    for t in ths: t.start()
    for t in ths: t.join()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / steps

n = starts.size
for k in range(1, len(ctxs) + 1):
    bounds = [n * i // k // 4 * 4 for i in range(k + 1)]; bounds[-1] = n
    parts = [(ctxs[i], streams[i], slice(bounds[i], bounds[i + 1])) for i in range(k)]
    run(parts)
    ref = feats.clone() if k == 1 else ref
    ms = [run(parts) for _ in range(3)]
    same = bool(torch.equal(ref, feats))
    print(f"{k} concurrent part(s): ms per step {['%.2f' % m for m in ms]}  rows identical to 1-part: {same}", flush=True)
